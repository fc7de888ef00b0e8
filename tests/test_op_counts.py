"""SURVEY 8(d) "authoritative op count": the numerators of the FP64 rooflines (bench.py F2_OPS /
F3_OPS = f64 add + multiply instructions per simplex octave incl. the Fbm accumulate) are pinned
  * for the form the kernels execute: from the SASS ptxas emits for the Fbm loops of
    csrc/vx_noise.cuh (tests/cuda/op_probe.cu, the library's own flags) -- compiler output, not a
    hand count;
  * for the crate's literal form: by the oracle's counting scalar (-DVGO_COUNT_OPS build), i.e. what
    the reference's CPU executes per octave.
CPU only (nvcc cross-compiles, cuobjdump disassembles)."""
import collections
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _loops(sass: str):
    """{kernel: [Counter of opcodes in the body of every backward branch]}"""
    out = {}
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n")[0].strip()
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
        bodies = []
        for addr, text in ins:
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                lo = int(m.group(1), 16)
                bodies.append(collections.Counter(
                    re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for a, t in ins if lo <= a <= addr))
        out[name] = bodies
    return out


def test_executed_form_counts_come_from_the_sass(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.exists(nvcc) and os.path.exists(cuobjdump)):
        pytest.skip("no CUDA toolkit")
    cubin = str(tmp_path / "op_probe.cubin")
    subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
                    "-cubin", "-o", cubin, os.path.join(ROOT, "tests", "cuda", "op_probe.cu")], check=True)
    sass = subprocess.run([cuobjdump, "-sass", cubin], check=True, capture_output=True, text=True).stdout
    assert "DFMA" not in sass  # -fmad=false: every multiply and add rounds on its own
    loops = _loops(sass)
    octave2 = max(loops["probe_fbm2"], key=lambda c: sum(c.values()))  # the other loop copies the tables
    octave3 = max(loops["probe_fbm3"], key=lambda c: sum(c.values()))
    assert octave2["DADD"] + octave2["DMUL"] == bench.F2_OPS == 58, octave2
    assert octave3["DADD"] + octave3["DMUL"] == bench.F3_OPS == 83, octave3
    # what else the FP64 pipe sees per octave, and the issue slots around it
    assert octave2["DSETP"] <= 1 and octave3["DSETP"] <= 7
    assert sum(octave2.values()) <= 92 and sum(octave3.values()) <= 172


_COUNT = r"""
import collections, numpy as np, oracle
assert oracle.use_counting_build()
s = oracle.Simplex(12345)
rng = np.random.default_rng(1)
c2, c3 = collections.Counter(), collections.Counter()
for _ in range(4000):
    x, y, z = rng.uniform(-100, 100, 3)
    oracle.op_counts()
    s.sample2(x, y); c2[oracle.op_counts()] += 1
    s.sample3(x, y, z); c3[oracle.op_counts()] += 1
print((sorted(c2), sorted(c3)))
f2 = oracle.Fbm(s, 6, 0.008, 1.8, 0.45)
f3 = oracle.Fbm(s, 8, 0.04, 1.2, 0.35)
one = oracle.Fbm(s, 1, 0.008, 1.8, 0.45)
oracle.op_counts(); s.sample2(0.3, 0.1); a = oracle.op_counts()
oracle.op_counts(); one.sample2(0.3 / 0.008, 0.1 / 0.008); b = oracle.op_counts()
print((a, b))
"""


def test_crate_form_counts_by_counting_scalar():
    r = subprocess.run([sys.executable, "-c", _COUNT], env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    c2, c3 = eval(lines[0])
    # (adds, multiplies) per simplex evaluation: a corner outside its radius returns before its
    # gradient arithmetic, so the count depends on how many corners contribute (1..3 of 3, 1..4 of 4)
    assert c2 == [(25, 14), (26, 19), (27, 24)]
    assert c3 == [(48, 21), (50, 27), (52, 33), (54, 39)]
    (a_add, a_mul), (b_add, b_mul) = eval(lines[1])
    # one Fbm octave around it: noise += amp * s (1 add, 1 multiply), point *= lacunarity (2 in 2-D),
    # amp *= persistence (1); once per Fbm: point *= frequency (2), * normalisation (1)
    assert (b_add - a_add, b_mul - a_mul) == (1, 4 + 2 + 1)
    # => the crate executes at most 27 + 24 + 1 + 4 = 56 per 2-D and 54 + 39 + 1 + 5 = 99 per 3-D octave;
    # the kernels issue 58 and 83 (floor as two adds, amplitudes precomputed, 3-D gradient dot product
    # as one signed add; every corner evaluated, outside its radius it adds +-0)
