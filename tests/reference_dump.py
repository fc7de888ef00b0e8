"""Matcher for a parity dump of the reference (rust/parity_dump/parity_dump.rs): which setting of
the libnoise-calibration switches reproduces it?  Test infrastructure (uses the CPU oracle)."""
from __future__ import annotations

import itertools
import json

import numpy as np

import oracle

NORM_2D, NORM_3D = 99.83685446303647, 76.88375854620836
FLAG_CANDIDATES = (0, 1, 8, 9)  # Fbm form x skew digits; the tie-breaks (2, 4) are value-neutral


def _f(bits) -> np.ndarray:
    return np.asarray(bits, dtype=np.uint64).view(np.float64)


def _bits(values) -> list:
    return [int(v) for v in np.asarray(values, dtype=np.float64).view(np.uint64)]


# ---------------------------------------------------------------------------------- synthetic dumps
def synthesize(flags=0, order2=None, order3=None, world_name="test", areas=((62500, 62500), (123, 456)), n=120):
    """A dump with the layout of the Rust harness, made by the oracle under a given setting."""
    rng = np.random.default_rng(99)
    seed = oracle.hash_world_name(world_name)
    M = (1 << 64) - 1
    seeds = [seed, seed ^ M, (seed + 1) & M, (seed + 10) & M]
    oracle.set_noise_variant(flags, order2, order3)
    try:
        simplex = []
        for s in seeds:
            g = oracle.Simplex(s)
            for dim in (2, 3):
                pts = np.concatenate([rng.uniform(-4, 4, (n, dim)), rng.uniform(79000, 81000, (n, dim)),
                                      np.repeat(rng.uniform(-50, 50, (n // 4, 1)), dim, axis=1)])
                vals = [g.sample2(*p) if dim == 2 else g.sample3(*p) for p in pts]
                simplex.append({"seed": s, "dim": dim, "points": _bits(pts.ravel()), "values": _bits(vals)})
                h, un = 0.01, (0.211324865405187 if dim == 2 else 1.0 / 6.0)
                side = 16 if dim == 2 else 7
                lattice, ppts = [], []
                for idx in itertools.product(range(side), repeat=dim):
                    c = [i - sum(idx) * un for i in idx]
                    for axis in range(dim):
                        lattice += list(idx) + [axis]
                        ppts.append([c[d] + (h if d == axis else 0.0) for d in range(dim)])
                pv = [g.sample2(*p) if dim == 2 else g.sample3(*p) for p in ppts]
                simplex.append({"seed": s, "dim": dim, "probe_h": _bits([h])[0], "lattice": lattice,
                                "points": _bits(np.asarray(ppts).ravel()), "values": _bits(pv)})
        fbm = []
        for name, s, dim, octv, f, l, p in (("height1", seeds[0], 2, 6, 0.008, 1.8, 0.45), ("height_mod", seeds[2], 2, 2, 0.004, 0.004, 0.2),
                                            ("cave2", seeds[1], 3, 8, 0.04, 1.2, 0.35), ("alt", seeds[2], 3, 2, 0.07, 2.0, 0.3),
                                            ("lake", seeds[3], 2, 2, 0.004, 1.3, 0.35)):
            g = oracle.Fbm(oracle.Simplex(s), octv, f, l, p)
            pts = np.floor(rng.uniform(980000, 1020000, (n, dim)))
            if dim == 3:
                pts[:, 2] = np.floor(rng.uniform(1, 128, n))
            vals = [g.sample2(*q) if dim == 2 else g.sample3(*q) for q in pts]
            fbm.append({"name": name, "seed": s, "dim": dim, "octaves": octv, "frequency": _bits([f])[0],
                        "lacunarity": _bits([l])[0], "persistence": _bits([p])[0], "points": _bits(pts.ravel()), "values": _bits(vals)})
        xy = np.asarray(areas, np.uint32)
        vox, mh, _, _ = oracle.generate_areas(seed, xy)
        out_areas = [{"x": int(a), "y": int(b), "voxels": vox[i].tobytes().hex(), "max_height": mh[i].tobytes().hex()}
                     for i, (a, b) in enumerate(xy)]
    finally:
        oracle.set_noise_variant()
    return {"format": 1, "crate": "synthetic (oracle)", "world_name": world_name, "seed": seed, "simplex": simplex,
            "fbm": fbm, "areas": out_areas}


# ---------------------------------------------------------------------------------- the matcher
def _derive_order(entry, canonical):
    """Gradient-table order from the corner probes of one (seed, dim): -> (order or None, problems)."""
    dim = entry["dim"]
    norm = NORM_2D if dim == 2 else NORM_3D
    h = float(_f([entry["probe_h"]])[0])
    lattice = np.asarray(entry["lattice"], np.int64).reshape(-1, dim + 1)
    values = _f(entry["values"])
    perm = oracle.Simplex(entry["seed"]).perm.astype(np.int64)
    n_grad = len(canonical)
    scale = norm * (0.5 - h * h) ** 4 * h
    order, problems = {}, []
    for c in range(0, len(lattice), dim):
        idx = lattice[c, :dim]
        g = values[c:c + dim] / scale  # the corner's gradient, component per probed axis
        k = int(np.argmax(canonical @ g))
        if not np.allclose(canonical[k], g, atol=1e-6):
            problems.append(f"corner {tuple(idx)}: gradient {g} is not in the canonical set")
            continue
        hashed = int(perm[idx[0]])
        for d in range(1, dim):
            hashed = int(perm[idx[d] + hashed])
        gi = hashed % n_grad
        if order.setdefault(gi, k) != k:
            problems.append(f"corner {tuple(idx)}: table entry {gi} maps to both {order[gi]} and {k} -- permutation table mismatch")
    if problems or len(order) != n_grad or sorted(order.values()) != list(range(n_grad)):
        if not problems:
            problems.append(f"only {len(order)} of {n_grad} table entries seen, or not a bijection")
        return None, problems
    return np.array([order[i] for i in range(n_grad)], np.uint8), []


def analyse(dump: dict) -> dict:
    """-> {"seed_ok", "order2", "order3", "problems", "flags": matching flag sets, "report": text}."""
    report, problems = [], []
    seed = oracle.hash_world_name(dump["world_name"])
    seed_ok = seed == int(dump["seed"])
    report.append(f"world-name hash: {'ok' if seed_ok else 'DIFFERS'} ({seed:#x} vs {int(dump['seed']):#x})")
    grad2 = oracle.grad2_table().reshape(24, 2)
    grad3 = oracle.grad3_table().reshape(12, 3)
    orders = {2: None, 3: None}
    for e in dump["simplex"]:
        if "probe_h" not in e:
            continue
        o, p = _derive_order(e, grad2 if e["dim"] == 2 else grad3)
        if o is None:
            problems += [f"seed {e['seed']:#x} {e['dim']}-D: {x}" for x in p[:3]]
        elif orders[e["dim"]] is None:
            orders[e["dim"]] = o
        elif not np.array_equal(orders[e["dim"]], o):
            problems.append(f"{e['dim']}-D gradient order differs between seeds")
    for dim in (2, 3):
        o = orders[dim]
        report.append(f"{dim}-D gradient table order: " + ("not derivable" if o is None else
                      ("canonical" if np.array_equal(o, np.arange(len(o))) else str(o.tolist()))))
    matching = []
    if orders[2] is not None and orders[3] is not None:
        xy = np.array([[a["x"], a["y"]] for a in dump["areas"]], np.uint32)
        want_vox = np.stack([np.frombuffer(bytes.fromhex(a["voxels"]), np.uint8) for a in dump["areas"]])
        want_mh = np.stack([np.frombuffer(bytes.fromhex(a["max_height"]), np.uint8) for a in dump["areas"]])
        for flags in FLAG_CANDIDATES:
            oracle.set_noise_variant(flags, orders[2], orders[3])
            try:
                first = None
                for e in dump["simplex"]:
                    g = oracle.Simplex(e["seed"])
                    pts = _f(e["points"]).reshape(-1, e["dim"])
                    got = np.array([g.sample2(*p) if e["dim"] == 2 else g.sample3(*p) for p in pts])
                    bad = np.flatnonzero(got.view(np.uint64) != np.asarray(e["values"], np.uint64))
                    if bad.size and first is None:
                        first = f"Simplex<{e['dim']}> seed {e['seed']:#x} at {pts[bad[0]].tolist()}: {got[bad[0]]!r} vs {_f(e['values'])[bad[0]]!r} ({bad.size} of {len(got)})"
                for e in dump["fbm"]:
                    g = oracle.Fbm(oracle.Simplex(e["seed"]), e["octaves"], *[float(x) for x in _f([e["frequency"], e["lacunarity"], e["persistence"]])])
                    pts = _f(e["points"]).reshape(-1, e["dim"])
                    got = np.array([g.sample2(*p) if e["dim"] == 2 else g.sample3(*p) for p in pts])
                    bad = np.flatnonzero(got.view(np.uint64) != np.asarray(e["values"], np.uint64))
                    if bad.size and first is None:
                        first = f"Fbm {e['name']} at {pts[bad[0]].tolist()}: {got[bad[0]]!r} vs {_f(e['values'])[bad[0]]!r} ({bad.size} of {len(got)})"
                vox, mh, _, _ = oracle.generate_areas(seed, xy)
                n_bad = int((vox != want_vox).sum()) + int((mh != want_mh).sum())
                if n_bad and first is None:
                    a = int(np.flatnonzero((vox != want_vox).any(axis=1) | (mh != want_mh).any(axis=1))[0])
                    first = f"area {tuple(xy[a])}: {n_bad} block IDs / heights differ in the dump's areas"
            finally:
                oracle.set_noise_variant()
            report.append(f"flags {flags}: " + ("REPRODUCES THE DUMP bit for bit" if first is None else "differs -- " + first))
            if first is None:
                matching.append(flags)
    report += ["problem: " + p for p in problems]
    return {"seed_ok": seed_ok, "order2": orders[2], "order3": orders[3], "problems": problems, "flags": matching,
            "report": "\n".join(report)}


def load(path: str) -> dict:
    with open(path) as f:
        return json.load(f)
