"""CPU-side checks of the drop-in boundary and the host logic: the C-ABI library loads and exports
every symbol include/voxel_cuda.h declares, fails loudly without a GPU (no CPU fallback), the
product tree never touches oracle/, the golden fixtures still match the oracle, and the multi-GPU
sharding logic is exercised with a world_size-2 gloo group.  No compute call is made on the
library here (there is no GPU in this container)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import oracle  # noqa: E402
import voxel_game_b200 as vg  # noqa: E402


# ------------------------------------------------------------------------------ ABI surface
def _header_functions():
    text = open(os.path.join(ROOT, "include", "voxel_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = vg.load()
    declared = _header_functions()
    assert declared == sorted(vg.ABI_SYMBOLS), "binding.ABI_SYMBOLS out of sync with voxel_cuda.h"
    for name in declared:
        assert hasattr(lib, name), f"libvoxel_cuda.so does not export {name}"
    # and via nm: they are real dynamic symbols of the in-tree .so, C linkage (no mangling)
    out = subprocess.run(["nm", "-D", "--defined-only", vg.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(vx_[a-z0-9_]+)$", out, flags=re.M))
    assert set(declared) <= exported


def test_rust_shim_declares_every_header_symbol_and_guards_the_pairs():
    """rust/voxel_cuda/src/lib.rs cannot be compiled here (no cargo): at least its extern block names
    every function of the header, and every wrapper that makes a query + read pair holds a
    Transaction (vx_lock / vx_unlock) across it (Pipeline is Sync)."""
    src = open(os.path.join(ROOT, "rust", "voxel_cuda", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (vx_[a-z0-9_]+)\s*\(", src))
    optional = {"vx_set_timing", "vx_get_timings", "vx_diag_fp64_rate"}  # diagnostics the shim does not bind
    assert set(_header_functions()) - optional <= declared, sorted(set(_header_functions()) - optional - declared)
    for read_call in ("vx_mesh_read(", "vx_mesh_read_compact(", "vx_encode_read(", "vx_visible_read(", "vx_update_read("):
        for m in re.finditer(re.escape("            " + read_call) + "|" + re.escape("unsafe { " + read_call), src):
            body_start = src.rfind("pub fn ", 0, m.start())
            assert "self.transaction()" in src[body_start:m.start()], f"{read_call} used without a Transaction"
    assert "vx_copy_last_error" in src


def test_variant_libraries_are_current():
    """The VX_NOISE_VARIANT builds (parity-pin kit) come from the same sources as the default library:
    every declared symbol is exported, so a stale variant cannot sneak onto the GPU box."""
    from voxel_game_b200.binding import VARIANT_LIBS

    declared = set(_header_functions())
    for mask, path in VARIANT_LIBS.items():
        assert os.path.exists(path), f"{path} missing: __graft_entry__.build()"
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        exported = set(re.findall(r"\sT\s+(vx_[a-z0-9_]+)$", out, flags=re.M))
        assert declared <= exported, (mask, sorted(declared - exported))
        assert os.path.getmtime(path) >= os.path.getmtime(vg.LIB_PATH) - 600, f"{path} is older than the default library"


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", vg.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header():
    assert vg.REC_DTYPE.itemsize == 16 and vg.VERTEX_DTYPE.itemsize == 40 and vg.EDIT_DTYPE.itemsize == 16
    assert vg.VERTEX_DTYPE.fields["uv"][1] == 12 and vg.VERTEX_DTYPE.fields["color"][1] == 20
    assert vg.VERTEX_DTYPE.fields["normal"][1] == 24
    from voxel_game_b200.binding import MeshInfo, Timings

    assert C.sizeof(MeshInfo) == 40
    assert C.sizeof(Timings) == 9 * 4 + 5 * 4 + 4 * 8
    from voxel_game_b200.binding import View, VisibleInfo

    assert C.sizeof(View) == 40 and C.sizeof(VisibleInfo) == 24 + 28 * 4 + 64 * 4


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    lib = vg.load()
    if lib.vx_device_count() > 0:
        pytest.skip("a GPU is present")
    assert lib.vx_device_count() == -1  # VX_ERR_NO_DEVICE
    with pytest.raises(vg.VxError) as e:
        vg.Context(0)
    assert e.value.status == -1 and "no CPU fallback" in str(e.value)
    h = C.c_void_p()
    assert lib.vx_create(0, C.byref(h)) == -1 and not h.value
    assert b"no CPU fallback" in lib.vx_strerror(-1)
    assert lib.vx_strerror(-5) == b"area not resident"


def test_world_name_hash_matches_oracle():
    """Two independent SipHash-1-3 implementations (csrc/vx_tables.cpp, oracle/noise_libnoise.c)."""
    for name in ["test", "abc", "a world name of 20ch", "1234567", "12345678", "üñí"]:
        assert vg.hash_world_name(name) == oracle.hash_world_name(name)


def test_product_tree_never_references_the_oracle():
    """The shipped path must not import, include, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "voxel_game_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"\boracle\b|vg_oracle|liboracle|vgo_", text):
                    offenders.append(os.path.join(base, f))
    hdr = open(os.path.join(ROOT, "include", "voxel_cuda.h")).read()
    assert not re.search(r"vgo_|liboracle", hdr)
    assert not offenders, offenders
    out = subprocess.run(["ldd", vg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def _make_words(text, name):
    """Words of `NAME := ...` in a Makefile, continuation lines included."""
    m = re.search(rf"^{name}\s*[:?]?=\s*((?:.*\\\n)*.*)$", text, flags=re.M)
    assert m, name
    return m.group(1).replace("\\\n", " ").split()


def test_source_lists_agree():
    """One list of translation units: csrc/Makefile.  rust/voxel_cuda/build.rs drives that Makefile
    (it must not carry a list of its own), and the nvcc line INTEGRATION.md shows names exactly the
    Makefile's sources -- a unit missing from either used to leave the documented build unlinkable."""
    csrc = os.path.join(ROOT, "voxel_game_b200", "csrc")
    mk = open(os.path.join(csrc, "Makefile")).read()
    src = _make_words(mk, "SRC")
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith((".cu", ".cpp")))
    assert sorted(src) == on_disk, "csrc/Makefile SRC does not list every .cu/.cpp in csrc/"
    hdr = _make_words(mk, "HDR")
    hdr_on_disk = sorted(f for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp")))
    assert sorted(h for h in hdr if not h.startswith("..")) == hdr_on_disk
    assert "--no-undefined" in mk
    build_rs = open(os.path.join(ROOT, "rust", "voxel_cuda", "build.rs")).read()
    assert 'Command::new("make")' in build_rs and "voxel_game_b200/csrc" in build_rs
    assert not re.findall(r"vx_[a-z_]+\.(?:cu|cpp)\b", build_rs), "build.rs must not keep its own source list"
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```\nnvcc .*?```", integ, flags=re.S)
    assert block, "INTEGRATION.md no longer shows the nvcc command"
    assert sorted(re.findall(r"vx_[a-z_]+\.(?:cu|cpp)\b", block.group(0))) == sorted(src)
    # every extern "C" function of the Rust shim is declared in the header
    lib_rs = open(os.path.join(ROOT, "rust", "voxel_cuda", "src", "lib.rs")).read()
    rust_fns = set(re.findall(r"pub fn (vx_[a-z0-9_]+)\s*\(", lib_rs))
    assert rust_fns <= set(_header_functions()), rust_fns - set(_header_functions())


# ------------------------------------------------------------------------------ golden fixtures
GOLDEN = ["spawn_8x8_test.npz", "caves_6x6_test.npz", "far_4x4_seed2.npz", "lake_4x4_test.npz", "cold_4x4_test.npz"]


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_reproduces_golden_fixture(name):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden

    g = np.load(os.path.join(ROOT, "tests", "golden", name))
    fresh = make_golden.region_case(str(g["world_name"]), int(g["origin"][0]), int(g["origin"][1]), int(g["n"]))
    assert int(fresh["seed"]) == int(g["seed"])
    assert np.array_equal(fresh["rows"], g["rows"])
    assert np.array_equal(fresh["stats"], g["stats"])


def test_golden_regions_cover_caves_lakes_cold_and_trees():
    """The fixtures exercise every generation branch (cave carving, lake override incl. the ice cap,
    cold/dry/wet surfaces, trees)."""
    st = {n: np.load(os.path.join(ROOT, "tests", "golden", n))["stats"] for n in GOLDEN}
    assert st["caves_6x6_test.npz"][2] > 10_000          # cave voxels evaluated
    assert all(s[4] > 0 for s in st.values())            # trees placed
    seed = oracle.hash_world_name("test")
    v, _, _, _ = oracle.generate_areas(seed, oracle.region_xy(62507, 62503, 6, 6))
    assert (v == oracle.V["WaterSource"]).sum() > 1000
    v, _, _, _ = oracle.generate_areas(seed, oracle.region_xy(62485, 62491, 6, 6))
    assert (v == oracle.V["Ice"]).sum() > 0 and (v == oracle.V["Snow"]).sum() > 0
    area = np.load(os.path.join(ROOT, "tests", "golden", "area_62500_62500_test.npz"))
    vox, mh, _ = oracle.Generator(int(area["seed"])).generate_area(62500, 62500)
    assert np.array_equal(vox, area["voxels"]) and np.array_equal(mh, area["max_height"])


# ------------------------------------------------------------------------------ sharding (N > 1)
def test_balanced_rows_partition_by_cost():
    """Equal-cost strips for the sharded step: contiguous, non-empty, covering every row once; on
    skewed costs the largest strip cost (ring rows included) beats the equal-height cut."""
    rng = np.random.default_rng(0)
    for nx, ny, n_ranks in ((128, 128, 8), (128, 128, 3), (16, 9, 9), (32, 40, 5)):
        cc = np.zeros((ny + 2, nx + 2))
        cc[ny // 3: ny // 2, nx // 4: 3 * nx // 4] = 256 * rng.random((ny // 2 - ny // 3, 3 * nx // 4 - nx // 4))
        gen, mesh = vg.row_costs(cc.ravel(), 40 * cc.ravel(), nx, ny)
        strips = vg.balanced_rows(gen, mesh, n_ranks)
        assert len(strips) == n_ranks and strips[0][0] == 0 and strips[-1][1] == ny
        assert all(a < b for a, b in strips) and all(strips[i][1] == strips[i + 1][0] for i in range(n_ranks - 1))
        worst = max(vg.strip_cost(gen, mesh, a, b) for a, b in strips)
        equal = max(vg.strip_cost(gen, mesh, *vg.shard_rows(ny, n_ranks, r)) for r in range(n_ranks))
        assert worst <= equal * 1.0001
        if ny >= 4 * n_ranks:
            assert worst < 0.8 * equal
    # uniform cost: the equal-height cut (up to the ring rows, which make the end strips a little cheaper)
    gen, mesh = vg.row_costs(np.zeros(130 * 130), np.zeros(130 * 130), 128, 128)
    assert all(14 <= b - a <= 18 for a, b in vg.balanced_rows(gen, mesh, 8))


def test_shard_rows_partition():
    for ny, ws in [(128, 1), (128, 2), (128, 8), (130, 8), (7, 3), (5, 8)]:
        spans = [vg.shard_rows(ny, ws, r) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == ny
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _shard_worker(rank, world_size, port, n, out_dir):
    """One rank of the sharded pipeline, CPU stand-in for the per-GPU work: generate own strip plus
    the regenerated halo ring, mesh the strip, all_gather the per-area face counts (the optional
    final gather; there is no data-path collective)."""
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    seed = oracle.hash_world_name("test")
    ax0, ay0 = 62496, 62496
    lo, hi = vg.shard_rows(n, world_size, rank)
    rows = hi - lo
    world = oracle.World.generate(seed, ax0 - 1, ay0 + lo - 1, n + 2, rows + 2, threads=1)  # strip + ring
    mine = vg.region_xy(ax0, ay0 + lo, n, rows)
    faces = torch.tensor([world.mesh_area(int(ax), int(ay)) for ax, ay in mine], dtype=torch.int64)
    vox_digest = torch.tensor([int(world.voxels[world.slot(int(ax), int(ay))].astype(np.uint64).sum()) for ax, ay in mine],
                              dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(sizes, torch.tensor([len(mine)], dtype=torch.int64))
    pad = int(max(s.item() for s in sizes))
    buf = torch.full((2, pad), -1, dtype=torch.int64)
    buf[0, :len(mine)] = faces
    buf[1, :len(mine)] = vox_digest
    gathered = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(gathered, buf)
    # timing reduction as bench.py does it: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.cat([g[:, :int(s.item())] for g, s in zip(gathered, sizes)], dim=1).numpy()
        np.save(os.path.join(out_dir, "gathered.npy"), full)
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_strips_with_regenerated_halo_equal_unsharded(tmp_path):
    """world_size 2 over gloo: every rank regenerates its own one-area halo ring from the seed, so
    the union of the strips equals the single-process result (areas are pure functions of seed and
    coordinates; trees never cross an area border, trees.rs:214-222)."""
    import torch.multiprocessing as mp

    n = 6
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    assert float(np.load(tmp_path / "tmax.npy")[0]) == 2.0
    seed = oracle.hash_world_name("test")
    world = oracle.World.generate(seed, 62495, 62495, n + 2, n + 2)
    xy = vg.region_xy(62496, 62496, n, n)
    faces = [world.mesh_area(int(ax), int(ay)) for ax, ay in xy]
    dig = [int(world.voxels[world.slot(int(ax), int(ay))].astype(np.uint64).sum()) for ax, ay in xy]
    assert got.shape == (2, n * n)
    assert got[0].tolist() == faces and got[1].tolist() == dig


# ------------------------------------------------------------------------------ host expander
def test_host_library_exports_and_links_the_cuda_library():
    L = vg.load_host()
    for name in vg.HOST_SYMBOLS:
        assert hasattr(L, name), name
    out = subprocess.run(["ldd", vg.HOST_LIB_PATH], capture_output=True, text=True).stdout
    assert "libvoxel_cuda.so" in out and "oracle" not in out


def _compact_from_oracle(world, xy):
    """Face masks of the oracle's meshes -> the 4-byte records vx_mesh_read_compact would deliver."""
    rec_first = np.zeros(len(xy) + 1, np.uint32)
    packed = []
    for i, (ax, ay) in enumerate(xy):
        s = world.slot(int(ax), int(ay))
        local = np.flatnonzero(world.face_mask[s]).astype(np.uint32)
        packed.append(local | (world.voxels[s][local].astype(np.uint32) << 15) | (world.face_mask[s][local].astype(np.uint32) << 20))
        rec_first[i + 1] = rec_first[i] + len(local)
    return rec_first, np.concatenate(packed)


@pytest.mark.parametrize("case", ["terrain", "all_types"])
def test_host_expander_equals_oracle_emission(case):
    """The C++ expander of the compact transport (host/voxel_game.hpp expand_voxel) against the
    oracle's get_verticies_for_voxel restatement (oracle/mesh.c, mesh_generator.rs:200-498): records,
    vertices and indices byte for byte -- generated terrain, and random areas with all 27 voxel types
    (partial-height water offsets, voxels with different faces)."""
    if case == "terrain":
        world = oracle.World.generate(oracle.hash_world_name("test"), 62499, 62499, 4, 4)
        xy = oracle.region_xy(62500, 62500, 2, 2)
    else:
        rng = np.random.default_rng(9)
        vox = rng.integers(0, 27, (9, 32768)).astype(np.uint8)
        vox[rng.random(vox.shape) < 0.7] = 0
        mh = np.stack([oracle.update_all_column_heights(v) for v in vox])
        world = oracle.World(70000, 123, 3, 3, vox, mh)
        xy = np.array([[70001, 124]], np.uint32)
    base, orecs, overts, oidx = 0, [], [], []
    for ax, ay in xy:
        nf = world.mesh_area(int(ax), int(ay))
        r, v, i = world.emit_area(int(ax), int(ay), first_face_base=base)
        orecs.append(r), overts.append(v), oidx.append(i)
        base += nf
    rec_first, packed = _compact_from_oracle(world, xy)
    for threads in (1, 3):
        ff, recs, verts, idx = vg.expand_records(xy, rec_first, packed, threads=threads)
        assert ff[-1] == base
        assert recs.tobytes() == np.concatenate(orecs).tobytes()
        assert verts.tobytes() == np.concatenate(overts).tobytes()
        assert np.array_equal(idx, np.concatenate(oidx))
    assert vg.unpack_records(xy, rec_first, packed).tobytes() == recs.tobytes()
    if case == "all_types":
        kinds = set(np.unique((packed >> 15) & 0x1F).tolist())
        assert kinds == set(range(1, 27)) and base > 20000
