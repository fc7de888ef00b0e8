"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seed and region.
Block IDs, column heights, records, vertices and indices are compared byte for byte."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg

pytestmark = pytest.mark.gpu

SPAWN = 62500


def _first_diff(a, b):
    d = np.flatnonzero(a.ravel() != b.ravel())
    return None if d.size == 0 else int(d[0])


@pytest.mark.parametrize("origin,n", [((SPAWN - 4, SPAWN - 4), 8), ((123, 456), 3), ((SPAWN - 16, SPAWN - 16), 32)])
def test_generate_block_ids_and_heights_bit_exact(ctx, seed, origin, n):
    """AreaGenerator::generate_area for an n x n region: all block IDs + heights byte-exact.
    The 32x32 case is BASELINE.json configs[1] (33 554 432 block IDs)."""
    xy = vg.region_xy(origin[0], origin[1], n, n)
    vox, mh = ctx.generate(xy)
    ovox, omh, stats, _ = oracle.generate_areas(seed, xy)
    bad = np.flatnonzero((vox != ovox).any(axis=1))
    assert bad.size == 0, (f"{bad.size} areas differ, first area {xy[bad[0]]}, first voxel "
                           f"{_first_diff(vox[bad[0]], ovox[bad[0]])}")
    assert np.array_equal(mh, omh)
    if n == 32:
        assert stats["cave_columns"] > 0 and stats["trees_placed"] > 0  # the region exercises both


def test_mesh_stream_byte_exact(ctx, seed):
    """Renderer::load_all_blocking over 8x8 areas: records, vertices and indices equal the
    oracle's stream (same voxel order z,y,x and face push order), byte for byte."""
    n = 8
    ax0, ay0 = SPAWN - 4, SPAWN - 4
    world = oracle.World.generate(seed, ax0 - 1, ay0 - 1, n + 2, n + 2)
    ctx.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
    xy = vg.region_xy(ax0, ay0, n, n)
    info = ctx.mesh(xy)
    rec_first, face_first, recs, verts, idx = ctx.mesh_read(info)
    base = 0
    for i, (ax, ay) in enumerate(xy):
        nf = world.mesh_area(int(ax), int(ay))
        orecs, overts, oidx = world.emit_area(int(ax), int(ay), first_face_base=base)
        assert face_first[i] == base and face_first[i + 1] == base + nf
        r = recs[rec_first[i]:rec_first[i + 1]]
        assert r.tobytes() == orecs.tobytes(), f"records differ in area {ax},{ay}"
        v = verts[4 * base:4 * (base + nf)]
        assert v.tobytes() == overts.tobytes(), f"vertices differ in area {ax},{ay}"
        assert np.array_equal(idx[6 * base:6 * (base + nf)], oidx)
        base += nf
    assert info["n_faces"] == base


def test_mesh_stream_byte_exact_config2_region(ctx, seed):
    """The whole 32 x 32 region of config 2 (1 024 areas, ~600 000 faces, caves, lakes, trees, all three
    biomes): records, vertices and indices of every area equal the oracle's stream byte for byte, in
    the reference's order.  (At config-3 size the mesh is checked through sampled face counts and
    properties, test_full_size_128x128_properties.)"""
    n = 32
    ax0, ay0 = SPAWN - 16, SPAWN - 16
    world = oracle.World.generate(seed, ax0 - 1, ay0 - 1, n + 2, n + 2)
    ctx.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
    faces = _compare_mesh(ctx, world, vg.region_xy(ax0, ay0, n, n))
    assert faces > 400_000


def _compare_mesh(ctx, world, xy):
    info = ctx.mesh(xy)
    rec_first, face_first, recs, verts, idx = ctx.mesh_read(info)
    base = 0
    for i, (ax, ay) in enumerate(xy):
        nf = world.mesh_area(int(ax), int(ay))
        orecs, overts, oidx = world.emit_area(int(ax), int(ay), first_face_base=base)
        assert face_first[i] == base and face_first[i + 1] == base + nf, f"face count differs in area {ax},{ay}"
        assert recs[rec_first[i]:rec_first[i + 1]].tobytes() == orecs.tobytes(), f"records differ in area {ax},{ay}"
        assert verts[4 * base:4 * (base + nf)].tobytes() == overts.tobytes(), f"vertices differ in area {ax},{ay}"
        assert np.array_equal(idx[6 * base:6 * (base + nf)], oidx)
        base += nf
    assert info["n_faces"] == base
    return base


@pytest.mark.parametrize("p_none", [0.5, 0.9, 0.1])
def test_mesh_random_voxels_all_types(seed, p_none):
    """Uploaded (disk-loaded) areas with every one of the 27 voxel types, including glass, ice, all
    six water kinds next to each other: exercises the cur != nbr rule for transparent voxels, the
    partial-height top-face rule and the enclosed-voxel pre-filter (area.rs:126-201)."""
    rng = np.random.default_rng(int(p_none * 100))
    n = 5
    ax0, ay0 = 1000, 2000
    vox = rng.integers(1, 27, size=(n * n, 32768), dtype=np.uint8)
    vox[rng.random(vox.shape) < p_none] = 0
    # a solid block of partial-height water so that interior enclosed Water1..4 voxels exist
    blk = vox[12].reshape(128, 16, 16)
    blk[40:60, 3:13, 3:13] = oracle.V["Water2"]
    blk[45:50, 5:9, 5:9] = oracle.V["Water3"]
    xy_all = vg.region_xy(ax0, ay0, n, n)
    with vg.Context(0, seed=seed) as c:
        mh = c.upload(xy_all, vox)
        omh = np.stack([oracle.update_all_column_heights(v) for v in vox])
        assert np.array_equal(mh, omh)
        world = oracle.World(ax0, ay0, n, n, vox, omh)
        faces = _compare_mesh(c, world, vg.region_xy(ax0 + 1, ay0 + 1, n - 2, n - 2))
        assert faces > 0


def test_mesh_checkerboard_maximum_faces(seed):
    """Worst case: a 3-D checkerboard has 6 faces on every second voxel (98 304 faces per area),
    far beyond the shared-memory face list, so the windowed flush path is exercised."""
    n = 3
    ax0, ay0 = 500, 500
    z, y, x = np.meshgrid(np.arange(128), np.arange(16), np.arange(16), indexing="ij")
    board = (((x + y + z) & 1) * oracle.V["Brick"]).astype(np.uint8).ravel()
    vox = np.zeros((n * n, 32768), np.uint8)
    vox[4] = board
    vox[5] = board  # a second one next to it: faces across the area border
    xy_all = vg.region_xy(ax0, ay0, n, n)
    with vg.Context(0, seed=seed) as c:
        c.upload(xy_all, vox, read_heights=False)
        omh = np.stack([oracle.update_all_column_heights(v) for v in vox])
        world = oracle.World(ax0, ay0, n, n, vox, omh)
        faces = _compare_mesh(c, world, vg.region_xy(ax0 + 1, ay0 + 1, 1, 1))
        assert faces == 16384 * 6 - 256  # 6 faces per solid voxel, minus Down at z=127 (128) and Up at z=0 (128)


def test_work_counters_match_oracle(ctx, seed):
    """The library's work accounting (roofline numerators) equals the oracle's evaluation counts."""
    xy = vg.region_xy(SPAWN - 16, SPAWN - 16, 32, 32)
    ctx.set_timing(True)
    ctx.generate(xy, read=False)
    t = ctx.timings()
    ctx.set_timing(False)
    _, _, stats, _ = oracle.generate_areas(seed, xy)
    # 2-D evaluations: the cave-zone and biome noises skip their second octave where it cannot
    # change the verdict (at most 2 of the reference's evaluations per column)
    columns = len(xy) * 256
    assert stats["simplex2_evals"] - 2 * columns <= t["simplex2_evals"] < stats["simplex2_evals"]
    assert t["cave_voxels"] == stats["cave_voxels"]
    assert t["cave_columns"] == stats["cave_columns"]
    # 3-D evaluations: the kernel skips the alternative-voxel noise where a lake overrides the voxel
    # anyway (the reference evaluates it and discards the result), so it may do fewer, never more
    # and a cave voxel stops evaluating octaves once its verdict is certain (2..14 of 14)
    assert 2 * stats["cave_voxels"] <= t["cave_evals"] <= 14 * stats["cave_voxels"]
    assert t["cave_evals"] < t["simplex3_evals"] <= stats["simplex3_evals"]


# ------------------------------------------------------------------------------ golden fixtures
import hashlib
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ["spawn_8x8_test.npz", "caves_6x6_test.npz", "far_4x4_seed2.npz", "lake_4x4_test.npz", "cold_4x4_test.npz"]


def _digest(a):
    return int(np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint64)[0])


@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_path_reproduces_golden_fixture(name):
    """Committed fixtures (tests/golden/make_golden.py): block IDs, heights, records, vertices and
    indices per area, as digests.  Needs nothing but the .npz on the GPU box."""
    g = np.load(os.path.join(GOLDEN_DIR, name))
    ax0, ay0, n = int(g["origin"][0]), int(g["origin"][1]), int(g["n"])
    with vg.Context(0, seed=int(g["seed"])) as c:
        assert c.seed == vg.hash_world_name(str(g["world_name"]))
        inner = vg.region_xy(ax0, ay0, n, n)
        vox, mh = c.generate(inner)           # ring NOT generated here: vx_mesh must do it itself
        info = c.mesh(inner)
        assert c.resident_count() == n * n + 4 * n  # the four edge strips were generated on demand
        rec_first, face_first, recs, verts, idx = c.mesh_read(info)
    for i, row in enumerate(g["rows"]):
        assert (int(row[0]), int(row[1])) == (int(inner[i][0]), int(inner[i][1]))
        assert _digest(vox[i]) == int(row[2]), f"block IDs of area {inner[i]}"
        assert _digest(mh[i]) == int(row[3]), f"heights of area {inner[i]}"
        f0, f1 = int(face_first[i]), int(face_first[i + 1])
        r = recs[rec_first[i]:rec_first[i + 1]].copy()
        r["first_face"] -= f0                  # fixtures store per-area face indices
        assert f1 - f0 == int(row[4]) and len(r) == int(row[5])
        assert _digest(r) == int(row[6]) and _digest(verts[4 * f0:4 * f1]) == int(row[7])
        assert _digest(idx[6 * f0:6 * f1]) == int(row[8])


# ------------------------------------------------------------------------------ light / residency / errors
def test_column_heights_operator_and_upload_roundtrip(ctx, seed):
    """Area::update_all_column_heights as an operator (area.rs:34-56), upload -> read back, unload."""
    rng = np.random.default_rng(5)
    vox = rng.integers(0, 27, size=(7, 32768), dtype=np.uint8)
    vox[0] = 0                                   # empty area: every height 127
    vox[1] = oracle.V["Glass"]                   # only transparent voxels: 127 too
    vox[2].reshape(128, 256)[:100] = 0           # first opaque voxel deep down
    want = np.stack([oracle.update_all_column_heights(v) for v in vox])
    assert np.array_equal(ctx.column_heights(vox), want)
    assert (want[0] == 127).all() and (want[1] == 127).all()
    xy = vg.region_xy(40000, 40000, 7, 1)
    before = ctx.resident_count()
    assert np.array_equal(ctx.upload(xy, vox), want)
    v2, h2 = ctx.read_areas(xy[::-1])            # any order
    assert np.array_equal(v2, vox[::-1]) and np.array_equal(h2, want[::-1])
    ctx.unload(xy)
    assert ctx.resident_count() == before
    with pytest.raises(vg.VxError) as e:
        ctx.read_areas(xy[:1])
    assert e.value.status == -5                  # VX_ERR_NOT_LOADED


def test_error_paths(seed):
    with vg.Context(0) as c:                     # no seed yet
        with pytest.raises(vg.VxError) as e:
            c.generate(vg.region_xy(1, 1, 1, 1))
        assert e.value.status == -4              # VX_ERR_NO_SEED
        with pytest.raises(vg.VxError) as e:
            c.mesh(vg.region_xy(1, 1, 1, 1))
        assert e.value.status == -5              # VX_ERR_NOT_LOADED
        c.set_seed(seed)
        assert c.mesh(np.zeros((0, 2), np.uint32))["n_faces"] == 0   # empty request
        c.generate(vg.region_xy(5, 5, 2, 2), read=False)
        bad = np.zeros(1, vg.EDIT_DTYPE)
        bad[0] = (5 * 16, 5 * 16, 128, 1)        # z out of range
        with pytest.raises(vg.VxError) as e:
            c.apply_edits(bad)
        assert e.value.status == -3
        bad[0] = (5 * 16, 5 * 16, 10, 27)        # voxel id out of range
        with pytest.raises(vg.VxError):
            c.apply_edits(bad)
        bad[0] = (900 * 16, 5 * 16, 10, 1)       # area not resident
        with pytest.raises(vg.VxError) as e:
            c.apply_edits(bad)
        assert e.value.status == -5


def test_regeneration_and_batch_split_are_deterministic(ctx, seed):
    """Same areas generated as one batch, as two batches in another order, and twice: identical."""
    xy = vg.region_xy(62480, 62520, 6, 4)
    a, ah = ctx.generate(xy)
    perm = np.random.default_rng(0).permutation(len(xy))
    with vg.Context(0, seed=seed) as c2:
        b1, bh1 = c2.generate(xy[perm[:10]])
        b2, bh2 = c2.generate(xy[perm[10:]])
    b = np.concatenate([b1, b2])[np.argsort(perm)]
    bh = np.concatenate([bh1, bh2])[np.argsort(perm)]
    assert np.array_equal(a, b) and np.array_equal(ah, bh)
    c, ch = ctx.generate(xy)
    assert np.array_equal(a, c) and np.array_equal(ah, ch)


def test_heightmap_image(ctx, seed):
    """HeightMap::generate_height_map pixel loop (height_map.rs:41-85) incl. the 31-area window,
    stale pixels and non-resident areas reading as height 127."""
    from voxel_game_b200 import world as w

    cam = (62500, 62500)
    with vg.Context(0, seed=seed) as c:
        zone = w.render_zone(cam, 17)                              # wider than the 31-area window
        loaded = vg.region_xy(cam[0] - 10, cam[1] - 10, 21, 21)    # only part of it is resident
        _, mh = c.generate(loaded)
        img = c.heightmap(zone, cam[0], cam[1])
        lookup = {(int(a), int(b)): mh[i] for i, (a, b) in enumerate(loaded)}
        heights = np.stack([lookup.get((int(a), int(b)), np.full(256, 127, np.uint8)) for a, b in zone])
        want = oracle.height_map(zone, heights, cam[0], cam[1])
        assert np.array_equal(img, want)
        # second frame, camera moved by one area, fewer areas listed: stale pixels persist
        img2 = c.heightmap(loaded[:50], cam[0] + 1, cam[1])
        want2 = oracle.height_map(loaded[:50], mh[:50], cam[0] + 1, cam[1], pixels=want.copy())
        assert np.array_equal(img2, want2)


# ------------------------------------------------------------------------------ edits (config 4, reduced)
def _edit_storm(rng_state, n_edits, ax0, ay0, n_side, world):
    """SplitMix64 stream; break if occupied, else place one of six plain blocks."""
    palette = [oracle.V[k] for k in ("Cobblestone", "Brick", "Boards", "Stone", "Lamp", "Glass")]
    L = oracle.lib()
    state = rng_state

    def draw():
        nonlocal state
        out = L.vgo_split_mix64(state)
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        return out

    edits = np.zeros(n_edits, vg.EDIT_DTYPE)
    span = n_side * 16
    for i in range(n_edits):
        gx = ax0 * 16 + draw() % span
        gy = ay0 * 16 + draw() % span
        gz = 1 + draw() % 126
        cur = world.voxels[world.slot(gx // 16, gy // 16)][(gx % 16) + 16 * (gy % 16) + 256 * gz]
        v = 0 if cur != 0 else palette[draw() % 6]
        edits[i] = (gx, gy, gz, v)
        world.apply_edit(int(gx), int(gy), int(gz), int(v))   # sequential reference semantics
    return edits


def test_edit_storm_batched_equals_sequential(seed):
    """Edits applied in GPU batches (last writer wins inside a batch) + relight of edited columns +
    remesh of dirty areas == the oracle applying World::set + Renderer::update_location one by
    one (world.rs:184-191, renderer.rs:239-286): voxels, heights, face sets, lamp sets."""
    n, ax0, ay0 = 6, 62496, 62496
    world = oracle.World.generate(seed, ax0 - 1, ay0 - 1, n + 2, n + 2)
    inner = vg.region_xy(ax0, ay0, n, n)
    for ax, ay in inner:
        world.mesh_area(int(ax), int(ay))
    with vg.Context(0, seed=seed) as c:
        c.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
        c.mesh(inner)
        edits = _edit_storm(0x5EEDED17, 3000, ax0, ay0, n, world)
        # duplicates inside one batch: the last writer must win
        dup = np.repeat(edits[1499:1500], 10)
        dup[5]["voxel"] = oracle.V["Glass"]
        dup[9]["voxel"] = oracle.V["Brick"]
        for e in dup:
            world.apply_edit(int(e["gx"]), int(e["gy"]), int(e["gz"]), int(e["voxel"]))
        edits = np.concatenate([edits, dup])
        dirty_all = set()
        for lo in range(0, len(edits), 1000):
            d = c.apply_edits(edits[lo:lo + 1000])
            dirty_all |= {(int(a), int(b)) for a, b in d}
        # every edited area is reported dirty, plus edge neighbours for border edits
        touched = {(int(e["gx"]) // 16, int(e["gy"]) // 16) for e in edits}
        assert touched <= dirty_all
        vox, mh = c.read_areas(inner)
        for i, (ax, ay) in enumerate(inner):
            s = world.slot(int(ax), int(ay))
            assert np.array_equal(vox[i], world.voxels[s]), f"voxels of {ax},{ay}"
            assert np.array_equal(mh[i], world.max_height[s]), f"heights of {ax},{ay}"
        dirty_inner = np.array(sorted(d for d in dirty_all if ax0 <= d[0] < ax0 + n and ay0 <= d[1] < ay0 + n), np.uint32)
        info = c.mesh(dirty_inner)
        rec_first, face_first, recs, verts, idx = c.mesh_read(info)
        lamps_gpu, lamps_cpu = set(), set()
        for i, (ax, ay) in enumerate(dirty_inner):
            s = world.slot(int(ax), int(ay))
            r = recs[rec_first[i]:rec_first[i + 1]]
            got = np.zeros(32768, np.uint8)
            li = (r["gx"] % 16) + 16 * (r["gy"] % 16) + 256 * (r["packed"] & 0xFF)
            got[li] = (r["packed"] >> 16) & 0xFF
            assert np.array_equal(got, world.face_mask[s]), f"face set of {ax},{ay}"
            orecs, overts, oidx = world.emit_area(int(ax), int(ay), first_face_base=int(face_first[i]))
            assert r.tobytes() == orecs.tobytes()
            assert verts[4 * face_first[i]:4 * face_first[i + 1]].tobytes() == overts.tobytes()
            lamps_gpu |= {(int(a["gx"]), int(a["gy"]), int(a["packed"]) & 0xFF) for a in r if (a["packed"] >> 8) & 0xFF == 11}
            fm = world.face_mask[s]
            for j in np.flatnonzero((world.voxels[s] == 11) & (fm != 0)):
                lamps_cpu.add((int(ax) * 16 + int(j) % 16, int(ay) * 16 + (int(j) // 16) % 16, int(j) // 256))
        assert lamps_gpu == lamps_cpu and len(lamps_cpu) > 0     # RenderArea.lights (renderer.rs:60-67)


# ------------------------------------------------------------------------------ full size (config 3)
def test_full_size_128x128_properties(seed):
    """BASELINE.json configs[2] at full size through size-independent properties: per-area digests
    of all 16 384 areas against the oracle (all host threads), column heights idempotent under the
    stand-alone operator, face counts consistent between records / offsets / totals, index pattern,
    every vertex within half a voxel of its voxel centre, and the mesh is bit-reproducible."""
    ax0, ay0, nx, ny = vg.spawn_region(128)
    ring = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2)
    inner = vg.region_xy(ax0, ay0, nx, ny)
    with vg.Context(0, seed=seed) as c:
        vox, mh = c.generate(ring)
        ovox, omh, _, _ = oracle.generate_areas(seed, ring)
        bad = np.flatnonzero((vox != ovox).any(axis=1))
        assert bad.size == 0, f"{bad.size} of {len(ring)} areas differ, first {ring[bad[0]]}"
        assert np.array_equal(mh, omh)
        sample = np.random.default_rng(1).choice(len(ring), 512, replace=False)
        assert np.array_equal(c.column_heights(vox[sample]), mh[sample])      # idempotence
        info = c.mesh(inner)
        rec_first, face_first, recs, verts, idx = c.mesh_read(info)
        info2 = c.mesh(inner)
        again = c.mesh_read(info2)
        assert info == info2 and all(a.tobytes() == b.tobytes() for a, b in zip((rec_first, face_first, recs, verts, idx), again))
    nf = (recs["packed"] >> 24).astype(np.int64)
    mask = ((recs["packed"] >> 16) & 0xFF).astype(np.uint8)
    assert int(nf.sum()) == info["n_faces"] == int(face_first[-1]) and int(rec_first[-1]) == info["n_recs"]
    assert np.array_equal(nf, np.unpackbits(mask[:, None], axis=1).sum(axis=1))
    assert np.array_equal(recs["first_face"].astype(np.int64), np.concatenate([[0], np.cumsum(nf)[:-1]]))
    k = np.arange(info["n_faces"]) - np.repeat(recs["first_face"].astype(np.int64), nf)     # ordinal in voxel
    want_idx = (np.array([0, 1, 2, 0, 2, 3])[None, :] + 4 * k[:, None]).astype(np.uint16).ravel()
    assert np.array_equal(idx, want_idx)
    centre = np.stack([recs["gx"].astype(np.int64) - 1_000_000, recs["gy"].astype(np.int64) - 1_000_000,
                       (recs["packed"] & 0xFF).astype(np.int64)], axis=1).astype(np.float32)
    per_vertex_centre = np.repeat(centre, 4 * nf, axis=0)
    assert np.abs(verts["position"] - per_vertex_centre).max() == 0.5
    assert (verts["color"] == 255).all() and (np.abs(verts["normal"]).sum(axis=1) == 1).all()
    # the oracle's face count for a 1 024-area sample of the region (serial CPU meshing is slow)
    world = oracle.World(ax0 - 1, ay0 - 1, nx + 2, ny + 2, ovox, omh)
    pick = np.random.default_rng(2).choice(len(inner), 1024, replace=False)
    for i in pick:
        assert world.mesh_area(int(inner[i][0]), int(inner[i][1])) == int(face_first[i + 1] - face_first[i])


def test_cpp_host_mirror(tmp_path):
    """The reference's own generator/world/renderer unit tests, restated in C++ against the host
    mirror (voxel_game_b200/host/voxel_game.hpp) that sits on the C ABI: tests/cpp/test_host.cpp."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_host")
    libdir = os.path.join(root, "voxel_game_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-o", exe, os.path.join(root, "tests", "cpp", "test_host.cpp"),
                    "-L" + libdir, "-lvoxel_cuda", "-Wl,-rpath," + libdir], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def test_light_extension_equals_flood_fill(seed):
    """L-ext (EXTENSION, no reference counterpart -- SURVEY.md F3): the CUDA relaxation reaches the
    same fixed point as the oracle's sequential BFS flood fill (oracle/light_ext.c) over a 5x5 grid
    with caves, tree canopies, and lamps placed underground and next to area borders."""
    n, ax0, ay0 = 5, 62470, 62500
    world = oracle.World.generate(seed, ax0, ay0, n, n)
    xy = vg.region_xy(ax0, ay0, n, n)
    lamps = [(ax0 * 16 + 40, ay0 * 16 + 40, 100), (ax0 * 16 + 15, ay0 * 16 + 33, 90), (ax0 * 16 + 16, ay0 * 16 + 70, 110),
             (ax0 * 16 + 79, ay0 * 16 + 79, 60), (ax0 * 16 + 1, ay0 * 16 + 1, 120)]
    edits = np.zeros(len(lamps) * 2, vg.EDIT_DTYPE)
    for i, (gx, gy, gz) in enumerate(lamps):
        edits[2 * i] = (gx, gy, gz, oracle.V["Lamp"])
        edits[2 * i + 1] = (gx, gy, gz - 1, 0)            # a pocket of air above the lamp
    for e in edits:
        world.apply_edit(int(e["gx"]), int(e["gy"]), int(e["gz"]), int(e["voxel"]))
    want = world.light_flood()
    with vg.Context(0, seed=seed) as c:
        c.generate(xy, read=False)
        c.apply_edits(edits)
        got, iters = c.light(xy)
    assert 1 <= iters < 64
    assert np.array_equal(got, want), f"{(got != want).sum()} light values differ"
    assert (want == 15).any() and ((want > 0) & (want < 15)).sum() > 500    # real gradients exist


def test_contexts_with_different_seeds_do_not_interfere():
    """Two contexts (two worlds) in one process, calls interleaved: each keeps its own noise tables
    (they travel as a kernel parameter, not as a module-wide __constant__ symbol)."""
    s1, s2 = oracle.hash_world_name("test1"), oracle.hash_world_name("test2")
    xy = vg.region_xy(123, 456, 2, 2)
    with vg.Context(0, seed=s1) as c1, vg.Context(0, seed=s2) as c2:
        a1, _ = c1.generate(xy)
        a2, _ = c2.generate(xy)
        b1, _ = c1.generate(xy)
    o1, _, _, _ = oracle.generate_areas(s1, xy)
    o2, _, _, _ = oracle.generate_areas(s2, xy)
    assert np.array_equal(a1, o1) and np.array_equal(b1, o1) and np.array_equal(a2, o2)
    assert not np.array_equal(o1, o2)


# ------------------------------------------------------------------------------ full-size configs 4 and 5
def test_config4_edit_storm_full_size(seed):
    """BASELINE.json configs[3]: 100 000 random place/break events over a loaded 64x64 region
    (SURVEY.md 8d C4 recipe), GPU in batches of 1 000 + remesh of dirty areas, against the oracle
    applying World::set + Renderer::update_location one event at a time."""
    n, (ax0, ay0) = 64, (62500 - 32, 62500 - 32)
    world = oracle.World.generate(seed, ax0 - 1, ay0 - 1, n + 2, n + 2)
    inner = vg.region_xy(ax0, ay0, n, n)
    world.mesh_areas(inner, threads=0)
    edits = _edit_storm(0x5EEDED17, 100_000, ax0, ay0, n, world)
    with vg.Context(0, seed=seed) as c:
        c.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
        c.mesh(inner)
        dirty_all = set()
        for lo in range(0, len(edits), 1000):
            dirty_all |= {(int(a), int(b)) for a, b in c.apply_edits(edits[lo:lo + 1000])}
        vox, mh = c.read_areas(inner)
        slots = [world.slot(int(ax), int(ay)) for ax, ay in inner]
        assert np.array_equal(vox, world.voxels[slots]) and np.array_equal(mh, world.max_height[slots])
        dirty_inner = np.array(sorted(d for d in dirty_all if ax0 <= d[0] < ax0 + n and ay0 <= d[1] < ay0 + n), np.uint32)
        assert len(dirty_inner) > 3000          # nearly every area of the region was touched
        info = c.mesh(dirty_inner)
        rec_first, face_first, recs, verts, idx = c.mesh_read(info)
    got = np.zeros((len(dirty_inner), 32768), np.uint8)
    area_of_rec = np.repeat(np.arange(len(dirty_inner)), np.diff(rec_first.astype(np.int64)))
    li = (recs["gx"] % 16) + 16 * (recs["gy"] % 16) + 256 * (recs["packed"] & 0xFF)
    got[area_of_rec, li] = (recs["packed"] >> 16) & 0xFF
    want = world.face_mask[[world.slot(int(ax), int(ay)) for ax, ay in dirty_inner]]
    bad = np.flatnonzero((got != want).any(axis=1))
    assert bad.size == 0, f"face sets differ in {bad.size} areas, first {dirty_inner[bad[0]]}"
    # vertices of a sample of the dirty areas, byte for byte
    for i in np.random.default_rng(3).choice(len(dirty_inner), 64, replace=False):
        ax, ay = int(dirty_inner[i][0]), int(dirty_inner[i][1])
        _, overts, oidx = world.emit_area(ax, ay)
        f0, f1 = int(face_first[i]), int(face_first[i + 1])
        assert verts[4 * f0:4 * f1].tobytes() == overts.tobytes() and np.array_equal(idx[6 * f0:6 * f1], oidx)


def test_config5_512x512_pregeneration_sample(seed):
    """BASELINE.json configs[4] at one GPU's share of the 8-GPU sweep and beyond: the whole 512x512
    region (262 144 areas, 8 GiB of block IDs) is generated and meshed on one B200; a fixed
    1 024-area sample and its meshes are compared with the oracle."""
    n = 512
    ax0, ay0, _, _ = vg.spawn_region(n)
    xy = vg.region_xy(ax0, ay0, n, n)
    rng = np.random.default_rng(11)
    with vg.Context(0, seed=seed) as c:
        c.set_timing(True)
        c.generate(xy, read=False)
        t_gen = c.timings()
        info = c.mesh(xy)                       # the ring of 2 048 edge neighbours is generated on demand
        t_mesh = c.timings()
        assert c.resident_count() == n * n + 4 * n
        assert info["n_faces"] > 100_000_000
        pick = np.sort(rng.choice(len(xy), 1024, replace=False))
        vox, mh = c.read_areas(xy[pick])
        ovox, omh, _, _ = oracle.generate_areas(seed, xy[pick])
        assert np.array_equal(vox, ovox) and np.array_equal(mh, omh)
        # meshes of 64 of them: re-mesh just those areas (neighbours are resident) and compare
        sub = xy[pick[::16]]
        info2 = c.mesh(sub)
        rec_first, face_first, recs, verts, idx = c.mesh_read(info2)
    for i, (ax, ay) in enumerate(sub):
        w = oracle.World.generate(seed, int(ax) - 1, int(ay) - 1, 3, 3, threads=0)
        nf = w.mesh_area(int(ax), int(ay))
        orecs, overts, oidx = w.emit_area(int(ax), int(ay), first_face_base=int(face_first[i]))
        assert int(face_first[i + 1] - face_first[i]) == nf
        assert recs[rec_first[i]:rec_first[i + 1]].tobytes() == orecs.tobytes()
        assert verts[4 * face_first[i]:4 * face_first[i + 1]].tobytes() == overts.tobytes()
    gen_ms = t_gen["gen_columns_ms"] + t_gen["gen_caves_ms"] + t_gen["gen_finish_ms"]
    print(f"config5 on one B200: gen {gen_ms:.1f} ms, mesh {t_mesh['mesh_count_ms'] + t_mesh['mesh_emit_ms']:.1f} ms "
          f"for {n * n} areas, {info['n_faces']} faces")


def test_concurrent_callers_share_one_context(seed):
    """The reference enters generation from rayon workers (world_persistence.rs:90-93): eight host
    threads hammer one context with disjoint batches; every result equals the oracle's."""
    import threading

    xy = vg.region_xy(62400, 62600, 16, 8)
    parts = np.array_split(np.arange(len(xy)), 8)
    out, errs = {}, []
    with vg.Context(0, seed=seed) as c:
        def worker(k):
            try:
                for _ in range(3):
                    out[k] = c.generate(xy[parts[k]])
            except Exception as e:  # pragma: no cover
                errs.append(e)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert not errs, errs
        assert c.resident_count() == len(xy)
    ovox, omh, _, _ = oracle.generate_areas(seed, xy)
    for k in range(8):
        assert np.array_equal(out[k][0], ovox[parts[k]]) and np.array_equal(out[k][1], omh[parts[k]])


def test_residency_churn_keeps_map_and_pool_consistent(seed):
    """World::retain_areas-style churn (world.rs:196-240): load, unload, load again in random
    batches, with pool growth in between.  After every round every resident area reads back
    bit-exact, unloaded ones are gone, and meshing a resident patch still matches the oracle
    (exercises the host/device area hash map incl. deletions, slot reuse, scattered slot runs)."""
    rng = np.random.default_rng(42)
    universe = vg.region_xy(62440, 62470, 24, 24)
    ovox, omh, _, _ = oracle.generate_areas(seed, universe)
    key = {(int(a), int(b)): i for i, (a, b) in enumerate(universe)}
    resident = set()
    with vg.Context(0, seed=seed) as c:
        for rnd in range(8):
            add = [tuple(map(int, universe[i])) for i in rng.choice(len(universe), 150, replace=False)]
            c.generate(np.array(add, np.uint32), read=False)
            resident |= set(add)
            drop = [a for a in resident if rng.random() < 0.4]
            c.unload(np.array(drop, np.uint32).reshape(-1, 2))
            resident -= set(drop)
            assert c.resident_count() == len(resident)
            lst = np.array(sorted(resident), np.uint32)
            perm = rng.permutation(len(lst))
            v, h = c.read_areas(lst[perm])
            want = [key[tuple(map(int, a))] for a in lst[perm]]
            assert np.array_equal(v, ovox[want]) and np.array_equal(h, omh[want]), f"round {rnd}"
            for a in drop[:3]:
                with pytest.raises(vg.VxError):
                    c.read_areas(np.array([a], np.uint32))
        # mesh a 6x6 patch (vx_mesh generates whatever neighbours the churn left missing)
        patch = vg.region_xy(62445, 62475, 6, 6)
        c.generate(patch, read=False)
        world = oracle.World.generate(seed, 62444, 62474, 8, 8)
        _compare_mesh(c, world, patch)


def test_config1_spawn_area_at_shipped_render_distance(seed):
    """BASELINE.json configs[0]: what VoxelEngine::load_world does at the default render distance 8
    (voxel_engine.rs:114-135): world.load_all_blocking over the 19x19 load zone, then
    renderer.load_all_blocking over the 15x15 render zone, zones in the reference's own order
    (active_zone.rs:10-56).  Everything compared with the oracle."""
    from voxel_game_b200 import world as w

    load = w.load_zone_on_world_load(w.SPAWN_AREA, 8)
    render = w.render_zone_on_world_load(w.SPAWN_AREA, 8)
    assert len(load) == 361 and len(render) == 225
    with vg.Context(0, seed=seed) as c:
        vox, mh = c.generate(load)
        ovox, omh, _, _ = oracle.generate_areas(seed, load)
        assert np.array_equal(vox, ovox) and np.array_equal(mh, omh)
        # oracle world as a dense grid (the zone lists are x-outer, the grid is y-outer)
        grid = vg.region_xy(int(load[:, 0].min()), int(load[:, 1].min()), 19, 19)
        gv, gh, _, _ = oracle.generate_areas(seed, grid)
        world = oracle.World(int(grid[0][0]), int(grid[0][1]), 19, 19, gv, gh)
        faces = _compare_mesh(c, world, render)
        assert c.resident_count() == 361        # the render zone's neighbours were all loaded already
        assert faces > 50_000


@pytest.mark.gpu
def test_generate_without_outputs_is_stream_ordered(seed):
    """vx_generate with no host buffers returns once the kernels are enqueued; whatever reads the
    areas next sees them complete, and the timings accumulate until they are read."""
    xy = vg.region_xy(SPAWN - 3, SPAWN - 3, 6, 6)
    with vg.Context(0, seed=seed) as c:
        c.set_timing(True)
        c.generate(xy, read=False)
        c.generate(xy, read=False)           # a second call before anything waited for the first
        vox, mh = c.read_areas(xy)
        t = c.timings()
        assert t["gen_launches"] == 6 and t["gen_columns_ms"] > 0 and t["gen_caves_ms"] >= 0
        assert t["simplex2_evals"] > 0
        again = c.timings()                  # reading clears the accumulators
        assert again["gen_launches"] == 0 and again["gen_columns_ms"] == 0
        ovox, omh, _, _ = oracle.generate_areas(seed, xy)
        assert np.array_equal(vox, ovox) and np.array_equal(mh, omh)
        # mesh returns when the totals are known, emit may still be running: the reads line up behind it
        inner = vg.region_xy(SPAWN - 2, SPAWN - 2, 4, 4)
        info = c.mesh(inner)
        first = c.mesh_read(info)
        c.generate(xy, read=False)           # regenerating right behind a mesh in flight changes nothing
        info2 = c.mesh(inner)
        second = c.mesh_read(info2)
        assert info == info2
        for a, b in zip(first, second):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_strips_on_cuda_equal_unsharded(seed, shards):
    """SURVEY 8(e) on the device: `shards` contexts (one per rank in bench.py --gpus N), each
    generating its strip of rows of ONE region plus the regenerated ring and meshing its strip,
    give -- concatenated in rank order -- the block IDs, heights, records, vertices and indices of
    one context doing the whole region, byte for byte (first_face rebased by the strips before)."""
    n = 32 if shards < 8 else 40  # 40 rows over 8 ranks: 5 each; 32 over 3: ragged 11/11/10
    ax0, ay0 = SPAWN - n // 2, SPAWN - n // 2
    whole_xy = vg.region_xy(ax0, ay0, n, n)
    with vg.Context(0, seed=seed) as one:
        one.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
        vox, mh = one.read_areas(whole_xy)
        info = one.mesh(whole_xy)
        rf, ff, recs, verts, idx = one.mesh_read(info)
    rec_base = face_base = row = 0
    for rank in range(shards):
        lo, hi = vg.shard_rows(n, shards, rank)
        assert lo == row
        row = hi
        with vg.Context(0, seed=seed) as c:
            c.generate(vg.region_xy(ax0 - 1, ay0 + lo - 1, n + 2, hi - lo + 2), read=False)
            sxy = vg.region_xy(ax0, ay0 + lo, n, hi - lo)
            svox, smh = c.read_areas(sxy)
            sinfo = c.mesh(sxy)
            srf, sff, srecs, sverts, sidx = c.mesh_read(sinfo)
        a0, a1 = lo * n, hi * n  # region_xy order: index = ix + nx * iy, so a strip of rows is contiguous
        assert np.array_equal(svox, vox[a0:a1]) and np.array_equal(smh, mh[a0:a1])
        assert np.array_equal(srf.astype(np.int64) + rec_base, rf[a0:a1 + 1])
        assert np.array_equal(sff.astype(np.int64) + face_base, ff[a0:a1 + 1])
        srecs = srecs.copy()
        srecs["first_face"] += np.uint32(face_base)
        assert srecs.tobytes() == recs[rec_base:rec_base + sinfo["n_recs"]].tobytes()
        assert sverts.tobytes() == verts[4 * face_base:4 * (face_base + sinfo["n_faces"])].tobytes()
        assert np.array_equal(sidx, idx[6 * face_base:6 * (face_base + sinfo["n_faces"])])
        rec_base += sinfo["n_recs"]
        face_base += sinfo["n_faces"]
    assert row == n and rec_base == info["n_recs"] and face_base == info["n_faces"]


def test_query_and_read_pairs_are_atomic_across_threads(seed):
    """vx_mesh -> vx_mesh_read (and the other query + read pairs) fetch what the LAST
    query on the context left.  Four threads share one context and mesh batches of different sizes;
    with the pair under vx_lock / vx_unlock (Context.mesh_and_read, the Rust / C++ `Transaction`)
    every thread gets exactly its own batch, equal to what the batch gives alone."""
    import threading

    n = 12
    ax0, ay0 = SPAWN - 6, SPAWN - 6
    rows = [(0, 1), (1, 4), (4, 6), (6, 12)]  # batches of 12, 36, 24, 72 areas
    lists = [vg.region_xy(ax0, ay0 + lo, n, hi - lo) for lo, hi in rows]
    with vg.Context(0, seed=seed) as c:
        c.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
        alone = [c.mesh_and_read(xy) for xy in lists]
        got, errs = {}, []

        def worker(k):
            try:
                for _ in range(25):
                    info, arrays = c.mesh_and_read(lists[k])
                    files, offs = c.encode_areas(lists[k])
                    got[k] = (info, arrays, files.copy(), offs)
            except Exception as e:  # pragma: no cover
                errs.append(e)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert not errs, errs
        for k in range(4):
            info, arrays, files, offs = got[k]
            assert info == alone[k][0]
            for a, b in zip(arrays, alone[k][1]):
                assert a.tobytes() == b.tobytes()
            assert len(offs) == len(lists[k]) + 1 and int(offs[-1]) == len(files)
            vox, _ = c.read_areas(lists[k])
            for i in range(len(lists[k])):
                dec = oracle.area_file_decode(bytes(files[int(offs[i]):int(offs[i + 1])]))
                assert dec is not None and np.array_equal(dec, vox[i])


def test_cave_columns_cost_hint_equals_is_cave_zone(ctx, seed):
    """vx_cave_columns = per area the columns for which CaveGenerator::is_cave_zone
    (cave_generator.rs:36-41) holds and their voxels in should_be_cave's range, from the seed alone;
    the oracle's per-column samples agree, and so do the work counters of a generate over the list."""
    xy = vg.region_xy(SPAWN - 8, SPAWN - 8, 16, 16)
    cols, vox = ctx.cave_columns(xy)
    gen = oracle.Generator(seed)
    want_c, want_v = np.zeros(len(xy), np.uint32), np.zeros(len(xy), np.uint32)
    for i, (ax, ay) in enumerate(xy):
        for x in range(16):
            for y in range(16):
                h, cave, _, _ = gen.sample_column(int(ax), int(ay), x, y)
                if cave:
                    want_c[i] += 1
                    want_v[i] += max(0, min(h, 114) - 25 + 1)
    assert np.array_equal(cols, want_c) and np.array_equal(vox, want_v)
    assert cols.max() == 256 and (cols == 0).sum() > 50  # whole cave areas and cave-free ones
    ctx.set_timing(True)
    ctx.generate(xy, read=False)
    t = ctx.timings()
    assert t["cave_columns"] == int(cols.sum()) and t["cave_voxels"] == int(vox.sum())
    ctx.set_timing(False)


def test_pool_grows_in_place_by_mapping_memory(seed):
    """The resident-area pool is a set of virtual-memory ranges on B200: growth maps more physical
    memory behind them -- resident areas keep their bytes (nothing is copied, nothing moves) and the
    committed size follows the capacity instead of doubling past it."""
    with vg.Context(0, seed=seed) as c:
        info0 = c.pool_info()
        assert info0["virtual_memory"], "the CUDA virtual-memory API is available on every B200 driver"
        assert info0["capacity_areas"] == 0 and info0["committed_bytes"] == 0
        first = vg.region_xy(100, 200, 20, 20)
        vox0, mh0 = c.generate(first)
        c.set_mesh_retention(True)
        c.mesh(vg.region_xy(101, 201, 18, 18))
        masks0, _ = c.read_face_masks(first)
        caps = [c.pool_info()["capacity_areas"]]
        for k in range(1, 6):  # each batch forces the pool to grow
            c.generate(vg.region_xy(1000 * k, 300, 40 * k, 40), read=False)
            caps.append(c.pool_info()["capacity_areas"])
        assert caps == sorted(caps) and caps[-1] > 4 * caps[0]
        vox1, mh1 = c.read_areas(first)
        masks1, _ = c.read_face_masks(first)
        assert np.array_equal(vox0, vox1) and np.array_equal(mh0, mh1) and np.array_equal(masks0, masks1)
        info = c.pool_info()
        per_area = 32768 + 256 + 8192 + 8 + 32768 + 1
        assert info["committed_bytes"] < info["capacity_areas"] * per_area + 6 * (64 << 20)
        c.set_mesh_retention(False)  # the face masks' physical memory goes back to the driver
        assert c.pool_info()["committed_bytes"] < info["committed_bytes"]
