"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md 8c).  Each test names the reference test it restates (file:line under
/root/reference/src).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import T, V


# --------------------------------------------------------------------------- face predicate
# graphics/mesh_generator.rs:633-744
FACE_ROWS = [
    # test_should_generate_face_solid
    ("Brick", "Brick", False), ("Brick", "Stone", False), ("Stone", "Brick", False),
    # test_should_generate_face_air
    ("Brick", "None", True), ("None", "Stone", False),
    # test_should_generate_face_transparent
    ("Brick", "Glass", True), ("Glass", "Stone", False), ("Ice", "Glass", True),
    ("Ice", "WaterSource", True), ("Glass", "WaterDown", True), ("Glass", "Glass", False),
    ("Ice", "Ice", False), ("WaterSource", "WaterSource", False),
    # test_should_generate_face_water
    ("WaterSource", "WaterDown", True), ("WaterDown", "WaterSource", True),
    ("WaterSource", "Brick", False), ("Brick", "WaterSource", True),
]
TOP_ROWS = [
    # test_should_generate_top_face_solid / _air / _water
    ("Boards", "Brick", False), ("Boards", "Boards", False), ("Boards", "None", True),
    ("Water1", "Brick", True),
]


@pytest.mark.parametrize("cur,nbr,expect", FACE_ROWS)
def test_should_generate_face_rows(cur, nbr, expect):
    assert oracle.should_generate_face(V[cur], V[nbr]) is expect


@pytest.mark.parametrize("cur,nbr,expect", TOP_ROWS)
def test_should_generate_top_face_rows(cur, nbr, expect):
    assert oracle.should_generate_top_face(V[cur], V[nbr]) is expect


def test_top_face_water_differs_from_plain_face():
    # mesh_generator.rs:733-743: Water1 over Brick has a top face but no plain face
    assert not oracle.should_generate_face(V["Water1"], V["Brick"])


def test_face_predicate_equals_algebraic_form():
    """cur != nbr && transparent(nbr) (the form the CUDA bit planes use) == the three-clause
    source form, for all 27 x 27 pairs."""
    transparent = {V[n] for n in ["None", "Glass", "WaterSource", "WaterDown", "Water1", "Water2",
                                  "Water3", "Water4", "Ice"]}  # voxel.rs:39-49
    for cur in range(27):
        for nbr in range(27):
            assert oracle.should_generate_face(cur, nbr) == (cur != nbr and nbr in transparent)


# --------------------------------------------------------------------------- column heights
def test_calculate_max_height():  # model/area.rs:233-249
    a = oracle.Area()
    a.set(0, 0, 10, V["Brick"]); assert a.sample_height(0, 0) == 10
    a.set(0, 0, 5, V["Brick"]); assert a.sample_height(0, 0) == 5
    a.set(0, 0, 20, V["Brick"]); assert a.sample_height(0, 0) == 5
    a.set(1, 0, 1, V["Brick"])
    assert a.sample_height(0, 0) == 5 and a.sample_height(1, 0) == 1


def test_calculate_max_height_transparent():  # model/area.rs:251-266
    a = oracle.Area()
    a.set(0, 0, 10, V["Brick"]); assert a.sample_height(0, 0) == 10
    a.set(0, 0, 5, V["Glass"]); assert a.sample_height(0, 0) == 10
    a.set(0, 0, 8, V["Brick"]); assert a.sample_height(0, 0) == 8
    a.set(1, 0, 1, V["Glass"]); assert a.sample_height(1, 0) == 127


def test_calculate_max_height_replace_existing():  # model/area.rs:268-286
    a = oracle.Area()
    a.set(0, 0, 10, V["Brick"]); assert a.sample_height(0, 0) == 10
    a.set(0, 0, 10, V["Brick"]); assert a.sample_height(0, 0) == 10
    a.set(0, 0, 10, V["Glass"]); assert a.sample_height(0, 0) == 127
    a.set(0, 0, 5, V["Brick"]); assert a.sample_height(0, 0) == 5
    a.set(0, 0, 5, V["None"]); assert a.sample_height(0, 0) == 127


def test_set_without_calculating_max_height():  # model/area.rs:288-304
    a = oracle.Area()
    a.set_without_updating_max_height(0, 0, 10, V["Brick"]); assert a.sample_height(0, 0) == 127
    a.set_without_updating_max_height(0, 0, 5, V["Brick"]); assert a.sample_height(0, 0) == 127
    a.update_all_column_heights()
    a.set(0, 0, 20, V["Brick"]); assert a.sample_height(0, 0) == 5
    a.set(1, 0, 1, V["Brick"])
    assert a.sample_height(0, 0) == 5 and a.sample_height(1, 0) == 1


# --------------------------------------------------------------------------- trees
def _u8p(a):
    import ctypes as C

    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def test_can_generate_tree_true():  # area_generation/trees.rs:247-256
    a = oracle.Area()
    assert oracle.lib().vgo_can_generate_tree(_u8p(a.voxels), 5, 5, 10, T["Tall"]) == 1


def test_can_generate_tree_false():  # area_generation/trees.rs:258-273
    a = oracle.Area()
    a.set(5, 5, 8, V["Brick"])
    assert oracle.lib().vgo_can_generate_tree(_u8p(a.voxels), 5, 5, 10, T["Tall"]) == 0
    assert oracle.lib().vgo_can_generate_tree(_u8p(a.voxels), 5, 5, 1, T["Tall"]) == 0  # out of bounds


# templates of trees.rs:15-75, restated independently of oracle/gen.c
TEMPLATES = {
    "Short": [((0, 0, -1), "Wood"), ((0, 0, -2), "Wood"), ((0, 0, -3), "Wood"), ((0, 0, -4), "Leaves"),
              ((0, -1, -3), "Leaves"), ((0, 1, -3), "Leaves"), ((1, 0, -3), "Leaves"), ((-1, 0, -3), "Leaves")],
    "Tall": [((0, 0, -1), "Wood"), ((0, 0, -2), "Wood"), ((0, 0, -3), "Wood"), ((0, 0, -4), "Wood"),
             ((0, 0, -5), "Leaves"), ((0, -1, -4), "Leaves"), ((0, 1, -4), "Leaves"), ((1, 0, -4), "Leaves"),
             ((-1, 0, -4), "Leaves")],
    "TallCactus": [((0, 0, -1), "Cactus"), ((0, 0, -2), "Cactus"), ((0, 0, -3), "Cactus")],
    "ShortCactus": [((0, 0, -1), "Cactus")],
    "DeadTree": [((0, 0, -1), "Wood"), ((0, 0, -2), "Wood")],
    "Bush": [((0, 0, -1), "Leaves")],
}


@pytest.mark.parametrize("tree", ["Short", "Tall", "TallCactus", "ShortCactus", "DeadTree", "HugeTree", "Bush"])
def test_generate_tree(tree):  # area_generation/trees.rs:275-299
    a = oracle.Area()
    x, y, z = 8, 8, 64
    oracle.lib().vgo_generate_tree(_u8p(a.voxels), x, y, z, T[tree])
    placed = np.flatnonzero(a.voxels)
    assert placed.size > 0
    if tree in TEMPLATES:
        want = {(x + dx) + 16 * (y + dy) + 256 * (z + dz): V[v] for (dx, dy, dz), v in TEMPLATES[tree]}
        assert {int(i): int(a.voxels[i]) for i in placed} == want
    else:  # HugeTree: 27 cells, 9 wood + 18 leaves, reaching 7 above the base
        assert placed.size == 27
        assert int((a.voxels == V["Wood"]).sum()) == 9 and int((a.voxels == V["Leaves"]).sum()) == 18
        assert a.get(x, y, z - 7) == V["Leaves"] and a.get(x + 2, y, z - 5) == V["Leaves"]


def test_tree_pick_distribution_and_base_rules():
    """should_generate_tree (trees.rs:142-196): ~1/60 of eligible columns, only types whose
    allowed base contains the voxel, never on Stone/Snow."""
    seed = oracle.hash_world_name("test")
    L = oracle.lib()
    allowed = {"Grass": {"Short", "Tall", "DeadTree", "HugeTree", "Bush"}, "Clay": {"Short", "Tall", "DeadTree"},
               "Dirt": {"DeadTree"}, "Sand": {"TallCactus", "ShortCactus", "DeadTree", "Bush"}}
    for base, ok in allowed.items():
        hits = 0
        for i in range(6000):
            t = L.vgo_should_generate_tree(V[base], seed, 100 + i % 7, 200 + i // 7, i % 16, (i // 16) % 16, 60)
            if t:
                hits += 1
                assert oracle.TREE_NAMES[t] in ok
        assert 50 <= hits <= 160  # 6000 / 60 = 100 expected
    for base in ("Stone", "Snow", "None", "WaterSource"):
        assert all(L.vgo_should_generate_tree(V[base], seed, 1, 2, x, 3, 60) == 0 for x in range(16))


def test_split_mix64_and_combine_seed_reference_values():
    """algorithms.rs:28-55: SplitMix64 is the published generator (first output for state 0 is
    0xE220A8397B1DCDAF); combine_seed is restated independently here."""
    L = oracle.lib()
    assert L.vgo_split_mix64(0) == 0xE220A8397B1DCDAF
    M = (1 << 64) - 1

    def rotl(v, r):
        return ((v << r) | (v >> (64 - r))) & M

    def combine(seed, ax, ay, lx, ly, lz):
        s = (seed * 0x517CC1B727220A95) & M
        for add, r in ((ax, 11), (ay, 13), (lx, 17), (ly, 19), (lz, 23)):
            s = rotl((s + add) & M, r)
        return (s * 0xFF51AFD7ED558CCD) & M

    for args in [(0, 0, 0, 0, 0, 0), (12345678901234567890, 62500, 62501, 3, 15, 77), (M, 1, 2, 3, 4, 5)]:
        assert L.vgo_combine_seed(*args) == combine(*args)


# --------------------------------------------------------------------------- generator invariants
@pytest.fixture(scope="module")
def area_123_456():
    g = oracle.Generator(oracle.hash_world_name("test"))
    return g.generate_area(123, 456)


def test_generate_area(area_123_456):  # area_generation/generator.rs:155-179
    vox, mh, _ = area_123_456
    transparent = np.isin(vox, [V[n] for n in ["None", "Glass", "WaterSource", "WaterDown", "Water1", "Water2",
                                               "Water3", "Water4", "Ice"]])
    cols = vox.reshape(128, 256)
    tr = transparent.reshape(128, 256)
    for c in range(256):
        opaque = np.flatnonzero(~tr[:, c])
        assert opaque.size > 0                              # max_height.is_some()
        assert (cols[:, c] == V["Stone"]).any()             # contains_stone
        assert int(opaque[0]) == int(mh[c])                 # == area.sample_height(x, y)


def test_generate_area_different_locations():  # generator.rs:181-186
    g = oracle.Generator(oracle.hash_world_name("test"))
    a, b = g.generate_area(123, 456), g.generate_area(999, 400)
    assert not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))


def test_generate_area_different_seeds():  # generator.rs:188-193
    a = oracle.Generator(oracle.hash_world_name("test1")).generate_area(123, 456)
    b = oracle.Generator(oracle.hash_world_name("test2")).generate_area(123, 456)
    assert not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))


def test_generate_same_area(area_123_456):  # generator.rs:195-200
    again = oracle.Generator(oracle.hash_world_name("test")).generate_area(123, 456)
    assert np.array_equal(area_123_456[0], again[0]) and np.array_equal(area_123_456[1], again[1])


def test_generate_area_heights_calculated_correctly():  # generator.rs:202-220
    g = oracle.Generator(oracle.hash_world_name("test"))
    for x in range(10):
        vox, mh, _ = g.generate_area(x, 123)
        assert np.array_equal(oracle.update_all_column_heights(vox), mh)


def test_terrain_height_range_and_batch_equals_single():
    """H = (..) as u32 + 32 lies in [32, 102] (terrain_type.rs:26-32); the OpenMP batch with tables
    rebuilt per area (generator.rs:51) equals the hoisted-table path."""
    seed = oracle.hash_world_name("test")
    g = oracle.Generator(seed)
    for (ax, ay, x, y) in [(62500, 62500, 0, 0), (62484, 62510, 7, 9), (1, 1, 15, 15), (123, 456, 3, 4)]:
        h, cave, lake, biome = g.sample_column(ax, ay, x, y)
        assert 32 <= h <= 102 and cave in (0, 1) and 0 <= lake <= 10 and biome in (0, 1, 2)
    xy = oracle.region_xy(62496, 62496, 4, 4)
    v1, h1, s1, _ = oracle.generate_areas(seed, xy, rebuild_tables_per_area=True)
    v2, h2, s2, _ = oracle.generate_areas(seed, xy, rebuild_tables_per_area=False, threads=1)
    assert np.array_equal(v1, v2) and np.array_equal(h1, h2) and s1 == s2
    single = g.generate_area(int(xy[5][0]), int(xy[5][1]))
    assert np.array_equal(single[0], v1[5]) and np.array_equal(single[1], h1[5])


# --------------------------------------------------------------------------- zones / height map
def test_zone_sizes():  # service/active_zone.rs:62-130
    from voxel_game_b200 import world as w

    rz = w.render_zone((10, 10), 2)
    assert len(rz) == 25 and (np.abs(rz.astype(int) - 10) <= 2).all()
    assert len(w.render_zone_on_world_load((10, 10), 9)) == 15 * 15            # minimum 7
    assert len(w.render_zone_on_world_load((10, 10), 16)) == (1 + 2 * 13) ** 2  # 16 - 3
    assert len(w.load_zone((10, 10), 2)) == (1 + 2 * 4) ** 2                    # + LOAD_EXTRA
    assert len(w.load_zone_on_world_load((10, 10), 4)) == (1 + 2 * 5) ** 2      # + INITIAL_LOAD_EXTRA
    # config 1 of BASELINE.json: default render distance 8 around the spawn area
    assert len(w.load_zone_on_world_load(w.SPAWN_AREA, 8)) == 361
    assert len(w.render_zone_on_world_load(w.SPAWN_AREA, 8)) == 225


def test_height_map_window_and_pixels():  # graphics/height_map.rs:41-101,110-126
    from voxel_game_b200 import world as w

    cam = (100, 100)
    zone = w.render_zone(cam, 20)
    mh = (np.arange(len(zone) * 256) % 251).astype(np.uint8).reshape(len(zone), 256)
    img = oracle.height_map(zone, mh, cam[0], cam[1])
    touched = (img != 255).any(axis=2)
    ys, xs = np.nonzero(touched | (img[..., 0] == 255) & False)
    # filter_distant_areas keeps |d| < 16 -> a 31-area (496 px) square starting at pixel 16
    cols = np.flatnonzero(touched.any(axis=0))
    rows = np.flatnonzero(touched.any(axis=1))
    assert cols.min() >= 16 and cols.max() <= 511 and rows.min() >= 16 and rows.max() <= 511
    # one pixel, by hand: area (103, 98), column (5, 7) -> image (256 + 3*16 + 5, 256 - 2*16 + 7)
    i = int(np.flatnonzero((zone[:, 0] == 103) & (zone[:, 1] == 98))[0])
    assert (img[256 - 32 + 7, 256 + 48 + 5] == mh[i, 5 + 16 * 7]).all()
    # an area 16 away is skipped: its pixels stay white
    far = np.array([[116, 100]], np.uint32)
    img2 = oracle.height_map(far, np.zeros((1, 256), np.uint8), cam[0], cam[1])
    assert (img2 == 255).all()


# --------------------------------------------------------------------------- mesh emission
def test_vertex_emission_by_hand():
    """One Grass voxel alone at global (1_000_003, 1_000_005, 10): six faces in push order
    Left, Right, Front, Back, Down, Up with the vertex tables of mesh_generator.rs:222-487."""
    ax, ay = 62500, 62500
    vox = np.zeros((9, 32768), np.uint8)
    vox[4][3 + 16 * 5 + 256 * 10] = V["Grass"]
    mh = np.stack([oracle.update_all_column_heights(v) for v in vox])
    w = oracle.World(ax - 1, ay - 1, 3, 3, vox, mh)
    assert w.mesh_area(ax, ay) == 6
    recs, verts, idx = w.emit_area(ax, ay)
    assert len(recs) == 1
    assert (int(recs[0]["gx"]), int(recs[0]["gy"])) == (1_000_003, 1_000_005)
    assert int(recs[0]["packed"]) == 10 | (V["Grass"] << 8) | (0x3F << 16) | (6 << 24)
    assert np.array_equal(idx[:12], [0, 1, 2, 0, 2, 3, 4, 5, 6, 4, 6, 7])
    assert np.array_equal(idx[30:36], np.array([0, 1, 2, 0, 2, 3]) + 20)
    x, y, z, h = 3.0, 5.0, 10.0, 0.5
    left = verts[0:4]  # x+1 face, normal +x, side uv of a "different faces" voxel (v: 1 / 0.75)
    assert np.allclose(left["position"], [[x + h, y - h, z + h], [x + h, y - h, z - h], [x + h, y + h, z - h], [x + h, y + h, z + h]])
    assert np.allclose(left["uv"], [[1, 1], [1, 0.75], [0, 0.75], [0, 1]])
    assert (left["normal"] == [1, 0, 0, 0]).all() and (left["color"] == 255).all()
    up = verts[20:24]  # z-1 face, top uv (0.625 / 0.375)
    assert np.allclose(up["position"], [[x + h, y - h, z - h], [x - h, y - h, z - h], [x - h, y + h, z - h], [x + h, y + h, z - h]])
    assert np.allclose(up["uv"], [[0, 0.625], [1, 0.625], [1, 0.375], [0, 0.375]])
    assert (up["normal"] == [0, 0, -1, 0]).all()
    down = verts[16:20]  # bottom uv (0.25 / 0)
    assert np.allclose(down["uv"], [[0, 0.25], [1, 0.25], [1, 0], [0, 0]])


def test_partial_height_water_and_enclosed_quirk():
    """Water2 lowers its top by 2/5 and shifts side uv (mesh_generator.rs:214-220,490-498); an
    enclosed interior Water voxel gets no mesh on full-area load but an Up face on update_location
    (world.rs:137-139 vs renderer.rs:173-180)."""
    ax, ay = 10, 10
    vox = np.zeros((9, 32768), np.uint8)
    a = vox[4].reshape(128, 16, 16)
    a[50, 8, 8] = V["Water2"]                      # alone in air
    a[60:63, 4:7, 4:7] = V["Stone"]
    a[61, 5, 5] = V["Water1"]                      # enclosed by stone on all six sides
    mh = np.stack([oracle.update_all_column_heights(v) for v in vox])
    w = oracle.World(ax - 1, ay - 1, 3, 3, vox, mh)
    w.mesh_area(ax, ay)
    fm = w.face_mask[4].reshape(128, 16, 16)
    assert fm[50, 8, 8] == 0x3F
    assert fm[61, 5, 5] == 0                       # dropped by the candidate pre-filter
    recs, verts, _ = w.emit_area(ax, ay)
    k = int(np.flatnonzero((recs["packed"] & 0xFF) == 50)[0])
    v = verts[4 * int(recs[k]["first_face"]):][:24]
    ho = np.float32(2.0) / np.float32(5.0)
    zt = np.float32(50.0) - np.float32(0.5) + ho
    assert v[1]["position"][2] == zt and v[0]["position"][2] == np.float32(50.5)  # Left: zb then zt
    assert v[0]["uv"][1] == np.float32(1.0) - ho and v[1]["uv"][1] == np.float32(0.0)
    assert (v[20:24]["position"][:, 2] == zt).all() and v[20]["uv"][1] == np.float32(1.0)  # Up: uv untouched
    # incremental update of the enclosed voxel: no pre-filter, top-face rule applies
    gx, gy = ax * 16 + 5, ay * 16 + 5
    w.apply_edit(gx, gy, 61, V["Water1"])
    assert w.face_mask[4].reshape(128, 16, 16)[61, 5, 5] == oracle.lib().vgo_should_generate_top_face(0, 0) * 0 + 32


def test_apply_edit_updates_six_neighbours_and_heights():
    """renderer.rs:239-286 + world.rs:184-191: breaking a surface voxel exposes the voxel below and
    the four side neighbours; the column height moves down."""
    seed = oracle.hash_world_name("test")
    w = oracle.World.generate(seed, 62499, 62499, 3, 3)
    ax = ay = 62500
    w.mesh_area(ax, ay)
    # an interior column whose top is plain terrain (not a tree canopy with air underneath)
    col = next(c for c in range(17, 239) if 0 < c % 16 < 15 and
               w.voxels[4][c + 256 * (int(w.max_height[4][c]) + 1)] == V["Stone"] or
               w.voxels[4][c + 256 * (int(w.max_height[4][c]) + 1)] == V["Dirt"])
    lx, ly = col % 16, col // 16
    top = int(w.max_height[4][col])
    gx, gy = ax * 16 + lx, ay * 16 + ly
    before = w.face_mask[4].copy()
    w.apply_edit(gx, gy, top, V["None"])
    assert w.voxels[4][col + 256 * top] == 0
    assert w.max_height[4][col] > top
    assert w.face_mask[4][col + 256 * top] == 0                       # mesh removed
    assert w.face_mask[4][col + 256 * (top + 1)] & 32                 # voxel below gains Up
    changed = np.flatnonzero(before != w.face_mask[4])
    assert 1 <= changed.size <= 7


def test_batched_updates_equal_sequential_updates():
    """The batching rule vx_update_locations relies on (csrc/vx_update.cu header): after a batch of
    world.set + update_location pairs applied one by one (renderer.rs:239-286), every touched voxel
    holds generate_mesh_for_voxel of the FINAL voxels, so "all sets first, then all updates" leaves
    the same Renderer.meshes -- including enclosed partial-height water that only an update gives
    an Up face.  Checked on the oracle alone, with all 27 voxel types and duplicate locations."""
    rng = np.random.default_rng(21)
    n = 3
    vox = rng.integers(0, 27, (n * n, 32768)).astype(np.uint8)
    vox[rng.random(vox.shape) < 0.4] = 0
    vox.reshape(n * n, 128, 16, 16)[:, 60:70] = rng.integers(16, 20, (n * n, 10, 16, 16))  # a water slab
    mh = np.stack([oracle.update_all_column_heights(v) for v in vox])
    seq = oracle.World(10, 20, n, n, vox.copy(), mh.copy())
    bat = oracle.World(10, 20, n, n, vox.copy(), mh.copy())
    for w in (seq, bat):
        w.mesh_area(11, 21)
    k = 4000
    gx = rng.integers(11 * 16 - 1, 12 * 16 + 1, k)
    gy = rng.integers(21 * 16 - 1, 22 * 16 + 1, k)
    gz = rng.integers(0, 128, k)
    gx[-50:], gy[-50:], gz[-50:] = gx[:50], gy[:50], gz[:50]
    new = rng.integers(0, 27, k)
    for i in range(k):
        seq.apply_edit(int(gx[i]), int(gy[i]), int(gz[i]), int(new[i]))
    for i in range(k):  # all sets (array order: the last writer wins) ...
        slot = bat.slot(int(gx[i]) // 16, int(gy[i]) // 16)
        bat.voxels[slot][(gx[i] % 16) + 16 * (gy[i] % 16) + 256 * gz[i]] = new[i]
    for i in rng.permutation(k):  # ... then all updates, in any order
        bat.update_location(int(gx[i]), int(gy[i]), int(gz[i]))
    assert np.array_equal(seq.voxels, bat.voxels)
    assert np.array_equal(seq.face_mask, bat.face_mask)
    water = (seq.voxels >= V["Water1"]) & (seq.voxels <= V["Water4"])
    assert ((seq.face_mask == 32) & water).sum() > 0
