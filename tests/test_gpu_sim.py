"""GPU parity of the simulator loops (vx_water_tick, vx_falling_check, vx_falling_land)
against the oracle's restatement of water_simulator.rs / falling_voxel_simulator.rs on ORDERED lists:
the same list in the same order gives the same changes in the same order, the same queue for the next
tick, the same voxels, column heights and -- through vx_update_locations -- the same meshes."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg

pytestmark = pytest.mark.gpu
V = oracle.V
AX0, AY0, N = 300, 400, 4


def _terrain(rng):
    """4 x 4 areas: rock below z = 90 with pits and overhangs, loose sand / dirt / grass / snow layers and
    pillars above caves, water sources and flowing water of every level scattered over it."""
    vox = np.zeros((N * N, 128, 16, 16), np.uint8)
    vox[:, 90:] = V["Stone"]
    r = rng.random((N * N, 128, 16, 16))
    vox[:, 90:110][r[:, 90:110] < 0.25] = V["None"]                 # caves in the rock
    top = vox[:, 70:90]
    rt = r[:, 70:90]
    loose = np.array([V["Sand"], V["Dirt"], V["Grass"], V["Snow"]], np.uint8)
    top[rt < 0.30] = loose[rng.integers(0, 4, int((rt < 0.30).sum()))]
    top[(rt >= 0.30) & (rt < 0.34)] = V["Stone"]
    top[(rt >= 0.34) & (rt < 0.36)] = V["WaterSource"]
    water = np.array([V["WaterDown"], V["Water1"], V["Water2"], V["Water3"], V["Water4"]], np.uint8)
    m = (rt >= 0.36) & (rt < 0.42)
    top[m] = water[rng.integers(0, 5, int(m.sum()))]
    top[(rt >= 0.42) & (rt < 0.43)] = V["Glass"]
    return vox.reshape(N * N, 32768)


def _load(ctx, rng):
    xy = vg.region_xy(AX0, AY0, N, N)
    vox = _terrain(rng)
    ctx.set_mesh_retention(True)
    mh = ctx.upload(xy, vox)
    world = oracle.World(AX0, AY0, N, N, vox.copy(), mh.copy())
    inner = vg.region_xy(AX0 + 1, AY0 + 1, N - 2, N - 2)
    ctx.mesh(inner)
    for ax, ay in inner:
        world.mesh_area(int(ax), int(ay))
    return xy, inner, world


def _inner_locations(rng, k, z_lo=60, z_hi=112):
    x = rng.integers((AX0 + 1) * 16, (AX0 + N - 1) * 16, k)
    y = rng.integers((AY0 + 1) * 16, (AY0 + N - 1) * 16, k)
    z = rng.integers(z_lo, z_hi, k)
    return np.stack([x, y, z], axis=1).astype(np.uint32)


def _same_world(ctx, world, xy, inner):
    vox, mh = ctx.read_areas(xy)
    assert np.array_equal(vox, world.voxels), "voxels differ"
    assert np.array_equal(mh, world.max_height), "column heights differ"
    # the S / T planes the mesher reads were kept current: a full re-mesh equals the oracle's
    info = ctx.mesh(inner)
    want = sum(world.mesh_area(int(a), int(b)) for a, b in inner)
    assert info["n_faces"] == want


def _inside(locs):
    """locations whose own column and four sides lie in the inner 2 x 2 areas (what the queue may hold)"""
    locs = np.asarray(locs, np.int64).reshape(-1, 3)
    lo_x, hi_x = (AX0 + 1) * 16 + 1, (AX0 + N - 1) * 16 - 1
    lo_y, hi_y = (AY0 + 1) * 16 + 1, (AY0 + N - 1) * 16 - 1
    ok = (locs[:, 0] >= lo_x) & (locs[:, 0] < hi_x) & (locs[:, 1] >= lo_y) & (locs[:, 1] < hi_y) & (locs[:, 2] < 128)
    return locs[ok].astype(np.uint32)


def test_water_ticks_equal_the_reference_loop_in_the_given_order(seed):
    rng = np.random.default_rng(21)
    with vg.Context(0, seed=seed) as ctx:
        xy, inner, world = _load(ctx, rng)
        check = _inside(_inner_locations(rng, 3000, 65, 95))
        total_changed = 0
        for tick in range(8):
            assert len(check) > 0
            changed, nxt = ctx.water_tick(check)
            ochanged, onxt = world.water_tick(check)
            assert np.array_equal(changed, ochanged), f"tick {tick}: changes differ"
            assert np.array_equal(nxt, onxt), f"tick {tick}: next tick's queue differs"
            total_changed += len(changed)
            if len(changed):  # renderer.update_location after every set: the batch form on the device
                ctx.update_locations(changed[:, :3], read=False)
                masks, _ = ctx.read_face_masks(xy)
                slots = [world.slot(int(a), int(b)) for a, b in inner]
                assert np.array_equal(masks[slots], world.face_mask[slots]), f"tick {tick}: meshes differ"
            # the next tick's set, in an order of our choosing (a HashSet would pick its own)
            uniq = np.unique(_inside(nxt), axis=0)
            check = uniq[rng.permutation(len(uniq))]
        assert total_changed > 500
        _same_world(ctx, world, xy, inner)


def test_water_order_is_part_of_the_input(seed):
    """The same set in two orders gives two different worlds (the reason the order is an argument), and
    each equals the oracle for ITS order."""
    rng = np.random.default_rng(22)
    results = []
    for order in (0, 1):
        with vg.Context(0, seed=seed) as ctx:
            xy, inner, world = _load(ctx, np.random.default_rng(5))
            check = np.unique(_inside(_inner_locations(np.random.default_rng(6), 4000, 65, 95)), axis=0)
            if order:
                check = check[::-1].copy()
            changed, nxt = ctx.water_tick(check)
            ochanged, onxt = world.water_tick(check)
            assert np.array_equal(changed, ochanged) and np.array_equal(nxt, onxt)
            vox, _ = ctx.read_areas(xy)
            results.append(vox)
    assert (results[0] != results[1]).any()


def test_falling_voxels_check_and_land(seed):
    rng = np.random.default_rng(23)
    with vg.Context(0, seed=seed) as ctx:
        xy, inner, world = _load(ctx, rng)
        # knock supports out: edits on both sides, then update_voxels for the edited locations in order
        locs = _inside(_inner_locations(rng, 1500, 70, 110))
        edits = np.zeros(len(locs), vg.EDIT_DTYPE)
        edits["gx"], edits["gy"], edits["gz"], edits["voxel"] = locs[:, 0], locs[:, 1], locs[:, 2], 0
        ctx.apply_edits(edits)
        for gx, gy, gz in locs:
            world.apply_edit(int(gx), int(gy), int(gz), 0)
        spawned, nxt = ctx.falling_check(locs)
        ospawned, onxt = world.falling_check(locs)
        assert np.array_equal(spawned, ospawned) and np.array_equal(nxt, onxt)
        assert len(spawned) > 300 and len(nxt) == 7 * len(spawned)
        # chains: some call removed several voxels of one column
        cols = spawned[:, 0].astype(np.int64) * (1 << 32) + spawned[:, 1]
        assert np.unique(cols).size < len(spawned)
        vox, mh = ctx.read_areas(xy)
        assert np.array_equal(vox, world.voxels) and np.array_equal(mh, world.max_height)

        # the entities fall: simulate_falling's arithmetic on the host (f32), the retain pass on the device
        pos = np.stack([spawned[:, 0].astype(np.int64) - 1_000_000, spawned[:, 1].astype(np.int64) - 1_000_000,
                        spawned[:, 2]], axis=1).astype(np.float32)
        types = spawned[:, 3].astype(np.uint8)
        vel = np.zeros(len(pos), np.float32)
        placed_total = 0
        for step in range(40):
            if len(pos) == 0:
                break
            vel = np.minimum(vel + np.float32(0.05) * np.float32(0.2) * np.float32(60), np.float32(3.0)).astype(np.float32)
            pos[:, 2] += vel
            keep, placed, wn = ctx.falling_land(pos, types)
            okeep, oplaced, own = world.falling_land(pos, types)
            assert np.array_equal(keep, okeep) and np.array_equal(placed, oplaced) and np.array_equal(wn, own), f"step {step}"
            placed_total += len(placed)
            k = keep.astype(bool)
            pos, types, vel = pos[k], types[k], vel[k]
        assert placed_total > 100 and len(pos) < 20
        ctx.update_locations(np.concatenate([spawned[:, :3]]), read=False)
        _same_world(ctx, world, xy, inner)


def test_simulator_error_paths(seed):
    with vg.Context(0, seed=seed) as ctx:
        xy, inner, world = _load(ctx, np.random.default_rng(3))
        before, _ = ctx.read_areas(xy)
        edge = [((AX0 + N) * 16 - 1, (AY0 + 1) * 16 + 3, 80)]  # its x + 1 side is in an area that is not resident
        for call in (ctx.water_tick, ctx.falling_check):
            with pytest.raises(vg.VxError) as e:
                call(edge)
            assert e.value.status == vg.binding.VX_ERR_NOT_LOADED
        with pytest.raises(vg.VxError) as e:
            ctx.water_tick([((AX0 + 1) * 16 + 3, (AY0 + 1) * 16 + 3, 128)])
        assert e.value.status == vg.binding.VX_ERR_INVALID
        with pytest.raises(vg.VxError) as e:
            ctx.falling_land([(5.0e6, 0.0, 10.0)], [V["Sand"]])
        assert e.value.status == vg.binding.VX_ERR_INVALID
        with pytest.raises(vg.VxError) as e:
            ctx.falling_land([(-999_000.0, 0.0, 10.0)], [V["Sand"]])  # an area that is not resident
        assert e.value.status == vg.binding.VX_ERR_NOT_LOADED
        after, _ = ctx.read_areas(xy)
        assert np.array_equal(before, after)
        # an output list that is too small is an error, not an overflow
        locs = _inside(_inner_locations(np.random.default_rng(4), 800, 70, 110))
        e2 = np.zeros(len(locs), vg.EDIT_DTYPE)
        e2["gx"], e2["gy"], e2["gz"] = locs[:, 0], locs[:, 1], locs[:, 2]
        ctx.apply_edits(e2)
        with pytest.raises(vg.VxError) as e:
            ctx.falling_check(locs, cap_spawned=4, cap_next=28)
        assert e.value.status == vg.binding.VX_ERR_INVALID
        assert len(ctx.water_tick(np.zeros((0, 3), np.uint32))[0]) == 0
