"""GPU parity of vx_explode against oracle/physics.c (BombSimulator::handle_explosions / explode_at,
bomb_simulator.rs:191-262): removed voxels and found bombs in the reference's visiting order, final
block IDs and column heights, dirty areas, and the meshes of the dirty areas after re-meshing."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg
from test_gpu_parity import _compare_mesh

pytestmark = pytest.mark.gpu

SPAWN = 62500
OFF = 1_000_000


def _setup(seed, n=5):
    ax0 = ay0 = SPAWN - n // 2
    world = oracle.World.generate(seed, ax0, ay0, n, n)
    c = vg.Context(0, seed=seed)
    xy = vg.region_xy(ax0, ay0, n, n)
    c.generate(xy, read=False)
    return c, world, xy, ax0, ay0


def _surface(world, ax0, n, x, y):
    """z of the first non-transparent voxel of the column at player coordinates (x, y)."""
    gx, gy = x + OFF, y + OFF
    slot = (gx // 16 - ax0) + n * (gy // 16 - ax0)
    return int(world.max_height[slot][gx % 16 + 16 * (gy % 16)])


def _edit_both(c, world, edits):
    for gx, gy, gz, v in edits:
        world.apply_edit(gx, gy, gz, v)
    c.apply_edits(np.array(edits, vg.EDIT_DTYPE))


def test_explosions_match_oracle(seed):
    n = 5
    c, world, xy, ax0, ay0 = _setup(seed, n)
    try:
        h = _surface(world, ax0, n, 8, 8)
        h2 = _surface(world, ax0, n, 15, 0)
        V = oracle.V
        # things an explosion treats differently: bombs (chain), water (not solid), glass (solid, transparent)
        _edit_both(c, world, [(OFF + 9, OFF + 8, h, V["Bomb"]), (OFF + 6, OFF + 9, h + 1, V["Bomb"]),
                              (OFF + 8, OFF + 10, h - 1, V["WaterSource"]), (OFF + 7, OFF + 7, h - 1, V["Glass"]),
                              (OFF + 11, OFF + 9, h + 2, V["Water2"]), (OFF + 14, OFF + 1, h2, V["Bomb"])])
        positions = [
            (8.5, 8.5, h + 0.5),        # in the ground at the spawn column
            (10.2, 9.7, h + 1.3),       # overlaps the first one
            (15.9, 0.1, h2 + 0.2),      # on an area corner: four areas, border columns
            (-20.5, 3.5, 2.0),          # high in the air: z_min clamps at 0, nothing solid
            (20.5, -9.5, 125.6),        # at the bottom: z_max clamps at 126, the bedrock slab z = 127 stays
            (8.5, 8.5, h + 0.5),        # the same place twice in one tick
        ]
        removed, bombs, dirty = c.explode(positions)
        oremoved, obombs = world.explode(positions)
        assert len(oremoved) > 300 and len(obombs) == 3
        assert np.array_equal(removed, oremoved), "removed voxels / visiting order differ"
        assert np.array_equal(bombs, obombs)
        vox, mh = c.read_areas(xy)
        assert np.array_equal(vox, world.voxels) and np.array_equal(mh, world.max_height)
        assert (vox.reshape(-1, 128, 256)[:, 127] != 0).all()  # z = 127 untouched (bomb_simulator.rs:236)
        # dirty areas: every area that lost a voxel, plus edge neighbours of border columns
        touched = {(int(x) // 16, int(y) // 16) for x, y, _ in removed}
        got = {(int(a), int(b)) for a, b in dirty}
        assert touched <= got
        assert len(got) <= len(touched) + 8 and got <= {(int(a), int(b)) for a, b in xy}
        # re-mesh what the renderer would touch: equals the oracle's meshes of the changed world
        inner = {(ax0 + i, ay0 + j) for i in range(1, n - 1) for j in range(1, n - 1)}
        remesh = np.array(sorted(got & inner), np.uint32)
        assert len(remesh) >= 4
        assert _compare_mesh(c, world, remesh) > 0
    finally:
        c.close()


def test_explosion_errors_and_limits(seed):
    c, world, xy, ax0, ay0 = _setup(seed, 3)
    try:
        before, _ = c.read_areas(xy)
        with pytest.raises(vg.VxError) as e:   # reaches into an area that is not resident: nothing happens
            c.explode([(8.5, 8.5, 70.0), (40.0, 8.0, 70.0)])
        assert e.value.status == -5
        after, _ = c.read_areas(xy)
        assert np.array_equal(before, after)
        with pytest.raises(vg.VxError):
            c.explode(np.zeros((33, 3), np.float32))  # more explosions than one tick can hold
        r, b, d = c.explode(np.zeros((0, 3), np.float32))
        assert len(r) == 0 and len(b) == 0 and len(d) == 0
    finally:
        c.close()
