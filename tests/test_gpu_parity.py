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
