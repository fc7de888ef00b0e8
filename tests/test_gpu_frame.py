"""GPU parity of the per-frame visibility pass (vx_visible, K7) against oracle/frame.c:
same visible areas, same drawn voxel meshes in the same order, same face total and lamps."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg

pytestmark = pytest.mark.gpu

SPAWN = 62500
OFF = 1_000_000


@pytest.fixture(scope="module")
def meshed(seed):
    """12 x 12 resident areas around the spawn, the inner 10 x 10 meshed, a few lamps placed."""
    c = vg.Context(0, seed=seed)
    ax0, ay0 = SPAWN - 6, SPAWN - 6
    c.generate(vg.region_xy(ax0, ay0, 12, 12), read=False)
    inner = vg.region_xy(ax0 + 1, ay0 + 1, 10, 10)
    # lamps (world_actions place path): into the air two voxels above the surface of a few columns
    vox, mh = c.read_areas(inner)
    edits = []
    for k in (0, 17, 55, 99):
        axk, ayk = int(inner[k][0]), int(inner[k][1])
        z = int(mh[k][3 + 16 * 5]) - 2
        edits.append((axk * 16 + 3, ayk * 16 + 5, z, oracle.V["Lamp"]))
    c.apply_edits(np.array(edits, vg.EDIT_DTYPE))
    info = c.mesh(inner)
    rec_first, face_first, recs, _, _ = c.mesh_read(info)
    yield c, inner, recs, rec_first
    c.close()


def _surface_z(recs):
    return int(np.median(recs["packed"] & 0xFF))


VIEWS = [
    # (position offset from the spawn corner, target offset, render_size, head_in_water, show_map)
    ((8.0, 8.0, -3.0), (9.0, 8.0, -3.0), 8, False, False),      # standing on the surface, looking +x
    ((8.0, 8.0, -3.0), (8.0, -9.0, -1.0), 8, False, False),     # looking -y, slightly down
    ((8.5, 8.25, -30.0), (8.5, 8.25, 0.0), 8, False, False),    # above the terrain looking straight down
    ((8.0, 8.0, -3.0), (9.0, 9.0, -8.0), 4, False, False),      # looking up and diagonally, short distance
    ((40.0, -20.0, 2.0), (10.0, 8.0, 2.0), 8, True, False),     # under water
    ((8.0, 8.0, -3.0), (9.0, 8.0, -3.0), 8, False, True),       # map view: everything
    ((8.0, 8.0, -3.0), (8.0, 8.0, -3.0), 8, False, False),      # degenerate camera: look = 0
    ((3000.0, 3000.0, 0.0), (3001.0, 3000.0, 0.0), 8, False, False),  # far away, looking away
]


@pytest.mark.parametrize("case", range(len(VIEWS)))
def test_visible_matches_oracle(meshed, case):
    c, inner, recs, rec_first = meshed
    pos, tgt, render_size, wet, show_map = VIEWS[case]
    z0 = _surface_z(recs)
    cam = (pos[0], pos[1], z0 + pos[2])
    target = (tgt[0], tgt[1], z0 + tgt[2])
    view = vg.make_view(cam, target, render_size, wet, show_map)
    oview = oracle.make_view(cam, target, render_size, wet, show_map)
    assert bytes(view) == bytes(oview)  # the two host helpers build the same 40 bytes
    info, idx, flags = c.visible(view)
    ref = oracle.filter_visible(recs, rec_first, inner, oview)
    assert np.array_equal(flags, ref["area_visible"])
    assert info["n_areas"] == int(ref["area_visible"].sum())
    assert np.array_equal(info["type_first"], ref["type_first"])
    assert np.array_equal(idx, ref["visible"])  # same groups, same order inside the groups
    assert info["n_visible"] == len(ref["visible"]) and info["n_faces"] == ref["n_faces"]
    assert info["n_lamps"] == len(ref["lamps"])
    assert np.array_equal(info["lamps"], ref["lamps"][:64])
    if case == 0:
        assert 0 < info["n_visible"] < len(recs) and 0 < info["n_areas"] < len(inner)
    if show_map:
        assert info["n_visible"] == len(recs) and info["n_lamps"] == 4
    if case == 7:
        assert info["n_visible"] == 0 and info["n_areas"] == 0


def test_visible_is_a_function_of_the_view_only(meshed):
    """Two calls with different views do not disturb each other (scratch state is reset)."""
    c, inner, recs, rec_first = meshed
    z0 = _surface_z(recs)
    a = vg.make_view((8.0, 8.0, z0 - 3.0), (9.0, 8.0, z0 - 3.0))
    b = vg.make_view((8.0, 8.0, z0 - 3.0), (7.0, 8.0, z0 - 3.0))
    ia, xa, fa = c.visible(a)
    ib, xb, fb = c.visible(b)
    ia2, xa2, fa2 = c.visible(a)
    assert np.array_equal(xa, xa2) and np.array_equal(fa, fa2) and ia["n_faces"] == ia2["n_faces"]
    assert not np.array_equal(fa, fb)


def test_visible_needs_a_mesh(seed):
    with vg.Context(0, seed=seed) as c:
        with pytest.raises(vg.VxError) as e:
            c.visible(vg.make_view((0, 0, 0), (1, 0, 0)), read=False)
        assert e.value.status == -7  # VX_ERR_NO_MESH
        c.generate(vg.region_xy(SPAWN, SPAWN, 1, 1), read=False)
        c.mesh(np.zeros((0, 2), np.uint32))
        info = c.visible(vg.make_view((0, 0, 0), (1, 0, 0)), read=False)
        assert info["n_visible"] == 0 and info["n_areas"] == 0


def test_visible_full_region_properties(ctx):
    """Config 3 size (128 x 128 areas): size-independent properties instead of a CPU comparison --
    every index appears once, groups are sorted by type, the map view selects every record."""
    n = 128
    xy = vg.region_xy(SPAWN - n // 2, SPAWN - n // 2, n, n)
    ctx.generate(xy, read=False)
    info = ctx.mesh(xy)
    rec_first, recs = np.empty(n * n + 1, np.uint32), np.empty(info["n_recs"], vg.REC_DTYPE)
    ctx.mesh_read(info, out=(rec_first, None, recs, None, None))  # records only: no vertices needed
    z0 = _surface_z(recs)
    view = vg.make_view((8.0, 8.0, z0 - 3.0), (9.0, 8.5, z0 - 2.0), render_size=64)
    vi, idx, flags = ctx.visible(view)
    assert len(np.unique(idx)) == len(idx)
    types = (recs["packed"][idx] >> 8) & 0xFF
    assert np.all(np.diff(types.astype(np.int32)) >= 0)
    assert np.array_equal(np.searchsorted(types, np.arange(28)), vi["type_first"])
    area_of = np.searchsorted(rec_first, idx, side="right") - 1
    assert flags[area_of].all()
    assert vi["n_faces"] == int((recs["packed"][idx] >> 24).sum())
    ref = oracle.filter_visible(recs, rec_first, xy, oracle.make_view((8.0, 8.0, z0 - 3.0), (9.0, 8.5, z0 - 2.0), 64))
    assert np.array_equal(idx, ref["visible"]) and np.array_equal(flags, ref["area_visible"])
    everything, idx2, flags2 = ctx.visible(vg.make_view((0, 0, 0), (1, 0, 0), show_map=True))
    assert everything["n_visible"] == len(recs) and flags2.all()
    assert np.array_equal(np.sort(idx2), np.arange(len(recs), dtype=np.uint32))
