"""GPU tests of the host side above the C ABI: the compact record transport expanded by the C++
host mirror equals the device's own vertex stream, and the mirror's Renderer
(load_all_blocking / update_locations, per-voxel Mesh maps) ends up holding the oracle's meshes."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg

pytestmark = pytest.mark.gpu

SPAWN = 62500
V = oracle.V


def test_host_expanded_vertices_equal_the_device_stream(ctx, seed):
    """vx_mesh_read_compact + expand_records (host, 4 B per voxel over PCIe) == vx_mesh_read (device
    vertex kernel, 16 B + 172 B per face over PCIe): rec_first, face_first, records, vertices,
    indices byte for byte over 12x12 areas with caves, a lake and trees."""
    n = 12
    ax0, ay0 = SPAWN + 2, SPAWN - 2
    ctx.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
    xy = vg.region_xy(ax0, ay0, n, n)
    info = ctx.mesh(xy)
    rec_first, face_first, recs, verts, idx = ctx.mesh_read(info)
    rf, packed = ctx.mesh_read_compact(info)
    assert np.array_equal(rf, rec_first) and len(packed) == info["n_recs"]
    for threads in (1, 4, 0):
        ff, hrecs, hverts, hidx = vg.expand_records(xy, rf, packed, threads=threads)
        assert np.array_equal(ff, face_first)
        assert hrecs.tobytes() == recs.tobytes()
        assert hverts.tobytes() == verts.tobytes()
        assert np.array_equal(hidx, idx)
    assert info["n_faces"] > 50_000


def _check_renderer_against_oracle(hr, world, areas):
    faces = 0
    for ax, ay in areas:
        got = hr.area_face_counts(int(ax), int(ay))
        assert got is not None, f"area {ax},{ay} has no RenderArea"
        counts, lamps = got
        fm = world.face_mask[world.slot(int(ax), int(ay))]
        expect = np.unpackbits(fm.reshape(-1, 1), axis=1).sum(axis=1).astype(np.uint8)
        assert np.array_equal(counts, expect), f"face counts differ in area {ax},{ay}"
        assert lamps == int(((world.voxels[world.slot(int(ax), int(ay))] == V["Lamp"]) & (fm != 0)).sum())
        faces += int(expect.sum())
    return faces


@pytest.mark.parametrize("transport", ["compact", "raw"])
def test_host_renderer_load_update_unload(seed, transport):
    """Renderer::load_all_blocking, update_location and unload_area of the C++ mirror against the
    oracle: the per-voxel Mesh maps (face counts per voxel, lamp sets, vertex bytes of sampled
    voxels) after a full load, after edits through the incremental path, and after an unload."""
    n = 5
    ax0, ay0 = SPAWN - 2, SPAWN - 2
    world = oracle.World.generate(seed, ax0 - 1, ay0 - 1, n + 2, n + 2)
    inner = vg.region_xy(ax0, ay0, n, n)
    with vg.Context(0, seed=seed) as c:
        c.generate(world.area_xy(), read=False)
        with vg.HostRenderer(c) as hr:
            ms = hr.load_all_blocking(inner, transport=transport, threads=4)
            assert ms > 0 and hr.area_count() == n * n
            for ax, ay in inner:
                world.mesh_area(int(ax), int(ay))
            assert _check_renderer_against_oracle(hr, world, inner) == hr.face_count()
            assert hr.voxel_count() == int((world.face_mask != 0).sum())
            before = hr.face_count()
            hr.load_all_blocking(inner, transport=transport)  # already meshed: no-op (renderer.rs:311)
            assert hr.face_count() == before
            # edits: lamps in the air, holes in the ground, water poured into a hole
            rng = np.random.default_rng(4)
            k = 300
            gx = rng.integers(ax0 * 16 + 1, (ax0 + n) * 16 - 1, k)  # neighbours stay inside the meshed block
            gy = rng.integers(ay0 * 16 + 1, (ay0 + n) * 16 - 1, k)
            surface = np.array([world.max_height[world.slot(x // 16, y // 16)][(x % 16) + 16 * (y % 16)] for x, y in zip(gx, gy)])
            gz = np.clip(surface + rng.integers(-2, 3, k), 1, 126)
            edits = np.zeros(k, vg.EDIT_DTYPE)
            edits["gx"], edits["gy"], edits["gz"] = gx, gy, gz
            edits["voxel"] = rng.choice(np.array([0, V["Lamp"], V["Water3"], V["Glass"], V["Grass"]], np.uint32), k)
            for e in edits:
                world.apply_edit(int(e["gx"]), int(e["gy"]), int(e["gz"]), int(e["voxel"]))
            c.apply_edits(edits)
            hr.update_locations(np.stack([gx, gy, gz], axis=1))
            assert _check_renderer_against_oracle(hr, world, inner) == hr.face_count()
            # vertex bytes of the edited voxels
            meshes = {}
            for ax, ay in inner:
                recs, verts, idx = world.emit_area(int(ax), int(ay))
                for r in recs:
                    nf, f0 = int(r["packed"]) >> 24, int(r["first_face"])
                    meshes[(int(r["gx"]), int(r["gy"]), int(r["packed"]) & 0xFF)] = (verts[4 * f0:4 * (f0 + nf)], idx[6 * f0:6 * (f0 + nf)])
            checked = 0
            for x, y, z in zip(gx, gy, gz):
                v, i = hr.voxel_mesh(int(x), int(y), int(z))
                if (int(x), int(y), int(z)) in meshes:
                    ov, oi = meshes[(int(x), int(y), int(z))]
                    assert v.tobytes() == ov.tobytes() and np.array_equal(i, oi)
                    checked += 1
                else:
                    assert len(v) == 0
            assert checked > 50
            hr.unload_area(ax0, ay0)
            assert hr.area_face_counts(ax0, ay0) is None and hr.area_count() == n * n - 1
            masks, meshed = c.read_face_masks(inner)
            assert meshed[0] == 0 and not masks[0].any() and meshed[1:].all()
