"""GPU parity of the incremental path: vx_update_locations (Renderer::update_location, M6), mesh
retention (the device mirror of Renderer.meshes), vx_visible_zone and the compact record transport,
each against the oracle's voxel-at-a-time restatement (oracle/mesh.c: vgo_apply_edit,
vgo_update_location, face_mask volume = RenderArea.mesh_map)."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg

pytestmark = pytest.mark.gpu

SPAWN = 62500
V = oracle.V


def _water_world(rng, n=4, ax0=200, ay0=300):
    """n x n uploaded areas: stone below z = 70, a thick body of partial-height water (Water1..4 in
    blobs, so that interior voxels are enclosed by their own kind and by other kinds) between
    z = 40 and z = 70, sprinkled with glass, ice, sources, lamps and holes."""
    vox = np.zeros((n * n, 128, 16, 16), np.uint8)
    vox[:, 70:] = V["Stone"]
    kinds = np.array([V["Water1"], V["Water2"], V["Water3"], V["Water4"]], np.uint8)
    blob = kinds[rng.integers(0, 4, (n * n, 8, 4, 4))]  # 4-voxel cubes of one kind
    vox[:, 40:70] = np.repeat(np.repeat(np.repeat(blob, 4, axis=1), 4, axis=2), 4, axis=3)[:, :30]
    extra = rng.random((n * n, 30, 16, 16))
    body = vox[:, 40:70]
    body[extra < 0.02] = V["None"]
    body[(extra >= 0.02) & (extra < 0.03)] = V["Glass"]
    body[(extra >= 0.03) & (extra < 0.04)] = V["Ice"]
    body[(extra >= 0.04) & (extra < 0.05)] = V["WaterSource"]
    body[(extra >= 0.05) & (extra < 0.055)] = V["Lamp"]
    body[(extra >= 0.055) & (extra < 0.06)] = V["WaterDown"]
    xy = vg.region_xy(ax0, ay0, n, n)
    return xy, vox.reshape(n * n, 32768)


def _load(ctx, xy, vox, n, ax0, ay0):
    mh = ctx.upload(xy, vox)
    world = oracle.World(ax0, ay0, n, n, vox.copy(), mh.copy())
    return world


def _voxel_meshes(world, xy):
    """{(gx, gy, gz): (packed, vertex bytes, index bytes)} of the oracle's current meshes."""
    out = {}
    for ax, ay in xy:
        recs, verts, idx = world.emit_area(int(ax), int(ay), first_face_base=0)
        for r in recs:
            nf = int(r["packed"]) >> 24
            f0 = int(r["first_face"])
            out[(int(r["gx"]), int(r["gy"]), int(r["packed"]) & 0xFF)] = (
                int(r["packed"]), verts[4 * f0:4 * (f0 + nf)].tobytes(), idx[6 * f0:6 * (f0 + nf)].tobytes())
    return out


def _check_update_records(world, inner_xy, recs, verts, idx, touched_expected):
    """Every record of an update equals the oracle's mesh of that voxel (or its absence)."""
    meshes = _voxel_meshes(world, inner_xy)
    seen = set()
    for r in recs:
        key = (int(r["gx"]), int(r["gy"]), int(r["packed"]) & 0xFF)
        assert key not in seen, f"voxel {key} reported twice"
        seen.add(key)
        nf = int(r["packed"]) >> 24
        f0 = int(r["first_face"])
        if key in meshes:
            packed, vb, ib = meshes[key]
            assert int(r["packed"]) == packed, f"record differs at {key}: {int(r['packed']):#x} vs {packed:#x}"
            assert verts[4 * f0:4 * (f0 + nf)].tobytes() == vb, f"vertices differ at {key}"
            assert idx[6 * f0:6 * (f0 + nf)].tobytes() == ib, f"indices differ at {key}"
        else:
            assert nf == 0 and ((int(r["packed"]) >> 16) & 0xFF) == 0, f"{key} should have been removed"
    assert seen == touched_expected


def _touched(world, locs):
    """The voxels update_location visits, restricted to loaded areas (get_without_loading)."""
    out = set()
    for gx, gy, gz in locs:
        cand = [(gx, gy, gz), (gx + 1, gy, gz), (gx - 1, gy, gz), (gx, gy + 1, gz), (gx, gy - 1, gz)]
        if gz > 0:
            cand.append((gx, gy, gz - 1))
        if gz < 127:
            cand.append((gx, gy, gz + 1))
        for c in cand:
            ax, ay = c[0] // 16, c[1] // 16
            if world.ax0 <= ax < world.ax0 + world.nx and world.ay0 <= ay < world.ay0 + world.ny:
                out.add(c)
    return out


def test_update_locations_reproduce_the_enclosed_water_quirk(seed):
    """SURVEY 'quirk to preserve': an interior partial-height water voxel enclosed by water has no
    mesh after a full-area load (world.rs:137-139) but gets its Up face when an update touches it
    (renderer.rs:173-180).  Batched GPU edits + vx_update_locations == the oracle's one-by-one
    world.set + update_location: face-mask volumes, records, vertices, indices."""
    rng = np.random.default_rng(11)
    n, ax0, ay0 = 4, 200, 300
    xy, vox = _water_world(rng, n, ax0, ay0)
    with vg.Context(0, seed=seed) as ctx:
        ctx.set_mesh_retention(True)
        world = _load(ctx, xy, vox, n, ax0, ay0)
        inner = vg.region_xy(ax0 + 1, ay0 + 1, n - 2, n - 2)
        ctx.mesh(inner)
        for ax, ay in inner:
            world.mesh_area(int(ax), int(ay))
        masks, meshed = ctx.read_face_masks(xy)
        assert np.array_equal(masks, world.face_mask), "full-load face masks differ"
        inner_slots = [world.slot(int(a), int(b)) for a, b in inner]
        assert meshed[inner_slots].all() and meshed.sum() == len(inner)
        # the quirk is present in the data: enclosed partial-height water voxels without a mesh
        body = world.voxels[inner_slots].reshape(-1, 128, 16, 16)[:, 45:65, 1:15, 1:15]
        fm = world.face_mask[inner_slots].reshape(-1, 128, 16, 16)[:, 45:65, 1:15, 1:15]
        assert ((body >= V["Water1"]) & (body <= V["Water4"]) & (fm == 0)).sum() > 1000

        total_quirk = 0
        for batch_no, batch in enumerate([1, 7, 64, 500, 2000]):
            gx = rng.integers((ax0 + 1) * 16, (ax0 + n - 1) * 16, batch)
            gy = rng.integers((ay0 + 1) * 16, (ay0 + n - 1) * 16, batch)
            gz = rng.integers(38, 72, batch)
            if batch >= 64:  # duplicates inside a batch: the last writer must win
                gx[-10:], gy[-10:], gz[-10:] = gx[:10], gy[:10], gz[:10]
            new = rng.choice(np.array([V["None"], V["Stone"], V["Water2"], V["Water4"], V["Glass"], V["Lamp"],
                                       V["WaterSource"], V["Sand"]], np.uint32), batch)
            edits = np.zeros(batch, vg.EDIT_DTYPE)
            edits["gx"], edits["gy"], edits["gz"], edits["voxel"] = gx, gy, gz, new
            before = world.face_mask.copy()
            for e in edits:
                world.apply_edit(int(e["gx"]), int(e["gy"]), int(e["gz"]), int(e["voxel"]))
            if batch_no % 2:  # the form that only enqueues (no dirty list): the caller's array may be reused at once
                scratch = edits.copy()
                assert ctx.apply_edits(scratch, want_dirty=False) is None
                scratch[:] = 0
            else:
                ctx.apply_edits(edits)
            locs = np.stack([gx, gy, gz], axis=1).astype(np.uint32)
            info, recs, verts, idx = ctx.update_locations(locs)
            assert info["n_recs"] == len(recs) and info["n_faces"] * 4 == len(verts)
            _check_update_records(world, xy, recs, verts, idx, _touched(world, locs.tolist()))
            masks, _ = ctx.read_face_masks(xy)
            bad = np.flatnonzero((masks != world.face_mask).any(axis=1))
            assert bad.size == 0, f"batch {batch_no}: face masks differ in areas {xy[bad].tolist()}"
            v4, _ = ctx.read_areas(xy)
            assert np.array_equal(v4, world.voxels)
            # voxels that gained ONLY an Up face while still enclosed partial-height water: the quirk
            gained = (before == 0) & (world.face_mask == 32) & (world.voxels >= V["Water1"]) & (world.voxels <= V["Water4"])
            total_quirk += int(gained.sum())
        assert total_quirk > 50, "the edit storm never exercised the enclosed-water quirk"
        # a full re-mesh would NOT reproduce this state (the round-1 behaviour)
        remeshed = oracle.World(ax0, ay0, n, n, world.voxels.copy(), world.max_height.copy())
        for ax, ay in inner:
            remeshed.mesh_area(int(ax), int(ay))
        assert not np.array_equal(remeshed.face_mask, world.face_mask)


def test_update_locations_random_all_types(seed):
    """Random areas with all 27 voxel types, random edits with all 27 types in batches, every
    update compared with the sequential oracle; also an update of an area that was never meshed
    (set_voxel_mesh creates the RenderArea, renderer.rs:208-212)."""
    rng = np.random.default_rng(5)
    n, ax0, ay0 = 3, 5000, 6000
    vox = rng.integers(0, 27, (n * n, 32768)).astype(np.uint8)
    vox[rng.random(vox.shape) < 0.5] = 0
    xy = vg.region_xy(ax0, ay0, n, n)
    with vg.Context(0, seed=seed) as ctx:
        ctx.set_mesh_retention(True)
        world = _load(ctx, xy, vox, n, ax0, ay0)
        centre = np.array([[ax0 + 1, ay0 + 1]], np.uint32)
        ctx.mesh(centre)
        world.mesh_area(ax0 + 1, ay0 + 1)
        has_render_area = {(ax0 + 1, ay0 + 1)}
        for batch in [3, 300, 3000]:
            # inside the centre area and one voxel into its neighbours (never meshed areas)
            gx = rng.integers((ax0 + 1) * 16 - 1, (ax0 + 2) * 16 + 1, batch)
            gy = rng.integers((ay0 + 1) * 16 - 1, (ay0 + 2) * 16 + 1, batch)
            gz = rng.integers(0, 128, batch)
            edits = np.zeros(batch, vg.EDIT_DTYPE)
            edits["gx"], edits["gy"], edits["gz"] = gx, gy, gz
            edits["voxel"] = rng.integers(0, 27, batch)
            for e in edits:
                world.apply_edit(int(e["gx"]), int(e["gy"]), int(e["gz"]), int(e["voxel"]))
            ctx.apply_edits(edits)
            locs = np.stack([gx, gy, gz], axis=1).astype(np.uint32)
            info, recs, verts, idx = ctx.update_locations(locs)
            _check_update_records(world, xy, recs, verts, idx, _touched(world, locs.tolist()))
            masks, meshed = ctx.read_face_masks(xy)
            assert np.array_equal(masks, world.face_mask)
            # set_voxel_mesh creates the RenderArea of every area it is called for, even for a removal
            has_render_area |= {(c[0] // 16, c[1] // 16) for c in _touched(world, locs.tolist())}
            assert {tuple(a) for a in xy[meshed != 0].tolist()} == has_render_area
        mh = ctx.read_areas(xy, voxels=False)[1]
        assert np.array_equal(mh, world.max_height)


def test_update_locations_at_the_edge_of_the_resident_set(seed):
    """Voxels in areas that are not resident are skipped (get_without_loading -> None); an area one
    voxel beyond a touched voxel is generated for its face test (get_with_cache -> load_area)."""
    ax0, ay0 = SPAWN - 1, SPAWN - 1
    big = oracle.World.generate(seed, ax0 - 2, ay0 - 2, 6, 6)  # what generation yields everywhere
    inner = vg.region_xy(ax0, ay0, 2, 2)
    with vg.Context(0, seed=seed) as ctx:
        ctx.generate(inner, read=False)
        assert ctx.resident_count() == 4
        _, mh = ctx.read_areas(inner)
        z = int(mh[0][15 + 16 * 3])  # surface of column (15, 3) of area (ax0, ay0)... and its -x twin
        locs = np.array([[ax0 * 16, ay0 * 16 + 3, int(mh[0][0 + 16 * 3])],          # x = 0: -x neighbour area not resident
                         [(ax0 + 1) * 16 + 15, (ay0 + 1) * 16 + 15, z]], np.uint32)   # corner of the resident block
        info, recs, verts, idx = ctx.update_locations(locs)
        # expectation: the oracle on the large world, restricted to voxels of the four resident areas
        for l in locs:
            big.update_location(*[int(v) for v in l])
        resident = {(int(a), int(b)) for a, b in inner}
        expected = {c for c in _touched(big, locs.tolist()) if (c[0] // 16, c[1] // 16) in resident}
        got = {(int(r["gx"]), int(r["gy"]), int(r["packed"]) & 0xFF) for r in recs}
        assert got == expected
        for r in recs:
            gx, gy, gz = int(r["gx"]), int(r["gy"]), int(r["packed"]) & 0xFF
            m = big.face_mask[big.slot(gx // 16, gy // 16)][(gx % 16) + 16 * (gy % 16) + 256 * gz]
            assert ((int(r["packed"]) >> 16) & 0xFF) == m
        assert ctx.resident_count() > 4  # the areas beyond the border were generated, not guessed


def _zone_reference(world, zone_xy, meshed_set, view):
    """The oracle's frame over Renderer.meshes restricted to the zone list."""
    keep = [i for i, (a, b) in enumerate(zone_xy) if (int(a), int(b)) in meshed_set]
    sub = zone_xy[keep]
    rec_first, recs = world.records(sub)
    ref = oracle.filter_visible(recs, rec_first, sub, view)
    area_of = np.repeat(np.arange(len(sub)), np.diff(rec_first.astype(np.int64)))
    def code(q):
        r = recs[q]
        local = (int(r["gx"]) % 16) + 16 * (int(r["gy"]) % 16) + 256 * (int(r["packed"]) & 0xFF)
        return (keep[area_of[q]] << 15) | local
    codes = np.array([code(q) for q in ref["visible"]], np.uint32)
    lamps = np.array([code(q) for q in ref["lamps"]], np.uint32)
    flags = np.zeros(len(zone_xy), np.uint8)
    flags[keep] = ref["area_visible"]
    return codes, flags, ref, lamps


def test_visible_zone_over_resident_meshes_with_sliding_zone_and_edits(seed):
    """Residency churn: a 7x7 render zone slides three times over generated terrain; each frame only
    the NEW row is meshed (the rest stays resident), rows that left are unloaded, lamps are placed by
    edits + vx_update_locations, and vx_visible_zone equals the oracle's frame over its `meshes`."""
    r = 3
    cx, cy = SPAWN, SPAWN
    side = 2 * r + 1
    span = side + 3 + 2  # zone + slides + ring
    wx0, wy0 = cx - r - 1, cy - r - 1
    world = oracle.World.generate(seed, wx0, wy0, span, side + 2)
    cams = [((8.0, 8.0, 60.0), (9.0, 8.5, 61.0)), ((40.0, -3.0, 45.0), (40.0, -3.0, 80.0)),
            ((20.0, 20.0, 70.0), (19.0, 20.0, 69.5))]
    with vg.Context(0, seed=seed) as ctx:
        ctx.set_mesh_retention(True)
        ctx.generate(world.area_xy(), read=False)
        meshed = set()
        mesh_calls = 0
        for step in range(4):
            zone = vg.render_zone((cx + step, cy), r)
            new = np.array([a for a in zone.tolist() if tuple(a) not in meshed], np.uint32).reshape(-1, 2)
            gone = [a for a in meshed if a not in {tuple(z) for z in zone.tolist()}]
            if gone:
                ctx.mesh_unload(np.array(gone, np.uint32))
                for a in gone:
                    world.face_mask[world.slot(*a)] = 0
                    meshed.discard(a)
            assert len(new) == (side * side if step == 0 else side), "only the new row is meshed"
            ctx.mesh(new)
            mesh_calls += len(new)
            for ax, ay in new:
                world.mesh_area(int(ax), int(ay))
                meshed.add((int(ax), int(ay)))
            # a lamp and a hole near the camera, through the incremental path
            lx, ly = (cx + step) * 16 + 5, cy * 16 + 7
            lz = int(world.max_height[world.slot(cx + step, cy)][5 + 16 * 7]) - 1
            edits = np.zeros(2, vg.EDIT_DTYPE)
            edits["gx"], edits["gy"], edits["gz"], edits["voxel"] = [lx, lx + 1], [ly, ly], [lz, lz + 1], [V["Lamp"], 0]
            for e in edits:
                world.apply_edit(int(e["gx"]), int(e["gy"]), int(e["gz"]), int(e["voxel"]))
            ctx.apply_edits(edits)
            ctx.update_locations(np.stack([edits["gx"], edits["gy"], edits["gz"]], axis=1), read=False)
            masks, flags_m = ctx.read_face_masks(world.area_xy())
            assert np.array_equal(masks, world.face_mask)
            assert {tuple(a) for a in world.area_xy()[flags_m != 0].tolist()} == meshed
            for cam, target in cams:
                for kw in [dict(), dict(show_map=True), dict(head_in_water=True)]:
                    pos = (cam[0] + 16.0 * step, cam[1], cam[2])
                    tgt = (target[0] + 16.0 * step, target[1], target[2])
                    out, codes, flags = ctx.visible_zone(zone, vg.make_view(pos, tgt, render_size=r, **kw))
                    rc, rf, ref, rl = _zone_reference(world, zone, meshed, oracle.make_view(pos, tgt, render_size=r, **kw))
                    assert np.array_equal(flags, rf)
                    assert np.array_equal(codes, rc), f"step {step} cam {cam} {kw}"
                    assert np.array_equal(out["type_first"], ref["type_first"]) and out["n_faces"] == ref["n_faces"]
                    assert out["n_lamps"] == len(rl) and np.array_equal(out["lamps"], rl[:64])
                    if kw.get("show_map"):
                        assert out["n_lamps"] == step + 1 and out["n_areas"] == len(meshed)
        assert mesh_calls == side * side + 3 * side  # zero full-zone re-meshes


def test_compact_records_expand_to_the_full_stream(ctx, seed):
    """vx_mesh_read_compact (4 B per visible voxel) carries everything: expanded records equal
    vx_mesh_read's, for generated terrain and for uploaded areas with all voxel types."""
    n = 6
    ax0, ay0 = SPAWN - 3, SPAWN - 3
    ctx.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
    xy = vg.region_xy(ax0, ay0, n, n)
    info = ctx.mesh(xy)
    rec_first, face_first, recs, verts, idx = ctx.mesh_read(info)
    rf2, packed = ctx.mesh_read_compact(info)
    assert np.array_equal(rf2, rec_first)
    assert vg.unpack_records(xy, rf2, packed).tobytes() == recs.tobytes()
    rng = np.random.default_rng(3)
    vox = rng.integers(0, 27, (9, 32768)).astype(np.uint8)
    vox[rng.random(vox.shape) < 0.6] = 0
    uxy = vg.region_xy(900, 900, 3, 3)
    ctx.upload(uxy, vox)
    one = np.array([[901, 901]], np.uint32)
    info = ctx.mesh(one)
    rec_first, _, recs, _, _ = ctx.mesh_read(info)
    rf2, packed = ctx.mesh_read_compact(info)
    assert vg.unpack_records(one, rf2, packed).tobytes() == recs.tobytes() and len(recs) > 5000


def test_mesh_after_column_heights_operator_uses_fresh_jobs(ctx, seed):
    """Regression: vx_column_heights overwrote the device job list that vx_mesh's
    same-list fast path trusted -- mesh, column_heights, mesh of the same list must agree."""
    n = 4
    ax0, ay0 = SPAWN - 2, SPAWN + 10
    ctx.generate(vg.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2), read=False)
    xy = vg.region_xy(ax0, ay0, n, n)
    a = ctx.mesh_read(ctx.mesh(xy))
    vox, mh = ctx.read_areas(xy)
    assert np.array_equal(ctx.column_heights(vox[:3]), mh[:3])
    b = ctx.mesh_read(ctx.mesh(xy))
    for u, v in zip(a, b):
        assert u.tobytes() == v.tobytes()


@pytest.mark.gpu
def test_deferred_reads_equal_blocking_reads(seed):
    """vx_set_deferred_reads: encode_read / mesh_read_compact return before their copies are done; after
    wait_reads the pinned buffers hold exactly what the blocking calls deliver -- also when the next step
    (another region) has been issued in between, which overwrites every device buffer the copies read."""
    xy_a = vg.region_xy(SPAWN - 6, SPAWN - 6, 12, 12)
    xy_b = vg.region_xy(SPAWN + 40, SPAWN - 6, 12, 12)
    with vg.Context(0, seed=seed) as c:
        want = {}
        for name, xy in (("a", xy_a), ("b", xy_b)):
            _, mh = c.generate(xy)
            total = c.encode_areas(xy, read=False)
            inf = c.mesh(xy[: len(xy) // 2])
            files, offsets = c.encode_read(len(xy), total)
            rf, packed = c.mesh_read_compact(inf)
            want[name] = (files.copy(), offsets.copy(), rf.copy(), packed.copy(), mh.copy())
        c.set_deferred_reads(True)
        bufs = {}
        for name, xy in (("a", xy_a), ("b", xy_b)):
            mh_pin = vg.PinnedArray((len(xy), 256), np.uint8)
            mh_pin.array[...] = 0xEE
            c.generate(xy, max_height_out=mh_pin.array, read=False)  # heights only: deferred as well
            total = c.encode_areas(xy, read=False)
            inf = c.mesh(xy[: len(xy) // 2])
            b = {"mh": mh_pin, "files": vg.PinnedArray((int(total),), np.uint8), "off": vg.PinnedArray((len(xy) + 1,), np.uint64),
                 "rf": vg.PinnedArray((len(xy) // 2 + 1,), np.uint32), "packed": vg.PinnedArray((inf["n_recs"],), np.uint32)}
            for k_, a in b.items():
                if k_ != "mh":
                    a.array.view(np.uint8)[...] = 0xEE
            c.encode_read(len(xy), total, out=b["files"].array, offsets_out=b["off"].array)
            c.mesh_read_compact(inf, out=(b["rf"].array, b["packed"].array))
            bufs[name] = b  # not waited for: step "b" is issued while the copies of "a" may still run
        c.wait_reads()
        for name in ("a", "b"):
            files, offsets, rf, packed, mh = want[name]
            b = bufs[name]
            assert np.array_equal(b["mh"].array, mh)
            assert np.array_equal(b["files"].array, files)
            assert np.array_equal(b["off"].array, offsets)
            assert np.array_equal(b["rf"].array, rf)
            assert np.array_equal(b["packed"].array, packed)
        # library-owned (pageable) buffers are complete on return even in this mode
        total = c.encode_areas(xy_a, read=False)
        files, offsets = c.encode_read(len(xy_a), total)
        assert np.array_equal(files, want["a"][0]) and np.array_equal(offsets, want["a"][1])
        c.set_deferred_reads(False)
        for a in [x for b in bufs.values() for x in b.values()]:
            a.free()
