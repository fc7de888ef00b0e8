"""Oracle for the per-frame visibility pass (oracle/frame.c) against hand-computed cases of
renderer.rs:324-346 (is_area_visible), :549-558 (calculate_area_render_distance) and :560-586
(is_voxel_visible).  The reference has no tests for these functions; the expectations below are
worked out from its source with f32 arithmetic."""
import numpy as np

import oracle

OFF = 1_000_000
f32 = np.float32


def test_area_render_distance_formula():
    # looking straight down: dot = 1 -> smooth = 1 -> 16 * render_size * 0.5
    assert oracle.area_render_distance([0, 0, 1], 8) == 64.0
    assert oracle.area_render_distance([0, 0, -1], 8) == 64.0  # abs()
    # horizontal look: dot = 0 -> 0 -> clamped to one area
    assert oracle.area_render_distance([1, 0, 0], 8) == 16.0
    # in between, in f32: 1 - (1 - d)^2, times 128, times 0.5
    d = f32(0.6)
    smooth = f32(1) - (f32(1) - d) * (f32(1) - d)
    expect = max(float(smooth * f32(128) * f32(0.5)), 16.0)
    assert oracle.area_render_distance([0.8, 0, 0.6], 8) == expect


def test_voxel_proximity_box_ignores_look_direction():
    cam = [0.0, 0.0, 60.0]
    look = [1.0, 0.0, 0.0]
    # 5.5 behind the camera on every axis: still inside the proximity box
    assert oracle.is_voxel_visible(OFF - 5, OFF + 5, 65, look, cam, 128.0)
    assert oracle.is_voxel_visible(OFF - 5, OFF, 60, look, [0.5, 0.0, 60.0], 128.0)  # |dx| = 5.5 <= 5.5
    assert not oracle.is_voxel_visible(OFF - 6, OFF, 60, look, cam, 128.0)  # behind and outside the box


def test_voxel_distance_and_cone():
    cam = [0.0, 0.0, 60.0]
    look = [1.0, 0.0, 0.0]
    assert oracle.is_voxel_visible(OFF + 100, OFF, 60, look, cam, 128.0)       # straight ahead
    assert not oracle.is_voxel_visible(OFF + 129, OFF, 60, look, cam, 128.0)   # beyond render distance
    assert oracle.is_voxel_visible(OFF + 128, OFF, 60, look, cam, 128.0)       # == max distance is kept
    # cone: dot > 0.71 * distance.  (50, 49): cos = 50 / 70.007 = 0.7142 > 0.71; (50, 50): 0.7071 < 0.71
    assert oracle.is_voxel_visible(OFF + 50, OFF + 49, 60, look, cam, 128.0)
    assert not oracle.is_voxel_visible(OFF + 50, OFF + 50, 60, look, cam, 128.0)


def test_area_visibility_distance_or_cone():
    v = oracle.make_view([8.0, 8.0, 60.0], [9.0, 8.0, 60.0], render_size=8)  # look = +x
    rd = oracle.area_render_distance(v.look, 8)
    assert rd == 16.0
    spawn = OFF // 16  # area 62500 has its middle at (8, 8)
    assert oracle.is_area_visible(spawn, spawn, v, rd)          # distance 0
    assert oracle.is_area_visible(spawn + 1, spawn, v, rd)      # 16 away: <= 16
    assert oracle.is_area_visible(spawn + 5, spawn + 1, v, rd)  # ahead, inside the 0.35 cone
    assert not oracle.is_area_visible(spawn - 2, spawn, v, rd)  # behind
    # cone boundary: direction (1, k): cos = 1 / sqrt(1 + k^2) >= 0.35 <=> k <= 2.676
    assert oracle.is_area_visible(spawn + 2, spawn + 5, v, rd)      # k = 2.5
    assert not oracle.is_area_visible(spawn + 2, spawn + 6, v, rd)  # k = 3


def test_make_view_normalizes_like_glam():
    v = oracle.make_view([1, 2, 3], [1, 2, 3])  # zero vector -> normalize_or_zero gives zero
    assert list(v.look) == [0.0, 0.0, 0.0]
    v = oracle.make_view([0, 0, 0], [3, 0, 4])
    assert np.allclose(list(v.look), [0.6, 0.0, 0.8], atol=1e-7)


def _records(points):
    """points: (gx, gy, gz, voxel, nfaces) -> REC_DTYPE array."""
    r = np.zeros(len(points), oracle.REC_DTYPE)
    for i, (gx, gy, gz, voxel, nf) in enumerate(points):
        r[i] = (gx, gy, gz | (voxel << 8) | (((1 << nf) - 1) << 16) | (nf << 24), 0)
    return r


def test_filter_groups_by_type_and_collects_lamps():
    V = oracle.V
    spawn = OFF // 16
    recs = _records([
        (OFF + 20, OFF + 8, 60, V["Stone"], 2),        # ahead: drawn
        (OFF + 21, OFF + 8, 60, V["Grass"], 1),        # ahead: drawn
        (OFF + 22, OFF + 8, 60, V["WaterSource"], 3),  # ahead: drawn unless the head is in water
        (OFF + 23, OFF + 8, 60, V["Lamp"], 6),         # ahead: drawn, and a light
        (OFF - 40, OFF + 8, 60, V["Lamp"], 6),         # second area, behind the camera: area not visible
        (OFF + 24, OFF + 8, 60, V["Grass"], 1),
    ])
    # area 0 = spawn+1 (x 16..31) holds records 0..3 and 5 -> reorder so that each area's records are contiguous
    recs = recs[[0, 1, 2, 3, 5, 4]]
    rec_first = np.array([0, 5, 6], np.uint32)
    area_xy = np.array([[spawn + 1, spawn], [spawn - 3, spawn]], np.uint32)
    view = oracle.make_view([8.0, 8.0, 60.0], [9.0, 8.0, 60.0], render_size=8)
    out = oracle.filter_visible(recs, rec_first, area_xy, view)
    assert list(out["area_visible"]) == [1, 0]
    # groups in voxel-id order: Grass (3) [1, 4], Stone (9) [0], Lamp (11) [3], WaterSource (14) [2]
    assert list(out["visible"]) == [1, 4, 0, 3, 2]
    tf = out["type_first"]
    assert tf[V["Grass"]] == 0 and tf[V["Grass"] + 1] == 2 and tf[V["Stone"]] == 2 and tf[27] == 5
    assert out["n_faces"] == 1 + 1 + 2 + 6 + 3
    assert list(out["lamps"]) == [3]  # only the lamp of the visible area
    wet = oracle.make_view([8.0, 8.0, 60.0], [9.0, 8.0, 60.0], render_size=8, head_in_water=True)
    assert list(oracle.filter_visible(recs, rec_first, area_xy, wet)["visible"]) == [1, 4, 0, 3]
    everything = oracle.make_view([8.0, 8.0, 60.0], [9.0, 8.0, 60.0], render_size=8, show_map=True)
    out = oracle.filter_visible(recs, rec_first, area_xy, everything)
    assert list(out["area_visible"]) == [1, 1] and len(out["visible"]) == 6 and list(out["lamps"]) == [3, 5]
