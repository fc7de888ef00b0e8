"""The water and falling-voxel simulators as functions of ORDERED lists (oracle/physics.c), pinned
against cases worked out by hand from the reference's source (service/physics/water_simulator.rs,
falling_voxel_simulator.rs, utils.rs:7-13).  The reference has no test of its own for them.  CPU only."""
import numpy as np

import oracle

V = oracle.V
AX, AY = 500, 700  # a 3 x 3 grid of areas; everything happens in the middle one


def _world(fill=None):
    vox = np.zeros((9, 128, 16, 16), np.uint8)
    vox[:, 100:] = V["Stone"]  # floor: z >= 100 is solid
    if fill:
        fill(vox[4])
    vox = vox.reshape(9, 32768)
    mh = np.stack([oracle.update_all_column_heights(v).reshape(-1) for v in vox])
    return oracle.World(AX, AY, 3, 3, vox, mh)


def _g(x, y):
    return (AX + 1) * 16 + x, (AY + 1) * 16 + y


def _get(w, x, y, z):
    gx, gy = _g(x, y)
    return int(w.voxels[4].reshape(128, 16, 16)[z, y, x])


def test_source_on_a_floor_spreads_one_level_per_tick():
    """flow_sides (water_simulator.rs:110-150): a source on solid ground sets its four None sides to
    Water1 in the order x+1, x-1, y+1, y-1 and queues them; next tick each Water1 spreads Water2 to its
    None sides (not back into the source, level 100 >= 98)."""
    w = _world(lambda a: a.__setitem__((99, 8, 8), V["WaterSource"]))
    gx, gy = _g(8, 8)
    changed, nxt = w.water_tick([(gx, gy, 99)])
    want = [(gx + 1, gy, 99), (gx - 1, gy, 99), (gx, gy + 1, 99), (gx, gy - 1, 99)]
    assert changed.tolist() == [list(p) + [V["Water1"]] for p in want]
    assert nxt.tolist() == [list(p) for p in want]
    changed2, nxt2 = w.water_tick(nxt)
    # (x+1) first: its sides x+2, (x is the source: skipped), y+1, y-1
    assert changed2[:3].tolist() == [[gx + 2, gy, 99, V["Water2"]], [gx + 1, gy + 1, 99, V["Water2"]], [gx + 1, gy - 1, 99, V["Water2"]]]
    assert _get(w, 8, 8, 99) == V["WaterSource"] and _get(w, 10, 8, 99) == V["Water2"]
    # (x, y+1) comes third: its side (x+1, y+1) already holds Water2 from the first location (98 > 97: set
    # again to Water2 -- the reference does not compare with what it would write), so order matters
    assert [c for c in changed2.tolist() if c[:3] == [gx + 1, gy + 1, 99]] == [[gx + 1, gy + 1, 99, V["Water2"]]] * 2


def test_water4_does_not_spread_and_unsupported_water_dries_up():
    """flow_sides returns for LOWEST_WATER (:118-120); check_flow_stopped (:79-108): flowing water with no
    bordering water of a higher level becomes None, and location_updated queues the location, its four
    sides, z - 1 and (the reference's condition: z > 0) z + 1."""
    def fill(a):
        a[99, 3, 3] = V["Water4"]
        a[99, 3, 4] = V["Water3"]
    w = _world(fill)
    g4, g3 = _g(3, 3), _g(4, 3)
    changed, nxt = w.water_tick([(g4[0], g4[1], 99)])
    assert len(changed) == 0 and len(nxt) == 0  # Water4 next to Water3 (higher): stays, spreads nothing
    changed, nxt = w.water_tick([(g3[0], g3[1], 99)])  # Water3 has only a LOWER neighbour: it stops
    assert changed.tolist() == [[g3[0], g3[1], 99, V["None"]]]
    x, y = g3
    assert nxt.tolist() == [[x, y, 99], [x + 1, y, 99], [x - 1, y, 99], [x, y + 1, 99], [x, y - 1, 99], [x, y, 98], [x, y, 100]]
    changed, _ = w.water_tick([(g4[0], g4[1], 99)])  # now the Water4 is unsupported too
    assert changed.tolist() == [[g4[0], g4[1], 99, V["None"]]]


def test_flow_down_and_the_water_down_column():
    """flow_down (:184-206): over None (or flowing water) the voxel below becomes WaterDown and only it is
    queued; a WaterDown with water directly above is not stopped (:100-103), with nothing above it is."""
    w = _world(lambda a: a.__setitem__((90, 5, 5), V["WaterSource"]))
    gx, gy = _g(5, 5)
    changed, nxt = w.water_tick([(gx, gy, 90)])
    assert changed.tolist() == [[gx, gy, 91, V["WaterDown"]]] and nxt.tolist() == [[gx, gy, 91]]
    changed, nxt = w.water_tick(nxt)
    assert changed.tolist() == [[gx, gy, 92, V["WaterDown"]]]  # supported from above, flows on
    w.apply_edit(gx, gy, 90, V["None"])  # take the source away
    changed, nxt = w.water_tick([(gx, gy, 91)])
    assert changed.tolist() == [[gx, gy, 91, V["None"]]] and len(nxt) == 7
    # a source over a source does not flow down (down == WaterSource, :196-198) but sideways
    w2 = _world(lambda a: (a.__setitem__((98, 5, 5), V["WaterSource"]), a.__setitem__((99, 5, 5), V["WaterSource"])))
    changed, _ = w2.water_tick([(gx, gy, 98)])
    assert [c[3] for c in changed.tolist()] == [V["Water1"]] * 4 and all(c[2] == 98 for c in changed.tolist())
    # ... and water over flowing water does not spread sideways (NON_SOURCE_WATER below, :125-127)
    w3 = _world(lambda a: (a.__setitem__((98, 5, 5), V["Water1"]), a.__setitem__((98, 5, 6), V["WaterSource"]),
                           a.__setitem__((99, 5, 5), V["Water2"])))
    changed, _ = w3.water_tick([(gx, gy, 98)])
    assert changed.tolist() == [[gx, gy, 99, V["WaterDown"]]]  # flow_down wins: Water2 below is flowing water


def test_falling_chain_climbs_the_column():
    """update_voxels (falling_voxel_simulator.rs:105-150): a falling voxel over a non-solid voxel is
    removed and becomes an entity; the voxel above it is checked recursively, so a pillar of sand over a
    hole comes down top to bottom in one call, each level also scanning its own six neighbours."""
    def fill(a):
        a[95:99, 7, 7] = V["Sand"]   # pillar z = 95..98 standing on ...
        a[99, 7, 7] = V["None"]      # ... air (the floor starts at 100)
        a[98, 7, 8] = V["Dirt"]      # a neighbour of the pillar's foot, also unsupported
        a[97, 7, 6] = V["Grass"]     # one further up on the other side, resting on nothing either
        a[96, 7, 8] = V["Stone"]     # not a falling type
    w = _world(fill)
    gx, gy = _g(7, 7)
    spawned, nxt = w.falling_check([(gx, gy, 99)])  # the voxel below the pillar was just removed
    # to_check of z = 99: itself and its sides (None), then z - 1 = the foot: Sand over None -> falls, and
    # update_voxels recurses into the voxel ABOVE it at once (depth first): 97 falls -> 96 -> 95; on the
    # way back every level scans its own sides: at z = 96 the Stone stays, at z = 97 the Grass (over
    # None) falls.  The Dirt next to the foot is at z = 98, a level nobody scans (the recursion starts at
    # the voxel above the foot): it stays, although nothing holds it -- the reference's behaviour.
    got = [tuple(s) for s in spawned.tolist()]
    assert got == [(gx, gy, 98, V["Sand"]), (gx, gy, 97, V["Sand"]), (gx, gy, 96, V["Sand"]), (gx, gy, 95, V["Sand"]),
                   (gx - 1, gy, 97, V["Grass"])]
    assert _get(w, 8, 7, 98) == V["Dirt"]
    assert len(nxt) == 7 * len(got)
    for x, y, z, _ in got:
        assert w.voxels[4].reshape(128, 16, 16)[z, y - (AY + 1) * 16, x - (AX + 1) * 16] == V["None"]
    # nothing falls twice
    spawned, _ = w.falling_check([(gx, gy, 99)])
    assert len(spawned) == 0


def test_falling_land_places_on_top_of_solid_ground():
    """retain_or_place_voxel (:189-221) with vector_to_location (utils.rs:7-13: f32 round, z clamped):
    an entity whose centre + 0.5 rounds into a solid voxel is put on the voxel above it (if that is not
    solid) and dropped; in air or water it keeps falling; at z <= 0 it is dropped without a trace."""
    w = _world(lambda a: a.__setitem__((99, 2, 2), V["WaterSource"]))
    x0, y0 = (AX + 1) * 16 - 1_000_000, (AY + 1) * 16 - 1_000_000  # player coordinates of the area's corner
    pos = [(x0 + 4.2, y0 + 4.4, 50.0),    # in the air: keeps falling
           (x0 + 5.0, y0 + 5.0, 99.6),    # 100.1 rounds to 100 = floor: placed at z = 99
           (x0 + 5.4, y0 + 4.6, 99.7),    # same column (5.4 -> 5, 4.6 -> 5): z = 99 is now solid, 100.2 -> 100 solid, up (99) solid: dropped
           (x0 + 2.0, y0 + 2.0, 98.6),    # 99.1 -> 99 = water, not solid: keeps falling
           (x0 + 9.0, y0 + 9.0, 300.0)]   # clamped to z = 127, solid; up = 126 solid: dropped, nothing placed
    keep, placed, nxt = w.falling_land(pos, [V["Sand"], V["Dirt"], V["Snow"], V["Sand"], V["Grass"]])
    assert keep.tolist() == [1, 0, 0, 1, 0]
    gx, gy = _g(5, 5)
    assert placed.tolist() == [[gx, gy, 99, V["Dirt"]]]
    assert nxt.tolist()[:2] == [[gx, gy, 99], [gx + 1, gy, 99]] and len(nxt) == 7
    assert _get(w, 5, 5, 99) == V["Dirt"]
    # the column height follows (Area::set with an opaque voxel: min(prev, z))
    assert w.max_height[4][5 + 16 * 5] == 99
