/*
 * physics.c -- CPU ORACLE (test infrastructure, see vg_oracle.h) for the one physics routine that
 * is a batch of block edits with a data-independent visiting order: bomb explosions.
 *
 * Follows service/physics/bomb_simulator.rs:191-262 (handle_explosions, explode_at), :32-33
 * (EXPLOSION_RADIUS 4.5), model/voxel.rs:117-128 (is_solid).  Everything is integer / f32 code
 * that is on disk; the reference has no test for it.
 *
 * Not restated: the water and falling-voxel simulators (water_simulator.rs:54-77,
 * falling_voxel_simulator.rs:105-150).  They update the world in place while iterating a HashSet /
 * recursing, so their result depends on std's randomised hash order -- there is no single answer
 * an oracle could pin.
 */
#include <math.h>

#include "vg_oracle.h"

#define EXPLOSION_RADIUS 4.5f
#define EXPLOSION_RADIUS_SQ (EXPLOSION_RADIUS * EXPLOSION_RADIUS)
#define HALF_SIZE 0.5f

static int is_solid(uint8_t v) { return !(v == VGO_NONE || (v >= VGO_WATER_SOURCE && v <= VGO_WATER4)); }

static long slot_of(const vgo_world* w, uint32_t ax, uint32_t ay) {
    const uint32_t ix = ax - w->ax0, iy = ay - w->ay0;
    return (ix < w->nx && iy < w->ny) ? (long)ix + (long)w->nx * iy : -1;
}

/* positions: n Vec3 (player coordinates, f32).  Processed from the last to the first
 * (explosion_positions.pop(), bomb_simulator.rs:199-207).  removed / bombs: internal coordinates
 * (x, y, z triples) in visiting order; returns 0, or -1 if a visited voxel is in an area outside the
 * grid (the reference would load / generate it). */
int vgo_explode(vgo_world* w, const float* positions, int n, uint32_t* removed, long* n_removed,
                uint32_t* bombs, long* n_bombs) {
    long nr = 0, nb = 0;
    for (int e = n - 1; e >= 0; --e) {
        const float px = positions[3 * e], py = positions[3 * e + 1], pz = positions[3 * e + 2];
        const int cx = (int)floorf(px), cy = (int)floorf(py), cz = (int)floorf(pz);
        const int r = (int)ceilf(EXPLOSION_RADIUS);
        const int z_min = cz - r > 0 ? cz - r : 0;
        const int z_max = cz + r < VGO_AREA_HEIGHT - 2 ? cz + r : VGO_AREA_HEIGHT - 2;
        for (int x = cx - r; x <= cx + r; ++x)
            for (int y = cy - r; y <= cy + r; ++y)
                for (int z = z_min; z <= z_max; ++z) {
                    const float dx = (float)x + HALF_SIZE - px, dy = (float)y + HALF_SIZE - py,
                                dz = (float)z + HALF_SIZE - pz;
                    if (dx * dx + dy * dy + dz * dz > EXPLOSION_RADIUS_SQ) continue;
                    const uint32_t gx = (uint32_t)(x + VGO_LOCATION_OFFSET), gy = (uint32_t)(y + VGO_LOCATION_OFFSET);
                    const long slot = slot_of(w, gx / VGO_AREA_SIZE, gy / VGO_AREA_SIZE);
                    if (slot < 0) return -1;
                    const uint8_t v = w->voxels[(size_t)slot * VGO_VOXELS_IN_AREA + gx % VGO_AREA_SIZE +
                                                VGO_AREA_SIZE * (gy % VGO_AREA_SIZE) + 256u * (uint32_t)z];
                    if (v == VGO_BOMB) { /* add_active_bomb_from_explosion (:266-277) */
                        bombs[3 * nb] = gx; bombs[3 * nb + 1] = gy; bombs[3 * nb + 2] = (uint32_t)z;
                        ++nb;
                    }
                    if (is_solid(v)) {
                        vgo_apply_edit(w, gx, gy, (uint32_t)z, VGO_NONE); /* world.set + update_location (:209-211) */
                        removed[3 * nr] = gx; removed[3 * nr + 1] = gy; removed[3 * nr + 2] = (uint32_t)z;
                        ++nr;
                    }
                }
    }
    *n_removed = nr;
    *n_bombs = nb;
    return 0;
}
