/*
 * physics.c -- CPU ORACLE (test infrastructure, see vg_oracle.h) for the one physics routine that
 * is a batch of block edits with a data-independent visiting order: bomb explosions.
 *
 * Follows service/physics/bomb_simulator.rs:191-262 (handle_explosions, explode_at), :32-33
 * (EXPLOSION_RADIUS 4.5), model/voxel.rs:117-128 (is_solid).  Everything is integer / f32 code
 * that is on disk; the reference has no test for it.
 *
 * The water and falling-voxel simulators (water_simulator.rs:54-77, falling_voxel_simulator.rs:105-221)
 * update the world in place while they iterate: the water tick walks a HashSet, so its result depends
 * on std's randomised hash order.  They are restated here as functions of an ORDERED list -- the
 * caller passes the locations in the order its own iteration yields them, and the list is applied one
 * location after the other, exactly as the reference's loop body does.  Given the same order the
 * result is the reference's; no order is privileged.  The reference has no test for them.
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

/* ---------------------------------------------------------------- water (water_simulator.rs) */
static int is_water(uint8_t v) { return v >= VGO_WATER_SOURCE && v <= VGO_WATER4; }           /* Voxel::WATER */
static int is_non_source_water(uint8_t v) { return v >= VGO_WATER_DOWN && v <= VGO_WATER4; }  /* NON_SOURCE_WATER */

/* get_water_level (water_simulator.rs:170-181) */
static int water_level(uint8_t v) {
    switch (v) {
    case VGO_WATER_SOURCE: return 100;
    case VGO_WATER_DOWN: return 99;
    case VGO_WATER1: return 98;
    case VGO_WATER2: return 97;
    case VGO_WATER3: return 96;
    case VGO_WATER4: return 95;
    case VGO_NONE: return 0;
    default: return 1000;
    }
}

/* reduce_water_level (:152-168); only ever called with a water voxel above Water4 */
static uint8_t reduce_water_level(uint8_t v) {
    switch (v) {
    case VGO_WATER_SOURCE:
    case VGO_WATER_DOWN: return VGO_WATER1;
    case VGO_WATER1: return VGO_WATER2;
    case VGO_WATER2: return VGO_WATER3;
    case VGO_WATER3: return VGO_WATER4;
    default: return VGO_NONE;
    }
}

static int sim_get(const vgo_world* w, uint32_t x, uint32_t y, uint32_t z, uint8_t* out) {
    const long slot = slot_of(w, x / VGO_AREA_SIZE, y / VGO_AREA_SIZE);
    if (slot < 0 || z >= VGO_AREA_HEIGHT) return -1; /* the reference would load / generate the area (or panic on z) */
    *out = w->voxels[(size_t)slot * VGO_VOXELS_IN_AREA + x % VGO_AREA_SIZE + VGO_AREA_SIZE * (y % VGO_AREA_SIZE) + 256u * z];
    return 0;
}

static void push3(uint32_t* a, long* n, uint32_t x, uint32_t y, uint32_t z) {
    a[3 * *n] = x; a[3 * *n + 1] = y; a[3 * *n + 2] = z;
    ++*n;
}
static void push4(uint32_t* a, long* n, uint32_t x, uint32_t y, uint32_t z, uint32_t v) {
    a[4 * *n] = x; a[4 * *n + 1] = y; a[4 * *n + 2] = z; a[4 * *n + 3] = v;
    ++*n;
}

/* WaterSimulator::location_updated (:38-53): the location, its four sides, z - 1 if `z - 1 < AREA_HEIGHT`
 * (u32 arithmetic: false for z = 0, where the subtraction wraps), z + 1 if z > 0 -- the reference's own
 * conditions, kept as they are. */
static void water_location_updated(uint32_t* next, long* n, uint32_t x, uint32_t y, uint32_t z) {
    push3(next, n, x, y, z);
    push3(next, n, x + 1, y, z);
    push3(next, n, x - 1, y, z);
    push3(next, n, x, y + 1, z);
    push3(next, n, x, y - 1, z);
    if ((uint32_t)(z - 1u) < VGO_AREA_HEIGHT) push3(next, n, x, y, z - 1);
    if (z > 0) push3(next, n, x, y, z + 1);
}

/* WaterSimulator::simulate_voxels (:55-77) over `check` (n locations, in the caller's iteration order).
 * changed: x, y, z, new voxel of every world.set + renderer.update_location, in order (room for 4 per
 * location); next: the locations inserted into check_locations for the next tick, in insertion order,
 * duplicates included (room for 7 per location).  Returns 0, or -1 if a voxel outside the grid is needed. */
int vgo_water_tick(vgo_world* w, const uint32_t* check, int n, uint32_t* changed, long* n_changed, uint32_t* next,
                   long* n_next) {
    long nc = 0, nn = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t x = check[3 * i], y = check[3 * i + 1], z = check[3 * i + 2];
        uint8_t voxel;
        if (sim_get(w, x, y, z, &voxel)) return -1;
        if (!is_water(voxel)) continue;
        /* get_bordering (:209-238): sides first, then up (z - 1) if z > 0, then down (z + 1) if it exists */
        uint32_t bz[6];
        uint8_t bv[6];
        const uint32_t bx[4] = {x + 1, x - 1, x, x}, by[4] = {y, y, y + 1, y - 1};
        int nb = 0, has_down = 0;
        uint8_t down = VGO_NONE;
        for (int k = 0; k < 4; ++k, ++nb) {
            if (sim_get(w, bx[k], by[k], z, &bv[nb])) return -1;
            bz[nb] = z;
        }
        if (z > 0) {
            if (sim_get(w, x, y, z - 1, &bv[nb])) return -1;
            bz[nb++] = z - 1;
        }
        if (z + 1 < VGO_AREA_HEIGHT) {
            if (sim_get(w, x, y, z + 1, &bv[nb])) return -1;
            bz[nb++] = z + 1;
            down = bv[nb - 1];
            has_down = 1;
        }
        /* check_flow_stopped (:79-108) */
        int stopped = voxel != VGO_WATER_SOURCE;
        if (stopped) {
            const int level = water_level(voxel);
            for (int k = 0; k < nb; ++k) {
                if (!is_water(bv[k]) || bz[k] > z) continue;
                if (water_level(bv[k]) > level || (voxel == VGO_WATER_DOWN && bz[k] != z)) {
                    stopped = 0;
                    break;
                }
            }
        }
        if (stopped) {
            vgo_apply_edit(w, x, y, z, VGO_NONE);
            push4(changed, &nc, x, y, z, VGO_NONE);
            water_location_updated(next, &nn, x, y, z);
            continue;
        }
        /* flow_down (:184-206) */
        if (z + 1 < VGO_AREA_HEIGHT && has_down && (down == VGO_NONE || (down != VGO_WATER_SOURCE && is_water(down)))) {
            vgo_apply_edit(w, x, y, z + 1, VGO_WATER_DOWN);
            push4(changed, &nc, x, y, z + 1, VGO_WATER_DOWN);
            push3(next, &nn, x, y, z + 1);
            continue;
        }
        /* flow_sides (:110-150) with the bordering voxels as they were read above */
        if (voxel == VGO_WATER4 || !has_down || is_non_source_water(down)) continue;
        const int level = water_level(voxel);
        for (int k = 0; k < 4; ++k) {
            if (level <= water_level(bv[k])) continue;
            const uint8_t to_set = reduce_water_level(voxel);
            vgo_apply_edit(w, bx[k], by[k], z, to_set);
            push4(changed, &nc, bx[k], by[k], z, to_set);
            push3(next, &nn, bx[k], by[k], z);
        }
    }
    *n_changed = nc;
    *n_next = nn;
    return 0;
}

/* ---------------------------------------------------------------- falling voxels (falling_voxel_simulator.rs) */
static int is_falling(uint8_t v) { return v == VGO_SAND || v == VGO_DIRT || v == VGO_GRASS || v == VGO_SNOW; } /* Voxel::FALLING */

typedef struct {
    vgo_world* w;
    uint32_t *spawned, *water_next;
    long ns, nw;
    int err;
} falling_ctx;

/* FallingVoxelSimulator::update_voxels (:105-150), recursion and all */
static void falling_update_voxels(falling_ctx* c, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t tx[7], ty[7], tz[7];
    int n = 0;
    /* get_voxels_to_check (:69-103) */
    tx[n] = x; ty[n] = y; tz[n++] = z;
    tx[n] = x + 1; ty[n] = y; tz[n++] = z;
    tx[n] = x - 1; ty[n] = y; tz[n++] = z;
    tx[n] = x; ty[n] = y + 1; tz[n++] = z;
    tx[n] = x; ty[n] = y - 1; tz[n++] = z;
    if (z > 0) { tx[n] = x; ty[n] = y; tz[n++] = z - 1; }
    if (z + 1 < VGO_AREA_HEIGHT) { tx[n] = x; ty[n] = y; tz[n++] = z + 1; }
    for (int k = 0; k < n && !c->err; ++k) {
        uint8_t voxel, lower;
        if (sim_get(c->w, tx[k], ty[k], tz[k], &voxel)) { c->err = 1; return; }
        if (!is_falling(voxel) || tz[k] + 1 >= VGO_AREA_HEIGHT) continue;
        if (sim_get(c->w, tx[k], ty[k], tz[k] + 1, &lower)) { c->err = 1; return; }
        if (is_solid(lower)) continue;
        vgo_apply_edit(c->w, tx[k], ty[k], tz[k], VGO_NONE);
        water_location_updated(c->water_next, &c->nw, tx[k], ty[k], tz[k]);
        push4(c->spawned, &c->ns, tx[k], ty[k], tz[k], voxel); /* SimulatedVoxel at the location, velocity 0 */
        if (tz[k] >= 1) {
            uint8_t up;
            if (sim_get(c->w, tx[k], ty[k], tz[k] - 1, &up)) { c->err = 1; return; }
            if (is_falling(up)) falling_update_voxels(c, tx[k], ty[k], tz[k] - 1);
        }
    }
}

/* update_voxels for each of `check` (n internal locations) in order.  spawned: x, y, z, voxel of every
 * falling entity created (= every voxel removed), in creation order; water_next: what
 * water_simulator.location_updated inserted.  The caller sizes both (a chain can empty whole columns).
 * Returns 0, -1 outside the grid, -2 if an output would overflow its capacity (in entries). */
int vgo_falling_check(vgo_world* w, const uint32_t* check, int n, uint32_t* spawned, long cap_spawned, long* n_spawned,
                      uint32_t* water_next, long cap_next, long* n_next) {
    falling_ctx c = {w, spawned, water_next, 0, 0, 0};
    for (int i = 0; i < n && !c.err; ++i) {
        if (c.ns + 1024 > cap_spawned || c.nw + 7 * 1024 > cap_next) return -2;
        falling_update_voxels(&c, check[3 * i], check[3 * i + 1], check[3 * i + 2]);
    }
    *n_spawned = c.ns;
    *n_next = c.nw;
    return c.err ? -1 : 0;
}

/* simulate_falling's retain pass (:152-166, retain_or_place_voxel :189-221) over the entities in Vec
 * order, AFTER the caller has advanced their positions: positions (3 f32 each, player coordinates),
 * voxel types.  keep[i] = the entity keeps falling; placed: x, y, z, voxel of every voxel put down. */
int vgo_falling_land(vgo_world* w, const float* positions, const uint8_t* types, int n, uint8_t* keep, uint32_t* placed,
                     long* n_placed, uint32_t* water_next, long* n_next) {
    long np = 0, nw = 0;
    for (int i = 0; i < n; ++i) {
        /* vector_to_location(position + vec3(0, 0, HALF_SIZE)) (utils.rs:7-13): f32 round, z clamped */
        const float px = positions[3 * i], py = positions[3 * i + 1], pz = positions[3 * i + 2] + HALF_SIZE;
        const int lx = (int)roundf(px), ly = (int)roundf(py);
        int lz = (int)roundf(pz);
        lz = lz < 0 ? 0 : (lz > VGO_AREA_HEIGHT - 1 ? VGO_AREA_HEIGHT - 1 : lz);
        const uint32_t gx = (uint32_t)(lx + VGO_LOCATION_OFFSET), gy = (uint32_t)(ly + VGO_LOCATION_OFFSET);
        uint8_t here, up;
        if (sim_get(w, gx, gy, (uint32_t)lz, &here)) return -1;
        if (!is_solid(here)) { keep[i] = 1; continue; }
        keep[i] = 0;
        if (lz <= 0) continue;
        if (sim_get(w, gx, gy, (uint32_t)lz - 1, &up)) return -1;
        if (!is_solid(up)) {
            vgo_apply_edit(w, gx, gy, (uint32_t)lz - 1, types[i]);
            push4(placed, &np, gx, gy, (uint32_t)lz - 1, types[i]);
            water_location_updated(water_next, &nw, gx, gy, (uint32_t)lz - 1);
        }
    }
    *n_placed = np;
    *n_next = nw;
    return 0;
}
