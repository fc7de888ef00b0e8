// vx_sim.cu -- the world-mutating loops of the water and falling-voxel simulators on the device
// (SURVEY.md 8 f-4), so that a device-resident world needs no host round trip per touched voxel.
//
//   water_tick     WaterSimulator::simulate_voxels (service/physics/water_simulator.rs:55-77) with
//                  check_flow_stopped (:79-108), flow_down (:184-206), flow_sides (:110-150),
//                  get_bordering (:209-238), location_updated (:38-53)
//   falling_check  FallingVoxelSimulator::update_voxels (falling_voxel_simulator.rs:105-150,
//                  get_voxels_to_check :69-103), recursion included
//   falling_land   the retain pass of simulate_falling (:152-166, retain_or_place_voxel :189-221)
//
// These loops mutate the world WHILE they iterate, the water tick over a HashSet: the outcome depends
// on the iteration order, which std randomises per process.  So the order is an INPUT here: the caller
// hands over its locations in the order its own iteration yields them and they are applied one after
// the other, exactly as the reference's loop body does -- given the same order, the same world (the
// water tick runs as one warp whose lanes fetch a location's seven voxels side by side; the falling
// loops as one thread).  That makes these kernels latency-bound by construction (a dependent memory
// round trip per location, microseconds each); what they buy is residency, not throughput: the edits land in the
// device's voxels, planes and column heights, and the list of changed locations goes straight into
// vx_update_locations (the reference calls renderer.update_location after every set).
#include "vx_kernels.hpp"

namespace vx {
namespace {

struct SimWorld {
    AreaMap map;
    uint8_t* voxels;
    uint8_t* heights;
    uint16_t* planes;
    int32_t* status;  // 0, or a negative vx_status
    // one-entry cache of the last area looked up: neighbouring voxels mostly share it
    uint32_t last_ax, last_ay;
    int last_slot;
};

__device__ __forceinline__ int sim_slot(SimWorld& w, uint32_t gx, uint32_t gy) {
    const uint32_t ax = gx / AREA_SIZE, ay = gy / AREA_SIZE;
    if (w.last_slot >= 0 && ax == w.last_ax && ay == w.last_ay) return w.last_slot;
    const int s = area_slot(w.map, ax, ay);
    if (s >= 0) {
        w.last_ax = ax;
        w.last_ay = ay;
        w.last_slot = s;
    }
    return s;
}

// World::get; false (and the status set) where the reference would load / generate an area, or index
// past the top of an area
__device__ __forceinline__ bool sim_get(SimWorld& w, uint32_t gx, uint32_t gy, uint32_t gz, uint32_t& v) {
    const int s = sim_slot(w, gx, gy);
    if (s < 0 || gz >= (uint32_t)AREA_HEIGHT) {
        *w.status = s < 0 ? -5 : -3;  // VX_ERR_NOT_LOADED / VX_ERR_INVALID
        return false;
    }
    v = w.voxels[(size_t)s * VOXELS_IN_AREA + (gx % AREA_SIZE) + (gy % AREA_SIZE) * AREA_SIZE + gz * 256u];
    return true;
}

// World::set (world.rs:184-191) -> Area::set (area.rs:99-111): the voxel, the column height (rescan if
// the new voxel is transparent, else min(prev, z)) and the S / T bit planes the mesher reads.
__device__ __forceinline__ void sim_set(SimWorld& w, uint32_t gx, uint32_t gy, uint32_t gz, uint32_t v) {
    const int s = sim_slot(w, gx, gy);  // resident: the location was read before it is written
    const uint32_t col = (gx % AREA_SIZE) + (gy % AREA_SIZE) * AREA_SIZE, local = col + gz * 256u;
    uint8_t* vox = w.voxels + (size_t)s * VOXELS_IN_AREA;
    vox[local] = (uint8_t)v;
    uint8_t* h = w.heights + (size_t)s * COLUMNS_IN_AREA + col;
    if (is_transparent(v)) {
        int z = 0;
        while (z < AREA_HEIGHT && is_transparent(vox[col + 256 * z])) ++z;
        *h = (uint8_t)(z < AREA_HEIGHT ? z : AREA_HEIGHT - 1);
    } else if (gz < *h) {
        *h = (uint8_t)gz;
    }
    const uint32_t row = local >> 4, bit = 1u << ((local & 15u) + 16u * (row & 1u));
    uint32_t* S = reinterpret_cast<uint32_t*>(w.planes + (size_t)s * (VOXELS_IN_AREA / 8)) + (row >> 1);
    uint32_t* T = S + VOXELS_IN_AREA / 32;
    *S = v != 0u ? (*S | bit) : (*S & ~bit);
    *T = is_transparent(v) ? (*T | bit) : (*T & ~bit);
}

struct List {
    uint32_t* p;
    uint32_t n, cap;  // entries
    int32_t* status;
};
__device__ __forceinline__ void push3(List& l, uint32_t x, uint32_t y, uint32_t z) {
    if (l.n >= l.cap) {
        *l.status = -3;  // VX_ERR_INVALID: the caller's buffer is too small
        return;
    }
    l.p[3 * l.n] = x;
    l.p[3 * l.n + 1] = y;
    l.p[3 * l.n + 2] = z;
    ++l.n;
}
__device__ __forceinline__ void push4(List& l, uint32_t x, uint32_t y, uint32_t z, uint32_t v) {
    if (l.n >= l.cap) {
        *l.status = -3;
        return;
    }
    reinterpret_cast<uint4*>(l.p)[l.n] = make_uint4(x, y, z, v);
    ++l.n;
}

// WaterSimulator::location_updated (water_simulator.rs:38-53), the reference's own conditions:
// z - 1 if `z - 1 < AREA_HEIGHT` in u32 arithmetic (false for z = 0), z + 1 if z > 0
__device__ __forceinline__ void water_location_updated(List& next, uint32_t x, uint32_t y, uint32_t z) {
    push3(next, x, y, z);
    push3(next, x + 1, y, z);
    push3(next, x - 1, y, z);
    push3(next, x, y + 1, z);
    push3(next, x, y - 1, z);
    if (z - 1u < (uint32_t)AREA_HEIGHT) push3(next, x, y, z - 1u);
    if (z > 0u) push3(next, x, y, z + 1u);
}

__device__ __forceinline__ bool is_water(uint32_t v) { return v >= V_WATER_SOURCE && v <= V_WATER4; }
__device__ __forceinline__ bool is_solid(uint32_t v) { return !(v == V_NONE || is_water(v)); }  // voxel.rs:117-128
__device__ __forceinline__ bool is_falling(uint32_t v) { return v == V_SAND || v == V_DIRT || v == V_GRASS || v == V_SNOW; }

// get_water_level (water_simulator.rs:170-181)
__device__ __forceinline__ int water_level(uint32_t v) {
    if (is_water(v)) return 100 - (int)(v - V_WATER_SOURCE);  // Source 100, Down 99, Water1..4 98..95
    return v == V_NONE ? 0 : 1000;
}

// One warp: lane k < 7 fetches the k-th voxel of the location's neighbourhood (itself, the four sides,
// up, down), so a location costs one round trip to memory instead of seven in a row; every lane then
// runs the (uniform) decision, lane 0 writes.  Still one location after the other.
__global__ void __launch_bounds__(32) water_tick_kernel(const uint32_t* __restrict__ check, int n, SimWorld w, List changed,
                                                        List next, uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xFFFFFFFFu;
    for (int i = 0; i < n; ++i) {
        const uint32_t x = check[3 * i], y = check[3 * i + 1], z = check[3 * i + 2];
        // get_bordering's order behind the location itself: x + 1, x - 1, y + 1, y - 1, z - 1 (if z > 0), z + 1 (if it exists)
        uint32_t mx = x, my = y, mz = z;
        bool exists = lane < 7;
        if (lane == 1) mx += 1u;
        else if (lane == 2) mx -= 1u;
        else if (lane == 3) my += 1u;
        else if (lane == 4) my -= 1u;
        else if (lane == 5) { exists = z > 0u; mz -= 1u; }
        else if (lane == 6) { exists = z + 1u < (uint32_t)AREA_HEIGHT; mz += 1u; }
        uint32_t mine = V_NONE;
        bool ok = true;
        if (exists) ok = sim_get(w, mx, my, mz, mine);
        if (__any_sync(FULL, !ok)) break;  // the status is set
        const uint32_t voxel = __shfl_sync(FULL, mine, 0);
        if (!is_water(voxel)) continue;
        uint32_t bv[6], bz[6];
        int nb = 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bv[k] = __shfl_sync(FULL, mine, k + 1);
            bz[k] = z;
        }
        const uint32_t up = __shfl_sync(FULL, mine, 5), down = __shfl_sync(FULL, mine, 6);
        const bool has_down = z + 1u < (uint32_t)AREA_HEIGHT;
        if (z > 0u) {
            bv[nb] = up;
            bz[nb++] = z - 1u;
        }
        if (has_down) {
            bv[nb] = down;
            bz[nb++] = z + 1u;
        }
        const uint32_t bx[4] = {x + 1, x - 1, x, x}, by[4] = {y, y, y + 1, y - 1};
        // check_flow_stopped
        bool stopped = voxel != V_WATER_SOURCE;
        if (stopped) {
            const int level = water_level(voxel);
            for (int k = 0; k < nb; ++k) {
                if (!is_water(bv[k]) || bz[k] > z) continue;
                if (water_level(bv[k]) > level || (voxel == V_WATER_DOWN && bz[k] != z)) {
                    stopped = false;
                    break;
                }
            }
        }
        if (lane == 0) {
            if (stopped) {
                sim_set(w, x, y, z, V_NONE);
                push4(changed, x, y, z, V_NONE);
                water_location_updated(next, x, y, z);
            } else if (has_down && (down == V_NONE || (down != V_WATER_SOURCE && is_water(down)))) {  // flow_down
                sim_set(w, x, y, z + 1u, V_WATER_DOWN);
                push4(changed, x, y, z + 1u, V_WATER_DOWN);
                push3(next, x, y, z + 1u);
            } else if (!(voxel == V_WATER4 || !has_down || (is_water(down) && down != V_WATER_SOURCE))) {
                // flow_sides, on the bordering voxels as they were read above
                const int level = water_level(voxel);
                const uint32_t lower = voxel <= V_WATER_DOWN ? (uint32_t)V_WATER1 : voxel + 1u;  // reduce_water_level
                for (int k = 0; k < 4; ++k) {
                    if (level <= water_level(bv[k])) continue;
                    sim_set(w, bx[k], by[k], z, lower);
                    push4(changed, bx[k], by[k], z, lower);
                    push3(next, bx[k], by[k], z);
                }
            }
        }
        __syncwarp();  // lane 0's writes are visible to the lanes that read the next location
        if (__shfl_sync(FULL, (int)(*changed.status != 0 && lane == 0), 0)) break;  // an output list overflowed
    }
    if (lane == 0) {
        counts[0] = changed.n;
        counts[1] = next.n;
    }
}

// update_voxels recurses into the voxel above every voxel it drops; the recursion sits inside the loop
// over the seven locations, so a frame is (location, position in that loop).  A chain climbs one
// column, at most AREA_HEIGHT frames.
__global__ void falling_check_kernel(const uint32_t* __restrict__ check, int n, SimWorld w, List spawned, List next,
                                     uint32_t* __restrict__ counts) {
    struct Frame { uint32_t x, y, z, k; };
    Frame stack[AREA_HEIGHT + 2];
    for (int i = 0; i < n && *w.status == 0; ++i) {
        int depth = 0;
        stack[0] = Frame{check[3 * i], check[3 * i + 1], check[3 * i + 2], 0u};
        while (depth >= 0 && *w.status == 0) {
            Frame& f = stack[depth];
            if (f.k >= 7u) {
                --depth;
                continue;
            }
            const uint32_t k = f.k++;
            // get_voxels_to_check: the location, x + 1, x - 1, y + 1, y - 1, z - 1 (if z > 0), z + 1 (if it exists)
            uint32_t tx = f.x, ty = f.y, tz = f.z;
            if (k == 1u) tx += 1u;
            else if (k == 2u) tx -= 1u;
            else if (k == 3u) ty += 1u;
            else if (k == 4u) ty -= 1u;
            else if (k == 5u) {
                if (f.z == 0u) continue;
                tz -= 1u;
            } else if (k == 6u) {
                if (f.z + 1u >= (uint32_t)AREA_HEIGHT) continue;
                tz += 1u;
            }
            uint32_t voxel, lower;
            if (!sim_get(w, tx, ty, tz, voxel)) break;
            if (!is_falling(voxel) || tz + 1u >= (uint32_t)AREA_HEIGHT) continue;
            if (!sim_get(w, tx, ty, tz + 1u, lower)) break;
            if (is_solid(lower)) continue;
            sim_set(w, tx, ty, tz, V_NONE);
            water_location_updated(next, tx, ty, tz);
            push4(spawned, tx, ty, tz, voxel);
            if (tz >= 1u) {
                uint32_t up;
                if (!sim_get(w, tx, ty, tz - 1u, up)) break;
                if (is_falling(up)) {
                    if (depth + 1 >= AREA_HEIGHT + 2) {  // cannot happen: a chain climbs one column
                        *w.status = -2;
                        break;
                    }
                    stack[++depth] = Frame{tx, ty, tz - 1u, 0u};
                }
            }
        }
    }
    counts[0] = spawned.n;
    counts[1] = next.n;
}

__global__ void falling_land_kernel(const float* __restrict__ positions, const uint8_t* __restrict__ types, int n,
                                    SimWorld w, uint8_t* __restrict__ keep, List placed, List next,
                                    uint32_t* __restrict__ counts) {
    for (int i = 0; i < n && *w.status == 0; ++i) {
        // vector_to_location(position + vec3(0, 0, HALF_SIZE)) (utils.rs:7-13): f32 round half away from
        // zero, z clamped to the area
        const float px = positions[3 * i], py = positions[3 * i + 1], pz = positions[3 * i + 2] + 0.5f;
        const int lx = (int)roundf(px), ly = (int)roundf(py);
        int lz = (int)roundf(pz);
        lz = lz < 0 ? 0 : (lz > AREA_HEIGHT - 1 ? AREA_HEIGHT - 1 : lz);
        const uint32_t gx = (uint32_t)(lx + LOCATION_OFFSET), gy = (uint32_t)(ly + LOCATION_OFFSET);
        uint32_t here, up;
        if (!sim_get(w, gx, gy, (uint32_t)lz, here)) break;
        if (!is_solid(here)) {
            keep[i] = 1;
            continue;
        }
        keep[i] = 0;
        if (lz <= 0) continue;
        if (!sim_get(w, gx, gy, (uint32_t)lz - 1u, up)) break;
        if (!is_solid(up)) {
            const uint32_t t = types[i];
            sim_set(w, gx, gy, (uint32_t)lz - 1u, t);
            push4(placed, gx, gy, (uint32_t)lz - 1u, t);
            water_location_updated(next, gx, gy, (uint32_t)lz - 1u);
        }
    }
    counts[0] = placed.n;
    counts[1] = next.n;
}

SimWorld make_world(AreaMap map, uint8_t* voxels, uint8_t* heights, uint16_t* planes, int32_t* status) {
    SimWorld w;
    w.map = map;
    w.voxels = voxels;
    w.heights = heights;
    w.planes = planes;
    w.status = status;
    w.last_ax = w.last_ay = 0;
    w.last_slot = -1;
    return w;
}

}  // namespace

void launch_water_tick(const uint32_t* check, int n, AreaMap map, uint8_t* voxels, uint8_t* heights, uint16_t* planes,
                       uint32_t* changed, uint32_t cap_changed, uint32_t* next, uint32_t cap_next, uint32_t* counts,
                       int32_t* status, cudaStream_t stream) {
    water_tick_kernel<<<1, 32, 0, stream>>>(check, n, make_world(map, voxels, heights, planes, status),
                                           List{changed, 0u, cap_changed, status}, List{next, 0u, cap_next, status}, counts);
}

void launch_falling_check(const uint32_t* check, int n, AreaMap map, uint8_t* voxels, uint8_t* heights, uint16_t* planes,
                          uint32_t* spawned, uint32_t cap_spawned, uint32_t* next, uint32_t cap_next, uint32_t* counts,
                          int32_t* status, cudaStream_t stream) {
    falling_check_kernel<<<1, 1, 0, stream>>>(check, n, make_world(map, voxels, heights, planes, status),
                                              List{spawned, 0u, cap_spawned, status}, List{next, 0u, cap_next, status}, counts);
}

void launch_falling_land(const float* positions, const uint8_t* types, int n, AreaMap map, uint8_t* voxels,
                         uint8_t* heights, uint16_t* planes, uint8_t* keep, uint32_t* placed, uint32_t cap_placed,
                         uint32_t* next, uint32_t cap_next, uint32_t* counts, int32_t* status, cudaStream_t stream) {
    falling_land_kernel<<<1, 1, 0, stream>>>(positions, types, n, make_world(map, voxels, heights, planes, status), keep,
                                             List{placed, 0u, cap_placed, status}, List{next, 0u, cap_next, status}, counts);
}

}  // namespace vx
