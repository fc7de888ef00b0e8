// vx_edit.cu -- batched block edits (K5) and the height-map image (K6).
//
// K5 replaces World::set (world.rs:184-191) -> Area::set (area.rs:99-111) for a batch that must
// behave as if applied one by one in array order:
//   pass 1  every edit claims its voxel in an open-addressing hash table and records its index
//           with atomicMax -> the highest index per voxel is the sequential winner;
//   pass 2  winners write their voxel, flag their area (and the edge neighbour when the voxel is
//           on a border column: Renderer::update_location re-meshes the six neighbours,
//           renderer.rs:244-285) and append newly flagged areas to the dirty list;
//   pass 3  every edited column is rescanned from the top (area.rs:46-56).  Area::set's
//           incremental rule (rescan if transparent, else min(prev, z)) yields the same value as a
//           rescan whenever max_height was consistent before, which generation guarantees.
// K6 replaces the pixel loop of HeightMap::generate_height_map (height_map.rs:41-85).
#include "vx_kernels.hpp"

namespace vx {
namespace {

constexpr unsigned long long EMPTY_KEY = ~0ull;

__device__ __forceinline__ uint32_t hash_u64(unsigned long long k, uint32_t cap_mask) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return (uint32_t)k & cap_mask;
}

struct Located {
    int slot;
    uint32_t local;  // x + 16 y + 256 z
    unsigned long long key;
};

__device__ __forceinline__ Located locate(const uint4 e, const AreaMap& map) {
    Located l;
    l.slot = area_slot(map, e.x / AREA_SIZE, e.y / AREA_SIZE);  // world.rs:40-52
    l.local = (e.x % AREA_SIZE) + (e.y % AREA_SIZE) * AREA_SIZE + e.z * 256u;
    l.key = (unsigned long long)(uint32_t)l.slot * VOXELS_IN_AREA + l.local;
    return l;
}

__global__ void edits_claim_kernel(const uint4* __restrict__ edits, int n, AreaMap map,
                                   unsigned long long* hash_keys, uint32_t* hash_vals,
                                   uint32_t cap_mask, int32_t* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 e = edits[i];
    const Located l = locate(e, map);
    if (l.slot < 0) {
        atomicMin(status, -5);  // VX_ERR_NOT_LOADED
        return;
    }
    uint32_t pos = hash_u64(l.key, cap_mask);
    while (true) {
        const unsigned long long prev = atomicCAS(&hash_keys[pos], EMPTY_KEY, l.key);
        if (prev == EMPTY_KEY || prev == l.key) break;
        pos = (pos + 1) & cap_mask;
    }
    atomicMax(&hash_vals[pos], (uint32_t)i + 1u);
}

__device__ __forceinline__ void flag_dirty(int slot, uint32_t* dirty_flags, uint32_t* dirty_slots,
                                           uint32_t* dirty_count) {
    if (slot < 0) return;  // neighbour not loaded: get_without_loading -> None, skipped
    if (atomicExch(&dirty_flags[slot], 1u) == 0u) dirty_slots[atomicAdd(dirty_count, 1u)] = (uint32_t)slot;
}

__global__ void edits_apply_kernel(const uint4* __restrict__ edits, int n, AreaMap map,
                                   uint8_t* pool_voxels, uint16_t* pool_planes,
                                   const unsigned long long* hash_keys,
                                   const uint32_t* hash_vals, uint32_t cap_mask,
                                   uint32_t* dirty_flags, uint32_t* dirty_slots,
                                   uint32_t* dirty_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 e = edits[i];
    const Located l = locate(e, map);
    if (l.slot < 0) return;
    uint32_t pos = hash_u64(l.key, cap_mask);
    while (hash_keys[pos] != l.key) pos = (pos + 1) & cap_mask;
    if (hash_vals[pos] != (uint32_t)i + 1u) return;  // a later edit of the same voxel wins
    pool_voxels[(size_t)l.slot * VOXELS_IN_AREA + l.local] = (uint8_t)e.w;
    {   // keep the area's S / T bit planes (vx_mesh.cu) current: one bit each, in 32-bit words
        const uint32_t row = l.local >> 4, bit = 1u << ((l.local & 15u) + 16u * (row & 1u));
        uint32_t* S = reinterpret_cast<uint32_t*>(pool_planes + (size_t)l.slot * (VOXELS_IN_AREA / 8)) + (row >> 1);
        uint32_t* T = S + VOXELS_IN_AREA / 32;
        if (e.w != 0u) atomicOr(S, bit); else atomicAnd(S, ~bit);
        if (is_transparent(e.w)) atomicOr(T, bit); else atomicAnd(T, ~bit);
    }
    flag_dirty(l.slot, dirty_flags, dirty_slots, dirty_count);
    const uint32_t ax = e.x / AREA_SIZE, ay = e.y / AREA_SIZE, lx = e.x % AREA_SIZE, ly = e.y % AREA_SIZE;
    if (lx == 0) flag_dirty(area_slot(map, ax - 1, ay), dirty_flags, dirty_slots, dirty_count);
    if (lx == AREA_SIZE - 1) flag_dirty(area_slot(map, ax + 1, ay), dirty_flags, dirty_slots, dirty_count);
    if (ly == 0) flag_dirty(area_slot(map, ax, ay - 1), dirty_flags, dirty_slots, dirty_count);
    if (ly == AREA_SIZE - 1) flag_dirty(area_slot(map, ax, ay + 1), dirty_flags, dirty_slots, dirty_count);
}

// One warp per edit: lane l looks at z = l, l + 32, l + 64, l + 96 of the edited column (four
// independent loads), a ballot per group finds the first voxel from the top that is not transparent.
// (One thread per edit walked the column as 128 dependent loads, 70 us per batch of 1 000.)
__global__ void __launch_bounds__(256) edits_heights_kernel(const uint4* __restrict__ edits, int n, AreaMap map,
                                                            const uint8_t* __restrict__ pool_voxels,
                                                            uint8_t* pool_heights) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint4 e = edits[i];
    const Located l = locate(e, map);
    if (l.slot < 0) return;
    const uint32_t col = l.local & 255u;
    const uint8_t* vox = pool_voxels + (size_t)l.slot * VOXELS_IN_AREA + col;
    bool opaque[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) opaque[g] = !is_transparent(vox[256 * (32 * g + lane)]);
    int z = AREA_HEIGHT - 1;  // calculate_column_height: nothing opaque -> AREA_HEIGHT - 1 (area.rs:46-56)
#pragma unroll
    for (int g = 3; g >= 0; --g) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, opaque[g]);
        if (m) z = 32 * g + __ffs(m) - 1;
    }
    if (lane == 0) pool_heights[(size_t)l.slot * COLUMNS_IN_AREA + col] = (uint8_t)z;
}

// ------------------------------------------------------------------------------ explosions
// BombSimulator::handle_explosions / explode_at (service/physics/bomb_simulator.rs:191-262): the
// explosions of one tick, taken from the last position to the first; every solid voxel
// (voxel.rs:117-128) whose centre is within EXPLOSION_RADIUS of a position becomes None, and Bomb
// voxels found on the way are reported.  Candidate c is the c-th voxel the reference's loops visit:
// processing order o = c / 1331 (explosion n - 1 - o), then x, y, z nested, z innermost.  A voxel
// inside several spheres belongs to the first explosion that visits it (later ones find None).
constexpr int EXPLOSION_R = 5;                   // EXPLOSION_RADIUS.ceil()
constexpr float EXPLOSION_RADIUS_SQ = 4.5f * 4.5f;
constexpr uint32_t EXPLOSION_SIDE = 2 * EXPLOSION_R + 1;

struct Candidate {
    bool valid;
    uint4 loc;  // internal x, y, z
};

__device__ __forceinline__ Candidate explosion_candidate(const float* __restrict__ positions, int n, uint32_t c) {
    Candidate k;
    const uint32_t o = c / EXPLOSION_CANDIDATES, local = c - o * EXPLOSION_CANDIDATES;
    const float* p = positions + 3 * (size_t)(n - 1 - (int)o);
    const float px = p[0], py = p[1], pz = p[2];
    const int cx = (int)floorf(px), cy = (int)floorf(py), cz = (int)floorf(pz);
    const int x = cx - EXPLOSION_R + (int)(local / (EXPLOSION_SIDE * EXPLOSION_SIDE));
    const int y = cy - EXPLOSION_R + (int)((local / EXPLOSION_SIDE) % EXPLOSION_SIDE);
    const int z = cz - EXPLOSION_R + (int)(local % EXPLOSION_SIDE);
    const int z_min = max(cz - EXPLOSION_R, 0), z_max = min(cz + EXPLOSION_R, AREA_HEIGHT - 2);
    const float dx = (float)x + 0.5f - px, dy = (float)y + 0.5f - py, dz = (float)z + 0.5f - pz;
    k.valid = z >= z_min && z <= z_max && !(dx * dx + dy * dy + dz * dz > EXPLOSION_RADIUS_SQ);
    k.loc = make_uint4((uint32_t)(x + LOCATION_OFFSET), (uint32_t)(y + LOCATION_OFFSET), (uint32_t)z, 0u);
    return k;
}

__global__ void explode_claim_kernel(const float* __restrict__ positions, int n, AreaMap map,
                                     unsigned long long* hash_keys, uint32_t* hash_vals, uint32_t cap_mask,
                                     int32_t* status) {
    const uint32_t total = (uint32_t)n * EXPLOSION_CANDIDATES;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    const Candidate k = explosion_candidate(positions, n, c);
    if (!k.valid) return;
    const Located l = locate(k.loc, map);
    if (l.slot < 0) {
        atomicMin(status, -5);  // VX_ERR_NOT_LOADED
        return;
    }
    uint32_t pos = hash_u64(l.key, cap_mask);
    while (true) {
        const unsigned long long prev = atomicCAS(&hash_keys[pos], EMPTY_KEY, l.key);
        if (prev == EMPTY_KEY || prev == l.key) break;
        pos = (pos + 1) & cap_mask;
    }
    atomicMax(&hash_vals[pos], total - c);  // the earliest visitor wins
}

// flags[c]: bit0 = removed by candidate c, bit1 = it was a Bomb
__global__ void explode_apply_kernel(const float* __restrict__ positions, int n, AreaMap map,
                                     uint8_t* pool_voxels, uint16_t* pool_planes,
                                     const unsigned long long* hash_keys, const uint32_t* hash_vals,
                                     uint32_t cap_mask, uint32_t* dirty_flags, uint32_t* dirty_slots,
                                     uint32_t* dirty_count, uint8_t* __restrict__ flags) {
    const uint32_t total = (uint32_t)n * EXPLOSION_CANDIDATES;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    flags[c] = 0;
    const Candidate k = explosion_candidate(positions, n, c);
    if (!k.valid) return;
    const Located l = locate(k.loc, map);
    if (l.slot < 0) return;
    uint32_t pos = hash_u64(l.key, cap_mask);
    while (hash_keys[pos] != l.key) pos = (pos + 1) & cap_mask;
    if (hash_vals[pos] != total - c) return;  // an earlier explosion of this tick visits it first
    uint8_t* voxel = pool_voxels + (size_t)l.slot * VOXELS_IN_AREA + l.local;
    const uint32_t v = *voxel;
    const bool solid = !(v == V_NONE || (v >= V_WATER_SOURCE && v <= V_WATER4));  // voxel.rs:117-128
    flags[c] = (uint8_t)((solid ? 1u : 0u) | (v == V_BOMB ? 2u : 0u));
    if (!solid) return;
    *voxel = V_NONE;  // world.set(loc, Voxel::None), bomb_simulator.rs:255-256
    {
        const uint32_t row = l.local >> 4, bit = 1u << ((l.local & 15u) + 16u * (row & 1u));
        uint32_t* S = reinterpret_cast<uint32_t*>(pool_planes + (size_t)l.slot * (VOXELS_IN_AREA / 8)) + (row >> 1);
        uint32_t* T = S + VOXELS_IN_AREA / 32;
        atomicAnd(S, ~bit);
        atomicOr(T, bit);
    }
    flag_dirty(l.slot, dirty_flags, dirty_slots, dirty_count);  // renderer.update_location, :209-211
    const uint32_t ax = k.loc.x / AREA_SIZE, ay = k.loc.y / AREA_SIZE, lx = k.loc.x % AREA_SIZE, ly = k.loc.y % AREA_SIZE;
    if (lx == 0) flag_dirty(area_slot(map, ax - 1, ay), dirty_flags, dirty_slots, dirty_count);
    if (lx == AREA_SIZE - 1) flag_dirty(area_slot(map, ax + 1, ay), dirty_flags, dirty_slots, dirty_count);
    if (ly == 0) flag_dirty(area_slot(map, ax, ay - 1), dirty_flags, dirty_slots, dirty_count);
    if (ly == AREA_SIZE - 1) flag_dirty(area_slot(map, ax, ay + 1), dirty_flags, dirty_slots, dirty_count);
}

// Ordered compaction of the flags into the two location lists (visiting order), then the column
// heights of the removed voxels (area.rs:99-111 via a rescan, as for edits).  One CTA.
__global__ void __launch_bounds__(1024) explode_collect_kernel(const float* __restrict__ positions, int n,
                                                               AreaMap map, const uint8_t* __restrict__ flags,
                                                               const uint8_t* __restrict__ pool_voxels,
                                                               uint8_t* pool_heights, uint32_t* __restrict__ removed,
                                                               uint32_t* __restrict__ bombs, uint32_t* counts) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t total = (uint32_t)n * EXPLOSION_CANDIDATES;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < total; base += 1024) {
        const uint32_t c = base + tid;
        const uint32_t f = c < total ? flags[c] : 0u;
        const uint32_t v = (f & 1u) | ((f >> 1) << 16);  // removed | bombs << 16
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w += o;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? s_w[wid - 1] : 0u) + inc - v;
        if (f) {
            const Candidate k = explosion_candidate(positions, n, c);
            if (f & 1u) {
                uint32_t* r = removed + 3 * (size_t)(before & 0xFFFFu);
                r[0] = k.loc.x; r[1] = k.loc.y; r[2] = k.loc.z;
                const Located l = locate(k.loc, map);
                const uint32_t col = l.local & 255u;
                const uint8_t* vox = pool_voxels + (size_t)l.slot * VOXELS_IN_AREA + col;
                int z = 0;
                while (z < AREA_HEIGHT && is_transparent(vox[256 * z])) ++z;
                pool_heights[(size_t)l.slot * COLUMNS_IN_AREA + col] = (uint8_t)(z < AREA_HEIGHT ? z : AREA_HEIGHT - 1);
            }
            if (f & 2u) {
                uint32_t* b = bombs + 3 * (size_t)(before >> 16);
                b[0] = k.loc.x; b[1] = k.loc.y; b[2] = k.loc.z;
            }
        }
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) {
        counts[0] = carry & 0xFFFFu;
        counts[1] = carry >> 16;
    }
}

__global__ void __launch_bounds__(256) heightmap_kernel(const uint2* __restrict__ area_xy, AreaMap map,
                                                        const uint8_t* __restrict__ pool_heights,
                                                        uint32_t cam_ax, uint32_t cam_ay,
                                                        uint32_t* __restrict__ rgba) {
    const uint2 a = area_xy[blockIdx.x];
    const int dx = (int)a.x - (int)cam_ax, dy = (int)a.y - (int)cam_ay;
    // filter_distant_areas (height_map.rs:90-101): |d| < 32 / 2
    if (dx <= -16 || dx >= 16 || dy <= -16 || dy >= 16) return;
    const int slot = area_slot(map, a.x, a.y);
    const int tid = threadIdx.x, x = tid & 15, y = tid >> 4;
    // get_area_without_loading falls back to an empty area: height 127 (world.rs:64-69)
    const uint32_t h = slot >= 0 ? pool_heights[(size_t)slot * COLUMNS_IN_AREA + tid] : 127u;
    const int image_x = 256 + dx * AREA_SIZE + x, image_y = 256 + dy * AREA_SIZE + y;
    rgba[image_x + image_y * 512] = h * 0x01010101u;  // [height; 4]
}

}  // namespace

void launch_apply_edits(const uint4* edits, int n, AreaMap map, uint8_t* pool_voxels,
                        uint8_t* pool_heights, uint16_t* pool_planes, unsigned long long* hash_keys,
                        uint32_t* hash_vals, uint32_t hash_cap, uint32_t* dirty_flags, uint32_t* dirty_slots,
                        uint32_t* dirty_count, int32_t* status, cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = (n + 255) / 256;
    edits_claim_kernel<<<blocks, 256, 0, stream>>>(edits, n, map, hash_keys, hash_vals, hash_cap - 1,
                                                   status);
    edits_apply_kernel<<<blocks, 256, 0, stream>>>(edits, n, map, pool_voxels, pool_planes, hash_keys, hash_vals,
                                                   hash_cap - 1, dirty_flags, dirty_slots,
                                                   dirty_count);
    edits_heights_kernel<<<(n + 7) / 8, 256, 0, stream>>>(edits, n, map, pool_voxels, pool_heights);
}

void launch_explode(const float* positions, int n, AreaMap map, uint8_t* pool_voxels, uint8_t* pool_heights,
                    uint16_t* pool_planes, unsigned long long* hash_keys, uint32_t* hash_vals, uint32_t hash_cap,
                    uint32_t* dirty_flags, uint32_t* dirty_slots, uint32_t* dirty_count, int32_t* status,
                    uint8_t* flags, uint32_t* removed, uint32_t* bombs, uint32_t* counts, cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = (n * (int)EXPLOSION_CANDIDATES + 255) / 256;
    explode_claim_kernel<<<blocks, 256, 0, stream>>>(positions, n, map, hash_keys, hash_vals, hash_cap - 1, status);
    explode_apply_kernel<<<blocks, 256, 0, stream>>>(positions, n, map, pool_voxels, pool_planes, hash_keys, hash_vals,
                                                     hash_cap - 1, dirty_flags, dirty_slots, dirty_count, flags);
    explode_collect_kernel<<<1, 1024, 0, stream>>>(positions, n, map, flags, pool_voxels, pool_heights, removed, bombs,
                                                   counts);
}

void launch_heightmap(const uint2* area_xy, int n, AreaMap map, const uint8_t* pool_heights,
                      uint32_t cam_ax, uint32_t cam_ay, uint32_t* rgba, cudaStream_t stream) {
    if (n <= 0) return;
    heightmap_kernel<<<n, 256, 0, stream>>>(area_xy, map, pool_heights, cam_ax, cam_ay, rgba);
}

}  // namespace vx
