// vx_light.cu -- L-ext: voxel light propagation (EXTENSION: the reference has no voxel light, its
// lighting is per-fragment GLSL from the column height map and lamp uniforms).
//
// Specification (DESIGN.md "L-ext"):
//   levels 0..15 per voxel; light CARRIERS are the voxels of Voxel::TRANSPARENT (voxel.rs:39-49);
//   sources: a carrier above its column's max_height (area.rs:34-56) has level 15 (sky), a Lamp has
//   level 15 (emitter, not a carrier); every other carrier has max(0, max over 6 neighbours - 1);
//   everything else 0.  Neighbours outside the listed areas or outside 0 <= z < 128 give nothing.
//
// The least fixed point is reached by chaotic relaxation: each CTA stages one area plus a one-voxel
// halo of its four edge neighbours' current levels in shared memory, relaxes max(nbr - 1) in place
// until nothing changes, and writes the area back; launches repeat until no CTA changed anything.
// Updates are monotone (levels only rise, never above their fixed point), so reading halos that
// other CTAs are updating concurrently is safe and the result is schedule-independent.
#include "vx_kernels.hpp"

namespace vx {
namespace {

constexpr int TW = 18;                                 // tile width with halo
constexpr int TILE = AREA_HEIGHT * TW * TW;            // 41 472 bytes
constexpr uint8_t OPAQUE = 0x80;                       // bit 7: does not carry light

__global__ void __launch_bounds__(256) light_init_kernel(const uint4* __restrict__ jobs,
                                                         const uint8_t* __restrict__ pool_voxels,
                                                         const uint8_t* __restrict__ pool_heights,
                                                         uint8_t* __restrict__ light) {
    const int tid = threadIdx.x;
    const uint32_t slot = jobs[blockIdx.x].z;
    const uint8_t* vox = pool_voxels + (size_t)slot * VOXELS_IN_AREA;
    uint8_t* out = light + (size_t)blockIdx.x * VOXELS_IN_AREA;
    const int mh = pool_heights[(size_t)slot * COLUMNS_IN_AREA + tid];
    for (int z = 0; z < AREA_HEIGHT; ++z) {
        const unsigned v = vox[tid + 256 * z];
        const bool lit = (is_transparent(v) && z < mh) || v == V_LAMP;
        out[tid + 256 * z] = lit ? 15 : 0;
    }
}

__global__ void light_slot_index_kernel(const uint4* __restrict__ jobs, int n, int32_t* __restrict__ slot_to_job) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_to_job[jobs[i].z] = i;
}

__global__ void __launch_bounds__(256) light_relax_kernel(const uint4* __restrict__ jobs,
                                                          const uint8_t* __restrict__ pool_voxels,
                                                          AreaMap map, const int32_t* __restrict__ slot_to_job,
                                                          uint8_t* __restrict__ light,
                                                          uint32_t* __restrict__ changed) {
    extern __shared__ uint8_t tile[];  // [z][y+1][x+1], level | OPAQUE
    const int tid = threadIdx.x;
    const uint4 job = jobs[blockIdx.x];
    const uint8_t* vox = pool_voxels + (size_t)job.z * VOXELS_IN_AREA;
    uint8_t* mine = light + (size_t)blockIdx.x * VOXELS_IN_AREA;
    const int x = tid & 15, y = tid >> 4;

    // neighbours' light arrays (null when the neighbour is not part of this call)
    const uint8_t* nl[4];
    {
        const uint32_t nx[4] = {job.x + 1, job.x - 1, job.x, job.x};
        const uint32_t ny[4] = {job.y, job.y, job.y + 1, job.y - 1};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int s = area_slot(map, nx[d], ny[d]);
            const int j = s >= 0 ? slot_to_job[s] : -1;
            nl[d] = j >= 0 ? light + (size_t)j * VOXELS_IN_AREA : nullptr;
        }
    }
    for (int i = tid; i < TILE; i += 256) tile[i] = OPAQUE;  // corners / missing neighbours
    __syncthreads();
    for (int z = 0; z < AREA_HEIGHT; ++z) {
        const unsigned v = vox[tid + 256 * z];
        const uint8_t l = mine[tid + 256 * z];
        tile[(z * TW + y + 1) * TW + x + 1] = (is_transparent(v) ? 0 : OPAQUE) | l;
    }
    // halo: tid -> (side, t) with 128 z values x 16 positions per side = 2048 entries per side
    for (int i = tid; i < 4 * 2048; i += 256) {
        const int side = i >> 11, k = i & 2047, z = k >> 4, t = k & 15;
        if (!nl[side]) continue;
        uint8_t l;
        int ty, tx;
        if (side == 0) { l = nl[0][0 + 16 * t + 256 * z]; tx = 17; ty = t + 1; }        // +x neighbour, its x = 0
        else if (side == 1) { l = nl[1][15 + 16 * t + 256 * z]; tx = 0; ty = t + 1; }  // -x neighbour, its x = 15
        else if (side == 2) { l = nl[2][t + 16 * 0 + 256 * z]; tx = t + 1; ty = 17; }  // +y neighbour, its y = 0
        else { l = nl[3][t + 16 * 15 + 256 * z]; tx = t + 1; ty = 0; }                 // -y neighbour, its y = 15
        // a neighbour voxel only matters through its level; treat it as a non-updating source
        tile[(z * TW + ty) * TW + tx] = OPAQUE | l;
    }
    __syncthreads();

    bool wrote = false;
    while (true) {
        bool moved = false;
        for (int z = 0; z < AREA_HEIGHT; ++z) {
            const int c = (z * TW + y + 1) * TW + x + 1;
            const uint8_t cur = tile[c];
            if ((cur & OPAQUE) || cur == 15) continue;
            int best = max(max(tile[c + 1] & 15, tile[c - 1] & 15), max(tile[c + TW] & 15, tile[c - TW] & 15));
            if (z + 1 < AREA_HEIGHT) best = max(best, tile[c + TW * TW] & 15);
            if (z > 0) best = max(best, tile[c - TW * TW] & 15);
            if (best - 1 > (int)cur) {
                tile[c] = (uint8_t)(best - 1);
                moved = true;
            }
        }
        wrote |= moved;
        if (!__syncthreads_or(moved)) break;
    }
    for (int z = 0; z < AREA_HEIGHT; ++z) {
        const uint8_t l = tile[(z * TW + y + 1) * TW + x + 1] & 15;
        if (l != mine[tid + 256 * z]) mine[tid + 256 * z] = l;
    }
    if (__syncthreads_or(wrote) && tid == 0) atomicOr(changed, 1u);
}

}  // namespace

void launch_light_init(const uint4* jobs, int n, const uint8_t* pool_voxels,
                       const uint8_t* pool_heights, uint8_t* light, int32_t* slot_to_job,
                       cudaStream_t stream) {
    if (n <= 0) return;
    light_init_kernel<<<n, 256, 0, stream>>>(jobs, pool_voxels, pool_heights, light);
    light_slot_index_kernel<<<(n + 255) / 256, 256, 0, stream>>>(jobs, n, slot_to_job);
}

void launch_light_relax(const uint4* jobs, int n, const uint8_t* pool_voxels, AreaMap map,
                        const int32_t* slot_to_job, uint8_t* light, uint32_t* changed,
                        cudaStream_t stream) {
    if (n <= 0) return;
    light_relax_kernel<<<n, 256, TILE, stream>>>(jobs, pool_voxels, map, slot_to_job, light, changed);
}

}  // namespace vx
