// vx_mesh.cu -- face-culled meshing (sm_100a), HBM-bound, no tensor cores.
//
// Replaces Renderer::load_full_area (renderer.rs:310-322):
//   get_renderable_voxels_for_area (world.rs:112-147) + Area::has_non_transparent_neighbours
//   (area.rs:126-201) + generate_mesh_for_voxel (renderer.rs:126-195) + should_generate_face
//   (mesh_generator.rs:501-518) + generate_mesh / get_verticies_for_voxel
//   (mesh_generator.rs:109-147,200-498).
//
// Data flow per area (one 256-thread CTA):
//   planes : each 16-voxel row (fixed y,z) is reduced to two 16-bit planes, S = "voxel != None"
//            and T = "voxel in Voxel::TRANSPARENT".  face(cur, nbr) = cur != nbr && T(nbr)
//            (the algebraic form of mesh_generator.rs:501-513), so for opaque voxels the six face
//            masks of a row are S & shifted/neighbouring T rows -- pure bit operations, with the
//            one-voxel halo taken from the four edge neighbours' planes.  Only transparent
//            non-None voxels (water, ice, glass) need the byte-wise cur != nbr refinement.
//   count  : popc of the masks, CTA reduction -> faces and visible voxels per area.
//   offsets: exclusive scan over areas -> where each area's records / faces start.
//   emit   : the same masks again; a CTA-wide exclusive scan puts voxels in the reference's
//            z,y,x order and faces in push order; face descriptors are staged in shared memory so
//            that the 160-byte vertex blocks and 12-byte index blocks leave as fully coalesced
//            16-byte / 4-byte stores.
#include "vx_face.cuh"
#include "vx_kernels.hpp"

namespace vx {
namespace {

constexpr int ROWS = VOXELS_IN_AREA / 16;  // 2048 rows of 16 voxels

// ------------------------------------------------------------------------------------- planes
// plane layout per slot: uint16 S[2048] followed by uint16 T[2048] (8 KiB)
// (the planes of a resident area are written by generation / upload / edits, never here)
// One thread per requested area: resolve its pool slot and those of its four edge neighbours
// through the device map and report what is not resident.  Keeps the host out of the per-area loop (no hash lookups on the CPU).
//   counters[0] = requested areas that are not resident (an error)
//   counters[1] = neighbour areas that are not resident (the host generates them and retries)
__global__ void mesh_prepare_kernel(const uint2* __restrict__ xy, int n, AreaMap map,
                                    uint4* __restrict__ jobs, uint4* __restrict__ nbr,
                                    uint32_t* __restrict__ counters,
                                    uint2* __restrict__ missing, uint32_t missing_cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 a = xy[i];
    const int slot = area_slot(map, a.x, a.y);
    if (slot < 0) {
        atomicAdd(&counters[0], 1u);
        jobs[i] = make_uint4(a.x, a.y, 0u, 0u);
        nbr[i] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    jobs[i] = make_uint4(a.x, a.y, (uint32_t)slot, 0u);
    const uint2 nb[4] = {make_uint2(a.x + 1, a.y), make_uint2(a.x - 1, a.y), make_uint2(a.x, a.y + 1),
                         make_uint2(a.x, a.y - 1)};
    uint32_t ns[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int sl = area_slot(map, nb[d].x, nb[d].y);
        ns[d] = sl < 0 ? (uint32_t)slot : (uint32_t)sl;  // placeholder; the host generates it and retries
        if (sl < 0) {
            const uint32_t k = atomicAdd(&counters[1], 1u);
            if (k < missing_cap) missing[k] = nb[d];
        }
    }
    nbr[i] = make_uint4(ns[0], ns[1], ns[2], ns[3]);  // +x, -x, +y, -y: the kernels below do no lookups
}

// ------------------------------------------------------------------------------------- masks
struct AreaView {
    const uint8_t* vox;        // own voxels
    const uint8_t* vox_n[4];   // +x, -x, +y, -y neighbour voxels
    const uint16_t* S;         // own planes
    const uint16_t* T;
    const uint16_t* T_n[4];    // neighbour T planes
};

struct RowMasks {
    uint32_t m[6];  // L(x+1) R(x-1) F(y+1) B(y-1) D(z+1) U(z-1): bit x set = voxel x has the face
};

// The planes one area's culling reads, staged in shared memory by fully coalesced loads: own S and
// T (2 x 4 KiB), per row the two x-halo bits (T of the +x neighbour's x = 0 and of the -x
// neighbour's x = 15), and the y-halo rows (T of the +y neighbour's y = 0 row and of the -y
// neighbour's y = 15 row, per z).
struct alignas(16) PlaneTile {
    uint16_t S[ROWS];
    uint16_t T[ROWS];
    uint8_t hx[ROWS];              // bit0: T(+x neighbour, x = 0), bit1: T(-x neighbour, x = 15)
    uint16_t ty[2][AREA_HEIGHT];   // [0]: +y neighbour row y = 0, [1]: -y neighbour row y = 15
    uint8_t slab[AREA_HEIGHT];     // per z: bit0 = some voxel != None, bit1 = some transparent voxel
                                   // in the slab or in its one-voxel halo
};

__device__ __forceinline__ void load_plane_tile(PlaneTile& p, const AreaView& a, int tid) {
    const uint4* s4 = reinterpret_cast<const uint4*>(a.S);
    const uint4* t4 = reinterpret_cast<const uint4*>(a.T);
    uint4* ds = reinterpret_cast<uint4*>(p.S);
    uint4* dt = reinterpret_cast<uint4*>(p.T);
    const uint4* xp4 = reinterpret_cast<const uint4*>(a.T_n[0]);
    const uint4* xm4 = reinterpret_cast<const uint4*>(a.T_n[1]);
    uint2* dh = reinterpret_cast<uint2*>(p.hx);
    static_assert(ROWS / 8 == 256, "one iteration per thread: thread i owns rows 8i .. 8i+7, half of slab i / 2");
    {
        const int i = tid;
        const uint4 sv = s4[i], tv = t4[i];
        const uint4 xp = xp4[i], xm = xm4[i];
        const int z = i >> 1;
        uint32_t ty_or = 0;
        if (!(i & 1)) {
            const uint32_t t0 = a.T_n[2][16 * z];       // row (y = 0, z) of the +y neighbour
            const uint32_t t1 = a.T_n[3][16 * z + 15];  // row (y = 15, z) of the -y neighbour
            p.ty[0][z] = (uint16_t)t0;
            p.ty[1][z] = (uint16_t)t1;
            ty_or = t0 | t1;
        }
        ds[i] = sv;
        dt[i] = tv;
        // per row one byte: bit0 = T(+x neighbour, x = 0), bit1 = T(-x neighbour, x = 15); a 32-bit
        // word holds two rows, so both bits of both rows are gathered with a mask, a shift and one
        // byte permute (bytes 0 and 2 of the gathered word)
        const uint32_t wp[4] = {xp.x, xp.y, xp.z, xp.w}, wm[4] = {xm.x, xm.y, xm.z, xm.w};
        uint32_t pair[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t g = (wp[k] & 0x00010001u) | ((wm[k] >> 14) & 0x00020002u);
            pair[k] = __byte_perm(g, 0u, 0x4420);  // rows 2k, 2k + 1 -> bytes 0, 1
        }
        const uint32_t h[2] = {pair[0] | (pair[1] << 16), pair[2] | (pair[3] << 16)};
        dh[i] = make_uint2(h[0], h[1]);
        uint32_t any_s = sv.x | sv.y | sv.z | sv.w;
        uint32_t any_t = tv.x | tv.y | tv.z | tv.w | h[0] | h[1] | ty_or;
        any_s |= __shfl_xor_sync(0xFFFFFFFFu, any_s, 1);
        any_t |= __shfl_xor_sync(0xFFFFFFFFu, any_t, 1);
        if (!(i & 1)) p.slab[z] = (uint8_t)((any_s != 0u) | ((any_t != 0u) << 1));
    }
}

// No voxel of slabs z0, z0 + 1 can have a face: they are all None, or nothing transparent (None
// included) touches them -- every face mask is "S & some T row of slab z - 1, z or z + 1 or of the
// halo".  Holds for most of an area: the air above the surface and the rock below it.
__device__ __forceinline__ bool slab_pair_empty(const PlaneTile& p, int z0) {
    const uint32_t f0 = p.slab[z0], f1 = p.slab[z0 + 1];
    const uint32_t below = z0 > 0 ? p.slab[z0 - 1] : 0u, above = z0 + 2 < AREA_HEIGHT ? p.slab[z0 + 2] : 0u;
    const bool faces0 = (f0 & 1u) && (((below | f0 | f1) & 2u) != 0u);
    const bool faces1 = (f1 & 1u) && (((f0 | f1 | above) & 2u) != 0u);
    return !(faces0 || faces1);
}

// bit x = byte x of a equals byte x of b (16 bytes, little endian over .x .y .z .w)
__device__ __forceinline__ uint32_t equal_bytes(const uint4 a, const uint4 b) {
    // __vcmpeq4: 0xFF per equal byte; the four top bits are gathered by a multiply whose partial
    // products land on distinct bits, the wanted ones on bits 28..31
    const uint32_t e0 = ((__vcmpeq4(a.x, b.x) & 0x80808080u) * 0x00204081u) >> 28;
    const uint32_t e1 = ((__vcmpeq4(a.y, b.y) & 0x80808080u) * 0x00204081u) >> 28;
    const uint32_t e2 = ((__vcmpeq4(a.z, b.z) & 0x80808080u) * 0x00204081u) >> 28;
    const uint32_t e3 = ((__vcmpeq4(a.w, b.w) & 0x80808080u) * 0x00204081u) >> 28;
    return e0 | (e1 << 4) | (e2 << 8) | (e3 << 12);
}

__device__ __forceinline__ RowMasks row_masks(const PlaneTile& p, const AreaView& a, int r) {
    RowMasks k;
    const uint32_t s = p.S[r];
    if (s == 0) {
#pragma unroll
        for (int d = 0; d < 6; ++d) k.m[d] = 0;
        return k;
    }
    const int y = r & 15, z = r >> 4;
    const uint32_t t = p.T[r], hx = p.hx[r];
    const uint32_t t_xp = (t >> 1) | ((hx & 1u) << 15);
    const uint32_t t_xm = ((t << 1) & 0xFFFFu) | (hx >> 1);
    const uint32_t t_yp = y < 15 ? p.T[r + 1] : p.ty[0][z];
    const uint32_t t_ym = y > 0 ? p.T[r - 1] : p.ty[1][z];
    const uint32_t t_zp = z < AREA_HEIGHT - 1 ? p.T[r + 16] : 0u;  // renderer.rs:165
    const uint32_t t_zm = z > 0 ? p.T[r - 16] : 0u;                // renderer.rs:173
    k.m[0] = s & t_xp;
    k.m[1] = s & t_xm;
    k.m[2] = s & t_yp;
    k.m[3] = s & t_ym;
    k.m[4] = s & t_zp;
    k.m[5] = s & t_zm;

    // transparent non-None voxels: apply cur != nbr, the top-face rule for partial-height water
    // (mesh_generator.rs:515-518) and the enclosed-voxel pre-filter (area.rs:126-201)
    uint32_t w = s & t;
#ifndef VX_MESH_BYTEWISE_REFINE
    // All 16 voxels of the row at once: the row and the rows it is compared with are 16-byte loads
    // (issued together, no per-voxel chain of byte loads through a water body), cur == nbr is a
    // byte-wise compare whose results are gathered into a 16-bit mask.
    if (w) {
        const uint4* rows = reinterpret_cast<const uint4*>(a.vox);
        const uint4 c = rows[r];
        if (k.m[0] & w) {
            const uint32_t halo = ((k.m[0] & w) >> 15) ? a.vox_n[0][16 * r] : 0u;
            const uint4 n = make_uint4(__funnelshift_r(c.x, c.y, 8), __funnelshift_r(c.y, c.z, 8),
                                       __funnelshift_r(c.z, c.w, 8), (c.w >> 8) | (halo << 24));
            k.m[0] &= ~(equal_bytes(c, n) & w);
        }
        if (k.m[1] & w) {
            const uint32_t halo = (k.m[1] & w & 1u) ? a.vox_n[1][16 * r + 15] : 0u;
            const uint4 n = make_uint4((c.x << 8) | halo, __funnelshift_l(c.x, c.y, 8), __funnelshift_l(c.y, c.z, 8),
                                       __funnelshift_l(c.z, c.w, 8));
            k.m[1] &= ~(equal_bytes(c, n) & w);
        }
        if (k.m[2] & w)
            k.m[2] &= ~(equal_bytes(c, y < 15 ? rows[r + 1] : reinterpret_cast<const uint4*>(a.vox_n[2])[r - 15]) & w);
        if (k.m[3] & w)
            k.m[3] &= ~(equal_bytes(c, y > 0 ? rows[r - 1] : reinterpret_cast<const uint4*>(a.vox_n[3])[r + 15]) & w);
        if (k.m[4] & w) k.m[4] &= ~(equal_bytes(c, rows[r + 16]) & w);  // a face bit implies z < AREA_HEIGHT - 1
        if (k.m[5] & w) k.m[5] &= ~(equal_bytes(c, rows[r - 16]) & w);  // a face bit implies z > 0
        const uint4 partial = make_uint4(c.x & 0xFCFCFCFCu, c.y & 0xFCFCFCFCu, c.z & 0xFCFCFCFCu, c.w & 0xFCFCFCFCu);
        static_assert(V_WATER1 == 16 && V_WATER4 == 19, "partial-height voxels are the IDs 0x10..0x13");
        const uint32_t ph = equal_bytes(partial, make_uint4(0x10101010u, 0x10101010u, 0x10101010u, 0x10101010u)) & w;
        if (ph && z > 0) {
            const uint32_t any_plain = k.m[0] | k.m[1] | k.m[2] | k.m[3] | k.m[4] | k.m[5];
            const uint32_t interior = (y > 0 && y < 15 && z < AREA_HEIGHT - 1) ? 0x7FFEu : 0u;
            // an enclosed interior voxel is dropped by get_renderable_voxels_for_area before the
            // top-face rule is ever consulted (world.rs:137-139)
            k.m[5] |= ph & (any_plain | ~interior) & 0xFFFFu;
        }
    }
#else
    while (w) {
        const int x = __ffs(w) - 1;
        w &= w - 1;
        const uint32_t bit = 1u << x;
        const int idx = x + 16 * r;
        const uint32_t cur = a.vox[idx];
        if (k.m[0] & bit) {
            const uint32_t n = x < 15 ? a.vox[idx + 1] : a.vox_n[0][idx - 15];
            if (n == cur) k.m[0] &= ~bit;
        }
        if (k.m[1] & bit) {
            const uint32_t n = x > 0 ? a.vox[idx - 1] : a.vox_n[1][idx + 15];
            if (n == cur) k.m[1] &= ~bit;
        }
        if (k.m[2] & bit) {
            const uint32_t n = y < 15 ? a.vox[idx + 16] : a.vox_n[2][idx - 240];
            if (n == cur) k.m[2] &= ~bit;
        }
        if (k.m[3] & bit) {
            const uint32_t n = y > 0 ? a.vox[idx - 16] : a.vox_n[3][idx + 240];
            if (n == cur) k.m[3] &= ~bit;
        }
        if ((k.m[4] & bit) && a.vox[idx + 256] == cur) k.m[4] &= ~bit;
        if ((k.m[5] & bit) && a.vox[idx - 256] == cur) k.m[5] &= ~bit;
        if (partial_height(cur) && z > 0) {
            const bool any_plain = ((k.m[0] | k.m[1] | k.m[2] | k.m[3] | k.m[4] | k.m[5]) & bit) != 0;
            const bool interior = x > 0 && x < 15 && y > 0 && y < 15 && z < AREA_HEIGHT - 1;
            // an enclosed interior voxel is dropped by get_renderable_voxels_for_area before the
            // top-face rule is ever consulted (world.rs:137-139)
            if (any_plain || !interior) k.m[5] |= bit;
        }
    }
#endif
    return k;
}

__device__ __forceinline__ AreaView make_view(const uint4 job, const uint4 nb,
                                              const uint8_t* pool_voxels,
                                              const uint16_t* pool_planes) {
    AreaView a;
    const int slot = (int)job.z;
    a.vox = pool_voxels + (size_t)slot * VOXELS_IN_AREA;
    a.S = pool_planes + (size_t)slot * (2 * ROWS);
    a.T = a.S + ROWS;
    const uint32_t ns[4] = {nb.x, nb.y, nb.z, nb.w};  // resident: vx_mesh generates missing ones first
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a.vox_n[i] = pool_voxels + (size_t)ns[i] * VOXELS_IN_AREA;
        a.T_n[i] = pool_planes + (size_t)ns[i] * (2 * ROWS) + ROWS;
    }
    return a;
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

__device__ __forceinline__ unsigned long long warp_inclusive_scan64(unsigned long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// ------------------------------------------------------------------------------------- count
// faces | visible voxels << 16 of one row
__device__ __forceinline__ uint32_t row_counts(const RowMasks& k) {
    const uint32_t vis = k.m[0] | k.m[1] | k.m[2] | k.m[3] | k.m[4] | k.m[5];
    return (__popc(k.m[0]) + __popc(k.m[1]) + __popc(k.m[2]) + __popc(k.m[3]) + __popc(k.m[4]) +
            __popc(k.m[5])) | ((uint32_t)__popc(vis) << 16);
}

// Per (256-row chunk, warp) of an area: faces (<= 3072, 12 bits) | visible voxels (<= 512, 10 bits)
// << 12 | rows that have faces (<= 32, 6 bits) << 22.
__device__ __forceinline__ uint32_t pack_warp_total(uint32_t faces, uint32_t recs, uint32_t rows) {
    return faces | (recs << 12) | (rows << 22);
}

// Per area: faces and visible voxels, in total and per (256-row chunk, warp) -- the latter, 64
// words per area, let the emit kernel place every row without recounting the area.
// Eight CTAs per SM (32 registers): the kernel is a chain of dependent round trips per area, so it
// is the resident warps that hide them (6 CTAs at 40 registers: 0.102 ms, 8: 0.076 ms).  Requesting
// the planes of the CTA that follows on the SM into L2 ahead of time changed nothing (measured).
#ifndef VX_COUNT_CTAS_PER_SM
#define VX_COUNT_CTAS_PER_SM 8
#endif
__global__ void __launch_bounds__(256, VX_COUNT_CTAS_PER_SM) mesh_count_kernel(const uint4* __restrict__ jobs,
                                                         const uint4* __restrict__ nbr,
                                                         const uint8_t* __restrict__ pool_voxels,
                                                         const uint16_t* __restrict__ pool_planes,
                                                         uint32_t* __restrict__ rec_count,
                                                         uint32_t* __restrict__ face_count,
                                                         uint32_t* __restrict__ warp_totals) {
    __shared__ uint32_t s_faces[8], s_recs[8];
    __shared__ uint32_t s_busy[2];  // bit c * 8 + w: (chunk c, warp w) may have faces
    __shared__ PlaneTile tile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // last area first: generation has just written the planes in list order, the end of the list is
    // what the L2 still holds (count 0.0769 -> 0.0754 ms; emit walks the list the same way: 0.3455 ->
    // 0.3435, profiles/r02_ab_mesh_order.log)
    const unsigned area = gridDim.x - 1u - blockIdx.x;
    const AreaView a = make_view(jobs[area], nbr[area], pool_voxels, pool_planes);
    load_plane_tile(tile, a, tid);
    __syncthreads();
    uint32_t* wt = warp_totals + (size_t)area * 64;
    if (tid < 64) {  // one thread per (chunk, warp): rows of two slabs
        const bool empty = slab_pair_empty(tile, 16 * (tid >> 3) + 2 * (tid & 7));
        if (empty) wt[tid] = 0u;
        const uint32_t busy = __ballot_sync(0xFFFFFFFFu, !empty);
        if (lane == 0) s_busy[wid] = busy;
    }
    __syncthreads();
    // this warp's chunks: bits wid, 8 + wid, ... of the 64-bit mask
    // (bit wid of each byte, gathered into a nibble by a multiply: the partial products land on
    // distinct bits, the wanted ones on bits 24..27)
    static_assert(ROWS / 256 == 8, "two words of four chunks each");
    const uint32_t mine_lo = ((((s_busy[0] >> wid) & 0x01010101u) * 0x01020408u) >> 24) & 15u;
    const uint32_t mine_hi = ((((s_busy[1] >> wid) & 0x01010101u) * 0x01020408u) >> 24) & 15u;
    uint32_t mine = mine_lo | (mine_hi << 4);
    uint32_t faces = 0, recs = 0;
    while (mine) {
        const int c = __ffs(mine) - 1;
        mine &= mine - 1;
        const uint32_t v = row_counts(row_masks(tile, a, c * 256 + tid));
        const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, v);  // <= 32 * 96 faces, 32 * 16 voxels
        const uint32_t rows = __popc(__ballot_sync(0xFFFFFFFFu, v != 0u));
        if (lane == 0) wt[c * 8 + wid] = pack_warp_total(tot & 0xFFFFu, tot >> 16, rows);
        faces += tot & 0xFFFFu;
        recs += tot >> 16;
    }
    if (lane == 0) {
        s_faces[wid] = faces;
        s_recs[wid] = recs;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t f = 0, q = 0;
        for (int i = 0; i < 8; ++i) {
            f += s_faces[i];
            q += s_recs[i];
        }
        face_count[area] = f;
        rec_count[area] = q;
    }
}

// ------------------------------------------------------------------------------------- offsets
// single-CTA exclusive scan over the per-area counts (n <= a few 100k: microseconds); four areas
// per thread and round, so 16 384 areas take four rounds
__global__ void __launch_bounds__(1024) mesh_offsets_kernel(const uint32_t* __restrict__ rec_count,
                                                            const uint32_t* __restrict__ face_count,
                                                            int n, uint32_t* __restrict__ rec_first,
                                                            uint32_t* __restrict__ face_first,
                                                            const uint32_t* __restrict__ counters,
                                                            uint32_t* __restrict__ host_totals) {
    __shared__ uint32_t s_r[32], s_f[32];
    __shared__ uint32_t carry_r, carry_f;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_r = carry_f = 0;
    __syncthreads();
    constexpr int PER = 16;  // areas per thread and round: four 16-byte loads of each array in flight
    for (int base = 0; base < n; base += 1024 * PER) {
        const int i = base + PER * tid;
        uint32_t r[PER], f[PER];
        if (i + PER <= n) {
#pragma unroll
            for (int q = 0; q < PER / 4; ++q) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rec_count + i + 4 * q);
                const uint4 fv = *reinterpret_cast<const uint4*>(face_count + i + 4 * q);
                r[4 * q] = rv.x; r[4 * q + 1] = rv.y; r[4 * q + 2] = rv.z; r[4 * q + 3] = rv.w;
                f[4 * q] = fv.x; f[4 * q + 1] = fv.y; f[4 * q + 2] = fv.z; f[4 * q + 3] = fv.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                r[k] = i + k < n ? rec_count[i + k] : 0u;
                f[k] = i + k < n ? face_count[i + k] : 0u;
            }
        }
        uint32_t sr = 0, sf = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            sr += r[k];
            sf += f[k];
        }
        const uint32_t ir = warp_inclusive_scan(sr, lane), jf = warp_inclusive_scan(sf, lane);
        if (lane == 31) {
            s_r[wid] = ir;
            s_f[wid] = jf;
        }
        __syncthreads();
        if (wid == 0) {
            const uint32_t wr = warp_inclusive_scan(s_r[lane], lane), wf = warp_inclusive_scan(s_f[lane], lane);
            s_r[lane] = wr;
            s_f[lane] = wf;
        }
        __syncthreads();
        uint32_t pr = carry_r + (wid ? s_r[wid - 1] : 0u) + ir - sr, pf = carry_f + (wid ? s_f[wid - 1] : 0u) + jf - sf;
        if (i + PER <= n) {
#pragma unroll
            for (int q = 0; q < PER / 4; ++q) {
                uint4 ro, fo;
                ro.x = pr; pr += r[4 * q]; ro.y = pr; pr += r[4 * q + 1]; ro.z = pr; pr += r[4 * q + 2]; ro.w = pr; pr += r[4 * q + 3];
                fo.x = pf; pf += f[4 * q]; fo.y = pf; pf += f[4 * q + 1]; fo.z = pf; pf += f[4 * q + 2]; fo.w = pf; pf += f[4 * q + 3];
                *reinterpret_cast<uint4*>(rec_first + i + 4 * q) = ro;
                *reinterpret_cast<uint4*>(face_first + i + 4 * q) = fo;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                if (i + k < n) {
                    rec_first[i + k] = pr;
                    face_first[i + k] = pf;
                }
                pr += r[k];
                pf += f[k];
            }
        }
        __syncthreads();
        if (tid == 1023) {
            carry_r = pr;
            carry_f = pf;
        }
        __syncthreads();
    }
    if (tid == 0) {
        rec_first[n] = carry_r;
        face_first[n] = carry_f;
        // what the host waits for, written straight into its (mapped, pinned) memory: it reads them
        // as soon as this kernel is done, while emit is already running
        host_totals[0] = counters[0];  // listed areas that are not resident (mesh_prepare)
        host_totals[1] = counters[1];  // edge neighbours that are not resident
        host_totals[2] = carry_r;
        host_totals[3] = carry_f;
    }
}

// ------------------------------------------------------------------------------------- emit
// (corner codes, face descriptors and vertex assembly: vx_face.cuh)
#ifndef VX_EMIT_LIST
#define VX_EMIT_LIST 3072
#endif
#ifndef VX_EMIT_ROWS
#define VX_EMIT_ROWS 512
#endif
#ifndef VX_EMIT_WARP_FACES
#define VX_EMIT_WARP_FACES 16
#endif
#ifndef VX_EMIT_CTAS
#define VX_EMIT_CTAS 5
#endif
constexpr int LIST_FACES = VX_EMIT_LIST;        // face descriptors buffered per flush (typical area: ~800)
constexpr int ROW_CAP = VX_EMIT_ROWS;           // rows with faces staged per pass (typical area: ~100)
constexpr int WARP_FACES = VX_EMIT_WARP_FACES;  // faces a warp assembles per staging round (2 lanes per face)
constexpr int FACE_U4 = 10;        // uint4 per face (4 vertices x 40 B)
constexpr int HALF_U4 = 5;         // a lane stages two vertices: 5 uint4 at (lane / 2) * 10 + (lane % 2) * 5,
                                   // so the 8 lanes of a store phase hit 8 different bank groups
constexpr int CHUNKS = ROWS / 256; // rows per thread

// The rows of the current pass that have faces, in any order (40 B per row).
struct RowEntries {
    uint4 vox[ROW_CAP];    // the row's 16 voxel bytes
    uint4 masks[ROW_CAP];  // face masks L | R << 16, F | B << 16, D | U << 16; row index
    uint2 first[ROW_CAP];  // area-local index of the row's first face and first record
};

struct EmitSmem {
    PlaneTile planes;
    uint32_t list[LIST_FACES];
    union {  // rows feed the descriptor list, then the list is flushed through the staging tiles
        RowEntries rows;
        uint4 stage[8][WARP_FACES * FACE_U4];  // one staging tile per warp
    } u;
    unsigned long long part[CHUNKS * 8 + 1];  // running sums before (chunk, warp); [64] = area totals
};
static_assert(sizeof(RowEntries) <= sizeof(uint4) * 8 * WARP_FACES * FACE_U4, "row entries alias the staging tiles");

// running sums: faces (<= 98 304) | visible voxels (<= 32 768) << 20 | rows (<= 2048) << 40
__device__ __forceinline__ unsigned long long widen(uint32_t w) {
    return (unsigned long long)(w & 0xFFFu) | ((unsigned long long)((w >> 12) & 0x3FFu) << 20) |
           ((unsigned long long)(w >> 22) << 40);
}
__device__ __forceinline__ uint32_t faces_of(unsigned long long p) { return (uint32_t)p & 0xFFFFFu; }
__device__ __forceinline__ uint32_t recs_of(unsigned long long p) { return (uint32_t)(p >> 20) & 0xFFFFFu; }
__device__ __forceinline__ uint32_t rows_of(unsigned long long p) { return (uint32_t)(p >> 40); }

// Write `count` buffered faces (descriptors sm.list[0..count)) starting at stream face `first`.
// Each warp works on its own: 16 faces are assembled in the warp's staging tile by lane pairs, then
// the tile leaves as one TMA bulk copy; no CTA barrier is involved.  (One lane per face with
// 80-byte half tiles needs fewer instructions but writes partial sectors: measured slower.)
__device__ __forceinline__ void flush_faces(EmitSmem& sm, uint32_t count, uint32_t first,
                                            const MeshOut& out, uint32_t gx0, uint32_t gy0, int tid) {
    const int lane = tid & 31, wid = tid >> 5;
    uint4* tile = sm.u.stage[wid];
#pragma unroll 1
    for (uint32_t base = wid * WARP_FACES; base < count; base += 8 * WARP_FACES) {
        const uint32_t nfaces = min((uint32_t)WARP_FACES, count - base);
        const uint32_t f = lane >> 1;
        if (f < nfaces) {
            const FaceParts p = decode_face(sm.list[base + f], gx0, gy0);
            build_half_face(&tile[f * FACE_U4 + (lane & 1) * HALF_U4], p, lane & 1);
        }
        // The tile leaves through the TMA unit as one bulk copy (shared -> global, up to 2 560 B):
        // the writers make their shared-memory stores visible to the async proxy, one lane issues
        // the copy and waits only until the tile has been READ -- the global writes drain behind
        // the warp's back while it assembles the next 16 faces.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint4* vdst = out.vertices + (size_t)(first + base) * FACE_U4;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(vdst), "r"((uint32_t)__cvta_generic_to_shared(tile)), "r"(nfaces * (FACE_U4 * 16u))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    // indices: [0,1,2,0,2,3] + 4*ord as three u32 words per face (mesh_generator.rs:44,131-139);
    // one thread per face, a warp covers 384 consecutive bytes with its three stores
    uint32_t* idst = out.indices + (size_t)first * 3;
    for (uint32_t fi = tid; fi < count; fi += 256) {
        const uint32_t b = ((sm.list[fi] >> 18) & 7u) * 4u;
        uint32_t* w = idst + 3 * fi;
        w[0] = b | ((b + 1u) << 16);
        w[1] = (b + 2u) | (b << 16);
        w[2] = (b + 2u) | ((b + 3u) << 16);
    }
}

// One CTA per area, three phases per pass (one pass for all but the densest areas):
//   A  one thread per row (the reference's z,y order): face masks from the plane tile, position of
//      the row's faces / records from the count kernel's (chunk, warp) totals and a warp scan; rows
//      that have faces are staged in shared memory with their 16 voxel bytes.  (chunk, warp) pairs
//      whose total is 0 -- all air or all buried, most of them -- are skipped.
//   B  one thread per visible voxel of the staged rows (binary search over the rows' record
//      prefixes): it writes its record and pushes its face descriptors.  The faces cluster in a few
//      z levels, i.e. in a few warps of phase A; spreading them over the CTA by voxel is what keeps
//      the warps evenly busy.
//   C  flush: lane pairs expand descriptors into 160-byte vertex blocks (flush_faces).
__global__ void __launch_bounds__(256, VX_EMIT_CTAS) mesh_emit_kernel(const uint4* __restrict__ jobs,
                                                        const uint4* __restrict__ nbr,
                                                        const uint8_t* __restrict__ pool_voxels,
                                                        const uint16_t* __restrict__ pool_planes,
                                                        const uint32_t* __restrict__ warp_totals,
                                                        MeshOut out, int n, uint32_t cap_faces,
                                                        uint32_t cap_recs) {
    __shared__ EmitSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // everything this CTA needs from global memory before the planes, requested together
    const uint32_t stream_faces = out.face_first[n], stream_recs = out.rec_first[n];
    const unsigned area = gridDim.x - 1u - blockIdx.x;  // as the count kernel
    const uint4 job = jobs[area];
    const uint4 nb = nbr[area];
    const uint32_t face_base = out.face_first[area];
    const uint32_t rec_base = out.rec_first[area];
    uint2 w2 = make_uint2(0u, 0u);
    if (wid == 0) w2 = reinterpret_cast<const uint2*>(warp_totals + (size_t)area * (CHUNKS * 8))[lane];
    // this warp's own (chunk, warp) totals, lane c = chunk c: which voxel rows phase A will read
    const uint32_t my_total = lane < CHUNKS ? warp_totals[(size_t)area * (CHUNKS * 8) + lane * 8 + wid] : 0u;
    // launched speculatively right behind the offsets scan: if the stream does not fit the arena
    // the whole grid backs off and the host, which sees the totals afterwards, enlarges it and
    // launches again
    if (stream_faces > cap_faces || stream_recs > cap_recs) return;
    const AreaView a = make_view(job, nb, pool_voxels, pool_planes);
    const uint32_t gx0 = job.x * AREA_SIZE, gy0 = job.y * AREA_SIZE;
    // bit c: this warp's group of chunk c has faces (lane c holds its total)
    const uint32_t busy_chunks = __ballot_sync(0xFFFFFFFFu, my_total != 0u);
    {   // The voxel rows of the busy groups are requested into L2 together with the planes: phase A
        // would otherwise start a third dependent round trip to HBM (32 rows = four 128-byte lines).
        uint32_t busy = busy_chunks;
        while (busy) {
            const int c = __ffs(busy) - 1;
            busy &= busy - 1;
            if ((lane & 7) == 0)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(a.vox + (size_t)(c * 256 + tid) * 16));
        }
    }
    load_plane_tile(sm.planes, a, tid);
    if (wid == 0) {  // running sums of the 64 (chunk, warp) totals, two per lane
        const unsigned long long t0 = widen(w2.x), t1 = widen(w2.y);
        const unsigned long long inc = warp_inclusive_scan64(t0 + t1, lane);
        sm.part[2 * lane] = inc - t0 - t1;
        sm.part[2 * lane + 1] = inc - t1;
        if (lane == 31) sm.part[CHUNKS * 8] = inc;
    }
    __syncthreads();

    const unsigned long long area_total = sm.part[CHUNKS * 8];
    const bool one_pass = rows_of(area_total) <= ROW_CAP && faces_of(area_total) <= LIST_FACES;
    int c0 = 0;
    while (c0 < CHUNKS) {
        // a pass = as many consecutive chunks as fit the row entries and the descriptor list
        const unsigned long long p0 = sm.part[c0 * 8];
        int c1 = one_pass ? CHUNKS : c0 + 1;
        while (c1 < CHUNKS) {
            const unsigned long long p = sm.part[(c1 + 1) * 8] - p0;  // field-wise: sums only grow
            if (rows_of(p) > ROW_CAP || faces_of(p) > LIST_FACES) break;
            ++c1;
        }
        const unsigned long long p1 = sm.part[c1 * 8];
        const uint32_t pass_lo = faces_of(p0), pass_hi = faces_of(p1), rows0 = rows_of(p0);
        const int nrows = (int)(rows_of(p1) - rows0);
        // a single chunk can hold up to 24 576 faces: then the list is filled in windows, and the
        // rows are staged again for each (the staging tiles of the flush overwrite them)
        const bool windowed = pass_hi - pass_lo > LIST_FACES;
        for (uint32_t lo = pass_lo; lo < pass_hi; lo += LIST_FACES) {
            const uint32_t hi = min(pass_hi, lo + LIST_FACES);
            // ---- A
#pragma unroll 1
            for (uint32_t todo = busy_chunks & ((1u << c1) - (1u << c0)); todo; todo &= todo - 1) {
                const int c = __ffs(todo) - 1;  // busy chunks of the pass only (uniform across the warp)
                const unsigned long long before = sm.part[c * 8 + wid];
                const int r = c * 256 + tid;
                const RowMasks k = row_masks(sm.planes, a, r);
                const uint32_t v = row_counts(k);
                const uint32_t ex = warp_inclusive_scan(v, lane) - v;  // rows of this warp before mine
                const uint32_t with_faces = __ballot_sync(0xFFFFFFFFu, v != 0u);
                if (v != 0u) {
                    const uint32_t e = rows_of(before) - rows0 + __popc(with_faces & ((1u << lane) - 1u));
                    sm.u.rows.vox[e] = reinterpret_cast<const uint4*>(a.vox)[r];
                    sm.u.rows.masks[e] = make_uint4(k.m[0] | (k.m[1] << 16), k.m[2] | (k.m[3] << 16),
                                                    k.m[4] | (k.m[5] << 16), (uint32_t)r);
                    sm.u.rows.first[e] = make_uint2(faces_of(before) + (ex & 0xFFFFu), recs_of(before) + (ex >> 16));
                }
            }
            __syncthreads();
            // ---- B: one thread per visible voxel of the pass (the record index), found in the staged
            //      rows by binary search over their record prefixes -- every lane has work, and the
            //      records leave as consecutive 16-byte stores
            const uint32_t rec0 = recs_of(p0), nrecs = recs_of(p1) - rec0;
#pragma unroll 1
            for (uint32_t q = tid; q < nrecs; q += 256) {
                const uint32_t target = rec0 + q;
                int elo = 0, ehi = nrows;  // largest entry whose first record is <= target
                while (ehi - elo > 1) {
                    const int mid = (elo + ehi) >> 1;
                    if (sm.u.rows.first[mid].y <= target) elo = mid; else ehi = mid;
                }
                const uint2 fq = sm.u.rows.first[elo];
                const uint4 m = sm.u.rows.masks[elo];
                const uint32_t vis = (m.x | (m.x >> 16) | m.y | (m.y >> 16) | m.z | (m.z >> 16)) & 0xFFFFu;
                const int x = (int)__fns(vis, 0, (int)(target - fq.y) + 1);  // the (k+1)-th visible voxel of the row
                const uint32_t below = (1u << x) - 1u, below2 = below | (below << 16);
                const uint32_t t0 = (m.x >> x) & 0x10001u, t1 = (m.y >> x) & 0x10001u, t2 = (m.z >> x) & 0x10001u;
                const uint32_t m6 = (t0 & 1u) | (t0 >> 15) | ((t1 & 1u) << 2) | (t1 >> 13) | ((t2 & 1u) << 4) | (t2 >> 11);
                const uint32_t f = fq.x + __popc(m.x & below2) + __popc(m.y & below2) + __popc(m.z & below2);
                const uint32_t voxel = reinterpret_cast<const uint8_t*>(&sm.u.rows.vox[elo])[x];
                const uint32_t r = m.w;
                if (!windowed || (f >= lo && f < hi))  // the window of the voxel's first face writes its record
                    out.recs[rec_base + target] =
                        make_uint4(gx0 + x, gy0 + (r & 15u), (r >> 4) | (voxel << 8) | (m6 << 16) | ((uint32_t)__popc(m6) << 24),
                                   face_base + f);
                const uint32_t desc0 = make_desc((int)r, x, 0, 0, voxel);
                uint32_t g = f - lo;  // list position of the voxel's next face; ordinal in bits 18..20
                uint32_t ord = 0;
#pragma unroll
                for (int d = 0; d < 6; ++d) {
                    if ((m6 >> d) & 1u) {
                        if (!windowed || g < hi - lo) sm.list[g] = desc0 | ((uint32_t)d << 15) | (ord << 18);
                        ++g;  // (g wraps below 0 for faces before the window: the unsigned compare rejects them)
                        ++ord;
                    }
                }
            }
            __syncthreads();
            // ---- C
            flush_faces(sm, hi - lo, face_base + lo, out, gx0, gy0, tid);
            if (c1 < CHUNKS || hi < pass_hi) __syncthreads();
        }
        c0 = c1;
    }
}

}  // namespace

void launch_mesh_prepare(const uint2* xy, int n, AreaMap map, uint4* jobs, uint4* nbr, uint32_t* counters,
                         uint2* missing, uint32_t missing_cap, cudaStream_t stream) {
    if (n <= 0) return;
    mesh_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(xy, n, map, jobs, nbr, counters, missing, missing_cap);
}

void launch_mesh_count(const uint4* jobs, const uint4* nbr, int n, const uint8_t* pool_voxels,
                       const uint16_t* pool_planes, uint32_t* rec_count,
                       uint32_t* face_count, uint32_t* warp_totals, cudaStream_t stream) {
    if (n <= 0) return;
    mesh_count_kernel<<<n, 256, 0, stream>>>(jobs, nbr, pool_voxels, pool_planes, rec_count,
                                             face_count, warp_totals);
}

void launch_mesh_offsets(const uint32_t* rec_count, const uint32_t* face_count, int n, uint32_t* rec_first,
                         uint32_t* face_first, const uint32_t* counters, uint32_t* host_totals, cudaStream_t stream) {
    mesh_offsets_kernel<<<1, 1024, 0, stream>>>(rec_count, face_count, n, rec_first, face_first, counters, host_totals);
}

void launch_mesh_emit(const uint4* jobs, const uint4* nbr, int n, const uint8_t* pool_voxels,
                      const uint16_t* pool_planes, const uint32_t* warp_totals, MeshOut out,
                      uint32_t cap_faces, uint32_t cap_recs, cudaStream_t stream) {
    if (n <= 0) return;
    mesh_emit_kernel<<<n, 256, 0, stream>>>(jobs, nbr, pool_voxels, pool_planes, warp_totals, out, n, cap_faces,
                                            cap_recs);
}

}  // namespace vx
