// vx_update.cu -- incremental re-meshing after edits (M6), the device-resident mirror of
// Renderer.meshes, and the compact record transport.
//
// Renderer::update_location (renderer.rs:239-286) re-runs generate_mesh_for_voxel
// (renderer.rs:126-195) + set_voxel_mesh (:197-219) for the edited voxel and its <= 6 neighbours,
// WITHOUT the enclosed-voxel pre-filter of a full-area load (world.rs:137-139): an interior
// partial-height water voxel that a full load drops gets its Up face the first time an edit touches
// it (should_generate_top_face, mesh_generator.rs:515-518).  For a batch of locations applied one by
// one, every touched voxel is re-meshed at or after the last change in its 6-neighbourhood (a change
// of a neighbour touches it again), so its final mesh is generate_mesh_for_voxel on the FINAL voxel
// state; voxels no location touches keep whatever mesh they had.  Hence the batch:
//   claim   candidate c = 7 * i + k (k: the location, x+1, x-1, y+1, y-1, z-1, z+1 -- the reference's
//           order) claims its voxel in a hash table; the lowest c per voxel owns it, so the output
//           order is the order in which the reference first touches each voxel;
//   count   owners evaluate the six face tests on the final voxels -> faces per 256-candidate block;
//   scan    (mesh_offsets_kernel of vx_mesh.cu over the blocks);
//   emit    one 16-byte record per touched voxel (n_faces = 0: RenderArea::remove) and its vertex /
//           index blocks; with mesh retention on, the resident face-mask volume is patched too.
//
// Mesh retention (the device mirror of `meshes: HashMap<AreaLocation, RenderArea>`,
// renderer.rs:97-101): per pool slot one byte per voxel = its face mask, 0 = no entry in
// RenderArea.mesh_map; `meshed[slot]` = the area has an entry in `meshes`.  A full-area mesh
// overwrites the slot's volume from the records it emitted; updates patch single bytes.
#include "vx_face.cuh"
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

// candidate k of location (x, y, z): false when it does not exist (z out of range)
__device__ __forceinline__ bool candidate_location(const uint32_t* __restrict__ xyz, uint32_t c, uint3& loc) {
    const uint32_t i = c / 7u, k = c - 7u * i;
    loc = make_uint3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    switch (k) {
    case 1: loc.x += 1u; break;
    case 2: loc.x -= 1u; break;
    case 3: loc.y += 1u; break;
    case 4: loc.y -= 1u; break;
    case 5:
        if (loc.z == 0u) return false;  // renderer.rs:265
        loc.z -= 1u;
        break;
    case 6:
        if (loc.z >= (uint32_t)AREA_HEIGHT - 1u) return false;  // renderer.rs:272
        loc.z += 1u;
        break;
    default: break;
    }
    return true;
}

__device__ __forceinline__ unsigned long long voxel_key(int slot, uint3 loc) {
    return (unsigned long long)(uint32_t)slot * VOXELS_IN_AREA +
           ((loc.x % AREA_SIZE) + (loc.y % AREA_SIZE) * AREA_SIZE + loc.z * 256u);
}

// owner[c] = pool slot of the candidate's area when the candidate exists and its area is resident
// NOW (world.get_without_loading, renderer.rs:241,281), else -1.  Evaluated before the call makes
// further areas resident for the face tests.
__global__ void update_claim_kernel(const uint32_t* __restrict__ xyz, uint32_t n_cand, AreaMap map,
                                    unsigned long long* hash_keys, uint32_t* hash_vals, uint32_t cap_mask,
                                    int32_t* __restrict__ cand_slot) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand) return;
    uint3 loc;
    int slot = -1;
    if (candidate_location(xyz, c, loc)) slot = area_slot(map, loc.x / AREA_SIZE, loc.y / AREA_SIZE);
    cand_slot[c] = slot;
    if (slot < 0) return;
    const unsigned long long key = voxel_key(slot, loc);
    uint32_t pos = hash_u64(key, cap_mask);
    while (true) {
        const unsigned long long prev = atomicCAS(&hash_keys[pos], EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) break;
        pos = (pos + 1) & cap_mask;
    }
    atomicMin(&hash_vals[pos], c);  // the first toucher owns the voxel
}

// World::get_with_cache for a face test: -1 when the area is not resident (the call generates such
// areas beforehand, as get_with_cache -> load_area does, so this means an internal error)
__device__ __forceinline__ int voxel_at(const uint8_t* __restrict__ pool_voxels, const AreaMap& map, uint32_t gx,
                                        uint32_t gy, uint32_t gz) {
    const int slot = area_slot(map, gx / AREA_SIZE, gy / AREA_SIZE);
    if (slot < 0) return -1;
    return pool_voxels[(size_t)slot * VOXELS_IN_AREA + (gx % AREA_SIZE) + (gy % AREA_SIZE) * AREA_SIZE + gz * 256u];
}

// generate_mesh_for_voxel (renderer.rs:126-195): face(cur, nbr) = cur != nbr && nbr transparent
// (the algebraic form of mesh_generator.rs:501-513, all 27 x 27 pairs checked in the tests);
// bit order = push order Left(x+1) Right(x-1) Front(y+1) Back(y-1) Down(z+1) Up(z-1).
// Returns 0x80 when a lateral neighbour's area is not resident.
__device__ __forceinline__ uint32_t update_face_mask(const uint8_t* __restrict__ pool_voxels, const AreaMap& map,
                                                     const uint8_t* __restrict__ own, uint3 loc, uint32_t cur) {
    if (cur == V_NONE) return 0u;
    const uint32_t lx = loc.x % AREA_SIZE, ly = loc.y % AREA_SIZE;
    const uint32_t local = lx + ly * AREA_SIZE + loc.z * 256u;
    uint32_t mask = 0u;
    int nb[4];
    nb[0] = lx < 15u ? own[local + 1] : voxel_at(pool_voxels, map, loc.x + 1u, loc.y, loc.z);
    nb[1] = lx > 0u ? own[local - 1] : voxel_at(pool_voxels, map, loc.x - 1u, loc.y, loc.z);
    nb[2] = ly < 15u ? own[local + 16] : voxel_at(pool_voxels, map, loc.x, loc.y + 1u, loc.z);
    nb[3] = ly > 0u ? own[local - 16] : voxel_at(pool_voxels, map, loc.x, loc.y - 1u, loc.z);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        if (nb[d] < 0) return 0x80u;
        if ((uint32_t)nb[d] != cur && is_transparent((uint32_t)nb[d])) mask |= 1u << d;
    }
    if (loc.z + 1u < (uint32_t)AREA_HEIGHT) {
        const uint32_t n = own[local + 256];
        if (n != cur && is_transparent(n)) mask |= 16u;
    }
    if (loc.z > 0u) {
        const uint32_t n = own[local - 256];
        if ((n != cur && is_transparent(n)) || partial_height(cur)) mask |= 32u;  // should_generate_top_face
    }
    return mask;
}

// cand_info[c] = 0 (not an owner) or 0x8000 | voxel << 8 | mask; per block: owners and faces
__global__ void __launch_bounds__(256) update_count_kernel(const uint32_t* __restrict__ xyz, uint32_t n_cand,
                                                           AreaMap map, const uint8_t* __restrict__ pool_voxels,
                                                           const unsigned long long* __restrict__ hash_keys,
                                                           const uint32_t* __restrict__ hash_vals, uint32_t cap_mask,
                                                           const int32_t* __restrict__ cand_slot,
                                                           uint16_t* __restrict__ cand_info,
                                                           uint32_t* __restrict__ rec_count,
                                                           uint32_t* __restrict__ face_count, int32_t* status) {
    __shared__ uint32_t s_r[8], s_f[8];
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    uint32_t recs = 0, faces = 0;
    if (c < n_cand) {
        uint32_t info = 0;
        const int slot = cand_slot[c];
        uint3 loc;
        if (slot >= 0 && candidate_location(xyz, c, loc)) {
            const unsigned long long key = voxel_key(slot, loc);
            uint32_t pos = hash_u64(key, cap_mask);
            while (hash_keys[pos] != key) pos = (pos + 1) & cap_mask;
            if (hash_vals[pos] == c) {
                const uint8_t* own = pool_voxels + (size_t)slot * VOXELS_IN_AREA;
                const uint32_t cur = own[(loc.x % AREA_SIZE) + (loc.y % AREA_SIZE) * AREA_SIZE + loc.z * 256u];
                uint32_t mask = update_face_mask(pool_voxels, map, own, loc, cur);
                if (mask & 0x80u) {
                    atomicMin(status, -5);  // VX_ERR_NOT_LOADED
                    mask = 0u;
                }
                info = 0x8000u | (cur << 8) | mask;
                recs = 1;
                faces = __popc(mask);
            }
        }
        cand_info[c] = (uint16_t)info;
    }
    recs = __reduce_add_sync(0xFFFFFFFFu, recs);
    faces = __reduce_add_sync(0xFFFFFFFFu, faces);
    if ((threadIdx.x & 31) == 0) {
        s_r[threadIdx.x >> 5] = recs;
        s_f[threadIdx.x >> 5] = faces;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t r = 0, f = 0;
        for (int i = 0; i < 8; ++i) {
            r += s_r[i];
            f += s_f[i];
        }
        rec_count[blockIdx.x] = r;
        face_count[blockIdx.x] = f;
    }
}

__global__ void __launch_bounds__(256) update_emit_kernel(const uint32_t* __restrict__ xyz, uint32_t n_cand,
                                                          const int32_t* __restrict__ cand_slot,
                                                          const uint16_t* __restrict__ cand_info,
                                                          const uint32_t* __restrict__ rec_first,
                                                          const uint32_t* __restrict__ face_first, MeshOut out,
                                                          uint8_t* __restrict__ face_volume,
                                                          uint8_t* __restrict__ meshed) {
    __shared__ uint32_t s_w[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    const uint32_t info = c < n_cand ? cand_info[c] : 0u;
    const uint32_t mask = info & 0x3Fu, nf = __popc(mask);
    const uint32_t v = (info >> 15) | (nf << 16);  // records | faces << 16
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint32_t before = inc - v;
    for (int w = 0; w < wid; ++w) before += s_w[w];
    if (!(info & 0x8000u)) return;
    uint3 loc;
    candidate_location(xyz, c, loc);
    const uint32_t voxel = (info >> 8) & 0x7Fu;
    const uint32_t rec = rec_first[blockIdx.x] + (before & 0xFFFFu);
    const uint32_t face0 = face_first[blockIdx.x] + (before >> 16);
    out.recs[rec] = make_uint4(loc.x, loc.y, loc.z | (voxel << 8) | (mask << 16) | (nf << 24), face0);
    const uint32_t lx = loc.x % AREA_SIZE, ly = loc.y % AREA_SIZE;
    if (face_volume) {  // set_voxel_mesh: insert or remove; a missing RenderArea is created (renderer.rs:208-212)
        const int slot = cand_slot[c];
        face_volume[(size_t)slot * VOXELS_IN_AREA + lx + ly * AREA_SIZE + loc.z * 256u] = (uint8_t)mask;
        meshed[slot] = 1;
    }
    uint32_t ord = 0;
#pragma unroll 1
    for (int d = 0; d < 6; ++d) {
        if (!((mask >> d) & 1u)) continue;
        const FaceParts p = decode_face(make_desc((int)(ly + 16u * loc.z), (int)lx, d, 0, voxel), loc.x - lx, loc.y - ly);
        uint4* dst = out.vertices + (size_t)(face0 + ord) * 10;
        build_half_face(dst, p, 0);
        build_half_face(dst + 5, p, 1);
        const uint32_t b = ord * 4u;
        uint32_t* w = out.indices + (size_t)(face0 + ord) * 3;
        w[0] = b | ((b + 1u) << 16);
        w[1] = (b + 2u) | (b << 16);
        w[2] = (b + 2u) | ((b + 3u) << 16);
        ++ord;
    }
}

// ------------------------------------------------------------------------------- retention
// After a full-area mesh: the slot's face-mask volume = exactly the records just emitted.
__global__ void __launch_bounds__(256) mesh_retain_kernel(const uint4* __restrict__ jobs,
                                                          const uint32_t* __restrict__ rec_first,
                                                          const uint4* __restrict__ recs,
                                                          uint8_t* __restrict__ face_volume,
                                                          uint8_t* __restrict__ meshed) {
    const int slot = (int)jobs[blockIdx.x].z;
    uint8_t* vol = face_volume + (size_t)slot * VOXELS_IN_AREA;
    uint4* v4 = reinterpret_cast<uint4*>(vol);
#pragma unroll
    for (int k = 0; k < VOXELS_IN_AREA / 16 / 256; ++k) v4[k * 256 + threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const uint32_t q0 = rec_first[blockIdx.x], q1 = rec_first[blockIdx.x + 1];
    for (uint32_t q = q0 + threadIdx.x; q < q1; q += 256) {
        const uint4 r = recs[q];
        vol[(r.x & 15u) + (r.y & 15u) * AREA_SIZE + (r.z & 0xFFu) * 256u] = (uint8_t)((r.z >> 16) & 0x3Fu);
    }
    if (threadIdx.x == 0) meshed[slot] = 1;
}

// compact record: local index (15) | voxel (5) << 15 | face mask (6) << 20 -- everything the 16-byte
// record and the vertex / index blocks are a pure function of, given the area's location
__global__ void mesh_pack_kernel(const uint4* __restrict__ recs, uint32_t n, uint32_t* __restrict__ packed) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint4 r = recs[q];
    packed[q] = (r.x & 15u) | ((r.y & 15u) << 4) | ((r.z & 0xFFu) << 8) | (((r.z >> 8) & 0x1Fu) << 15) |
                (((r.z >> 16) & 0x3Fu) << 20);
}

__global__ void map_patch_kernel(const uint4* __restrict__ patches, int n, uint4* __restrict__ entries) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) entries[patches[2 * i].x] = patches[2 * i + 1];
}

}  // namespace

void launch_map_patch(const uint4* patches, int n, uint4* entries, cudaStream_t stream) {
    if (n <= 0) return;
    map_patch_kernel<<<(n + 255) / 256, 256, 0, stream>>>(patches, n, entries);
}

void launch_update_claim(const uint32_t* xyz, int n, AreaMap map, unsigned long long* hash_keys, uint32_t* hash_vals,
                         uint32_t hash_cap, int32_t* cand_slot, cudaStream_t stream) {
    if (n <= 0) return;
    const uint32_t n_cand = 7u * (uint32_t)n;
    update_claim_kernel<<<(n_cand + 255) / 256, 256, 0, stream>>>(xyz, n_cand, map, hash_keys, hash_vals, hash_cap - 1,
                                                                 cand_slot);
}

void launch_update_count(const uint32_t* xyz, int n, AreaMap map, const uint8_t* pool_voxels,
                         const unsigned long long* hash_keys, const uint32_t* hash_vals, uint32_t hash_cap,
                         const int32_t* cand_slot, uint16_t* cand_info, uint32_t* rec_count, uint32_t* face_count,
                         int32_t* status, cudaStream_t stream) {
    if (n <= 0) return;
    const uint32_t n_cand = 7u * (uint32_t)n;
    update_count_kernel<<<(n_cand + 255) / 256, 256, 0, stream>>>(xyz, n_cand, map, pool_voxels, hash_keys, hash_vals,
                                                                 hash_cap - 1, cand_slot, cand_info, rec_count,
                                                                 face_count, status);
}

void launch_update_emit(const uint32_t* xyz, int n, const int32_t* cand_slot, const uint16_t* cand_info,
                        const uint32_t* rec_first, const uint32_t* face_first, MeshOut out, uint8_t* face_volume,
                        uint8_t* meshed, cudaStream_t stream) {
    if (n <= 0) return;
    const uint32_t n_cand = 7u * (uint32_t)n;
    update_emit_kernel<<<(n_cand + 255) / 256, 256, 0, stream>>>(xyz, n_cand, cand_slot, cand_info, rec_first,
                                                                face_first, out, face_volume, meshed);
}

void launch_mesh_retain(const uint4* jobs, int n, const uint32_t* rec_first, const uint4* recs, uint8_t* face_volume,
                        uint8_t* meshed, cudaStream_t stream) {
    if (n <= 0) return;
    mesh_retain_kernel<<<n, 256, 0, stream>>>(jobs, rec_first, recs, face_volume, meshed);
}

void launch_mesh_pack(const uint4* recs, uint64_t n, uint32_t* packed, cudaStream_t stream) {
    if (n == 0) return;
    mesh_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(recs, (uint32_t)n, packed);
}

}  // namespace vx
