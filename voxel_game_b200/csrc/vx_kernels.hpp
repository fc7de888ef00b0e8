// vx_kernels.hpp -- launchers of the device kernels (implemented in vx_gen.cu, vx_mesh.cu,
// vx_edit.cu, vx_light.cu); called only from vx_api.cu.  All launches are asynchronous on
// `stream`; all pointers are device pointers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "vx_common.cuh"
#include "vx_tables.hpp"

namespace vx {

// ---- generation (K1, K1b, K2) and column heights (L1)
// jobs: n entries {area x, area y, slot, 0}
// One cave-zone column for K1b: slot * 256 + column.  (A 16-byte item that also carries the terrain
// height and the global x, y -- one load per claim instead of a chain of three -- was measured: no gain
// for K1b, K1 loses 1 % to the wider scattered stores; profiles/r02_ab_cave_claims.log.)
typedef uint32_t CaveItem;
void launch_gen_columns(const NoiseTables& nt, const uint4* jobs, int n, uint8_t* pool_voxels, uint8_t* pool_heights,
                        uint2* slot_xy, CaveItem* cave_items, uint32_t* cave_count,
                        unsigned long long* stats /* [5]: simplex2, alt simplex3, cave voxels, cave evals, cursor */,
                        uint16_t* pool_planes, uint32_t* cave_areas /* jobs K2 has to finish; counted at cave_count[1] */,
                        cudaStream_t stream);
void launch_gen_caves(const NoiseTables& nt, const CaveItem* cave_items, const uint32_t* cave_count,
                      uint8_t* pool_voxels, const uint8_t* pool_heights, const uint2* slot_xy,
                      unsigned long long* stats, int sm_count, cudaStream_t stream);
void launch_gen_finish(const uint4* jobs, const uint32_t* cave_areas, const uint32_t* cave_area_count, int n,
                       uint8_t* pool_voxels, uint8_t* pool_heights, uint16_t* pool_planes, uint64_t seed,
                       int sm_count, cudaStream_t stream);
// pool_planes may be null (temporary areas that are never meshed)
void launch_column_heights(const uint4* jobs, int n, const uint8_t* pool_voxels,
                           uint8_t* pool_heights, uint2* slot_xy, uint16_t* pool_planes,
                           cudaStream_t stream);

// ---- meshing (K4): count -> offsets -> emit
struct MeshOut {
    uint32_t* rec_first;   // n+1
    uint32_t* face_first;  // n+1
    uint4* recs;
    uint4* vertices;       // 10 uint4 per face
    uint32_t* indices;     // 3 u32 per face
};
// planes: per slot uint16 S[2048] ("voxel != None") then uint16 T[2048] ("transparent")
// The planes of every resident area are kept current by whoever writes its voxels (generation,
// upload, edits).
// prepare: xy -> jobs {ax, ay, slot, 0} and nbr {slots of the +x, -x, +y, -y neighbours};
// counters[0] = requested areas not resident, counters[1] = neighbours not resident (-> missing[])
void launch_mesh_prepare(const uint2* xy, int n, AreaMap map, uint4* jobs, uint4* nbr, uint32_t* counters,
                         uint2* missing, uint32_t missing_cap, cudaStream_t stream);
void launch_mesh_count(const uint4* jobs, const uint4* nbr, int n, const uint8_t* pool_voxels,
                       const uint16_t* pool_planes, uint32_t* rec_count,
                       uint32_t* face_count, uint32_t* warp_totals /* 64 per area */, cudaStream_t stream);
void launch_mesh_offsets(const uint32_t* rec_count, const uint32_t* face_count, int n, uint32_t* rec_first,
                         uint32_t* face_first, const uint32_t* counters, uint32_t* host_totals, cudaStream_t stream);
void launch_mesh_emit(const uint4* jobs, const uint4* nbr, int n, const uint8_t* pool_voxels,
                      const uint16_t* pool_planes, const uint32_t* warp_totals, MeshOut out,
                      uint32_t cap_faces, uint32_t cap_recs /* arena capacity; no-op if exceeded */,
                      cudaStream_t stream);

// ---- per-frame visibility (K7), vx_frame.cu
constexpr int FRAME_MAX_LAMPS = 64;
struct FrameView {
    float cx, cy, cz, target_z;  // camera position, camera target z
    float lx, ly, lz;            // look direction
    uint32_t render_size, head_in_water, show_map;
};
struct FrameTotals {  // zero n_faces / n_areas before the launch
    unsigned long long n_faces;
    uint32_t n_areas, n_lamps;
    uint32_t type_first[28];
    uint32_t lamps[FRAME_MAX_LAMPS];
};
// counts: 28 * n words; row_total: 28 words; done: one word, zero before the first launch
void launch_frame_visible(const uint4* recs, const uint32_t* rec_first, const uint2* area_xy, int n,
                          const FrameView& view, uint8_t* area_flag, uint32_t* counts, uint32_t* row_total,
                          uint32_t* done, FrameTotals* totals, uint32_t* visible, uint8_t* codes, cudaStream_t stream);

// the same frame over the device-resident mirror of Renderer.meshes (mesh retention): zone list of
// areas, per-slot face-mask volumes; visible[] entries are zone index << 15 | x + 16 y + 256 z
void launch_frame_visible_zone(const uint2* area_xy, int n, AreaMap map, const uint8_t* meshed,
                               const uint8_t* face_volume, const uint8_t* pool_voxels, const FrameView& view,
                               uint8_t* area_flag, int32_t* zone_slot, uint32_t* counts, uint32_t* row_total,
                               uint32_t* done, FrameTotals* totals, uint32_t* visible, uint32_t cap_visible,
                               cudaStream_t stream);

// ---- incremental re-meshing (M6), mesh retention, compact records: vx_update.cu
// hash_keys / hash_vals filled with 0xFF bytes; cand_slot: 7 n words; cand_info: 7 n half-words;
// rec_count / face_count: one word per 256 candidates
void launch_update_claim(const uint32_t* xyz, int n, AreaMap map, unsigned long long* hash_keys, uint32_t* hash_vals,
                         uint32_t hash_cap, int32_t* cand_slot, cudaStream_t stream);
void launch_update_count(const uint32_t* xyz, int n, AreaMap map, const uint8_t* pool_voxels,
                         const unsigned long long* hash_keys, const uint32_t* hash_vals, uint32_t hash_cap,
                         const int32_t* cand_slot, uint16_t* cand_info, uint32_t* rec_count, uint32_t* face_count,
                         int32_t* status, cudaStream_t stream);
void launch_update_emit(const uint32_t* xyz, int n, const int32_t* cand_slot, const uint16_t* cand_info,
                        const uint32_t* rec_first, const uint32_t* face_first, MeshOut out,
                        uint8_t* face_volume /* may be null */, uint8_t* meshed, cudaStream_t stream);
void launch_mesh_retain(const uint4* jobs, int n, const uint32_t* rec_first, const uint4* recs, uint8_t* face_volume,
                        uint8_t* meshed, cudaStream_t stream);
void launch_mesh_pack(const uint4* recs, uint64_t n, uint32_t* packed, cudaStream_t stream);
// patches: n pairs {index, -, -, -}, {entry} applied to the device copy of the area map
void launch_map_patch(const uint4* patches, int n, uint4* entries, cudaStream_t stream);

// ---- area save files (K8 encode, K9 decode), vx_codec.cu; jobs = {ax, ay, slot, 0}
constexpr uint32_t AREA_FILE_MAX_BYTES = 33300;  // 4 + 3 + 32768 literals + token / length bytes
// kinds: n * 2048 bytes of scratch written by the size pass and read by the emit pass
void launch_codec_sizes(const uint4* jobs, int n, const uint8_t* pool_voxels, uint8_t* kinds, uint32_t* sizes,
                        unsigned long long* offsets /* n + 1 */, cudaStream_t stream);
void launch_codec_emit(const uint4* jobs, int n, const uint8_t* pool_voxels, uint8_t* kinds,
                       const unsigned long long* offsets, uint8_t* out, unsigned long long cap_bytes /* backs off beyond */,
                       cudaStream_t stream);
void launch_codec_decode(const uint4* jobs, int n, const uint8_t* files, const unsigned long long* offsets,
                         uint8_t* pool_voxels, uint8_t* status /* n */, cudaStream_t stream);

// ---- edits (K5) and height map (K6)
// hash_keys must be filled with 0xFF bytes, hash_vals / dirty_flags / dirty_count / status with 0
void launch_apply_edits(const uint4* edits, int n, AreaMap map, uint8_t* pool_voxels,
                        uint8_t* pool_heights, uint16_t* pool_planes, unsigned long long* hash_keys, uint32_t* hash_vals,
                        uint32_t hash_cap, uint32_t* dirty_flags /* per slot */,
                        uint32_t* dirty_slots, uint32_t* dirty_count, int32_t* status,
                        cudaStream_t stream);
// explosions of one tick (bomb_simulator.rs:191-262); buffers as for launch_apply_edits, plus
// flags: n * 1331 bytes, removed / bombs: 3 * n * 1331 words each, counts: 2 words
constexpr uint32_t EXPLOSION_CANDIDATES = 1331;  // (2 * ceil(4.5) + 1)^3 voxels visited per explosion
void launch_explode(const float* positions, int n, AreaMap map, uint8_t* pool_voxels, uint8_t* pool_heights,
                    uint16_t* pool_planes, unsigned long long* hash_keys, uint32_t* hash_vals, uint32_t hash_cap,
                    uint32_t* dirty_flags, uint32_t* dirty_slots, uint32_t* dirty_count, int32_t* status,
                    uint8_t* flags, uint32_t* removed, uint32_t* bombs, uint32_t* counts, cudaStream_t stream);
void launch_heightmap(const uint2* area_xy, int n, AreaMap map, const uint8_t* pool_heights,
                      uint32_t cam_ax, uint32_t cam_ay, uint32_t* rgba, cudaStream_t stream);

// ---- L-ext
// light: n * 32768 bytes in job order; slot_to_job must be filled with -1 (0xFF bytes) beforehand
void launch_light_init(const uint4* jobs, int n, const uint8_t* pool_voxels,
                       const uint8_t* pool_heights, uint8_t* light, int32_t* slot_to_job,
                       cudaStream_t stream);
void launch_light_relax(const uint4* jobs, int n, const uint8_t* pool_voxels, AreaMap map,
                        const int32_t* slot_to_job, uint8_t* light, uint32_t* changed,
                        cudaStream_t stream);

// ---- diagnostics
// vx_sim.cu: the world-mutating loops of the water / falling-voxel simulators, one ordered list each
void launch_water_tick(const uint32_t* check, int n, AreaMap map, uint8_t* voxels, uint8_t* heights, uint16_t* planes,
                       uint32_t* changed, uint32_t cap_changed, uint32_t* next, uint32_t cap_next, uint32_t* counts,
                       int32_t* status, cudaStream_t stream);
void launch_falling_check(const uint32_t* check, int n, AreaMap map, uint8_t* voxels, uint8_t* heights, uint16_t* planes,
                          uint32_t* spawned, uint32_t cap_spawned, uint32_t* next, uint32_t cap_next, uint32_t* counts,
                          int32_t* status, cudaStream_t stream);
void launch_falling_land(const float* positions, const uint8_t* types, int n, AreaMap map, uint8_t* voxels,
                         uint8_t* heights, uint16_t* planes, uint8_t* keep, uint32_t* placed, uint32_t cap_placed,
                         uint32_t* next, uint32_t cap_next, uint32_t* counts, int32_t* status, cudaStream_t stream);
void launch_cave_columns(const NoiseTables& nt, const uint2* area_xy, int n, uint32_t* columns, uint32_t* voxels,
                         cudaStream_t stream);
void launch_fp64_rate(double* sink, int blocks, int iters, cudaStream_t stream);

}  // namespace vx
