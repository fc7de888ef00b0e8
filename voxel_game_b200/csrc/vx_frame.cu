// vx_frame.cu -- per-frame visibility (K7): which meshed areas and which of their voxel meshes the
// renderer draws this frame, grouped by voxel type, plus the lamps of the visible areas.
//
// Replaces, for the mesh of the last vx_mesh call,
//   Renderer::prepare_visible_areas / is_area_visible / calculate_area_render_distance
//     (renderer.rs:324-346,507-519,549-558),
//   Renderer::filter_visible_voxels / is_voxel_visible -- the per-frame rayon loop over every
//     visible voxel (renderer.rs:521-547,560-586),
//   Renderer::optimise_render_order (renderer.rs:349-361) and prepare_lights (:496-501).
// The input is the record stream already resident in HBM (16 B per visible voxel), so a frame costs
// two passes over it: count per (voxel type, area), scan, scatter.  HBM-bound, no tensor cores.
//
// f32 arithmetic: glam 0.27 scalar Vec3 -- dot = (x*x + y*y) + z*z, length = sqrt(dot),
// normalize_or_zero = v * (1 / length) when that reciprocal is finite and positive.  The library is
// compiled with -fmad=false and IEEE sqrt / division, so the verdicts equal a scalar CPU evaluation.
#include "vx_kernels.hpp"

namespace vx {
namespace {

// renderer.rs:36-39
constexpr float AREA_RENDER_THRESHOLD = 0.35f;
constexpr float LOOK_DOWN_RENDER_MULTIPLIER = 0.5f;
constexpr float VOXEL_RENDER_THRESHOLD = 0.71f;
constexpr float VOXEL_PROXIMITY_THRESHOLD = 5.5f;

constexpr int TYPES = V_VARIANTS;     // 27 groups of optimise_render_order
constexpr int AREAS_PER_CTA = 8;      // one warp per area

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return (ax * bx + ay * by) + az * bz;
}

// calculate_area_render_distance (renderer.rs:549-558)
__device__ __forceinline__ float area_render_distance(const FrameView& v) {
    const float dot_product = fabsf(dot3(v.lx, v.ly, v.lz, 0.0f, 0.0f, 1.0f));
    const float smooth = 1.0f - (1.0f - dot_product) * (1.0f - dot_product);
    const float render_distance = smooth * (float)(AREA_SIZE * v.render_size);
    return fmaxf(render_distance * LOOK_DOWN_RENDER_MULTIPLIER, (float)AREA_SIZE);
}

// is_area_visible (renderer.rs:324-346)
__device__ __forceinline__ bool area_visible(uint32_t ax, uint32_t ay, const FrameView& v, float render_distance) {
    const float mx = (float)((int)(ax * AREA_SIZE + AREA_SIZE / 2) - LOCATION_OFFSET);
    const float my = (float)((int)(ay * AREA_SIZE + AREA_SIZE / 2) - LOCATION_OFFSET);
    const float mz = v.target_z;
    const float dx = v.cx - mx, dy = v.cy - my, dz = v.cz - mz;  // position.distance(area_vec)
    if (sqrtf(dot3(dx, dy, dz, dx, dy, dz)) <= render_distance) return true;
    const float tx = mx - v.cx, ty = my - v.cy, tz = mz - v.cz;
    const float rcp = 1.0f / sqrtf(dot3(tx, ty, tz, tx, ty, tz));
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;  // normalize_or_zero
    if (isfinite(rcp) && rcp > 0.0f) {
        nx = tx * rcp;
        ny = ty * rcp;
        nz = tz * rcp;
    }
    return dot3(nx, ny, nz, v.lx, v.ly, v.lz) >= AREA_RENDER_THRESHOLD;
}

// filter_visible_voxels' predicate (renderer.rs:536-540) with is_voxel_visible (:560-586)
__device__ __forceinline__ bool voxel_drawn(const uint4 rec, const FrameView& v, float max_distance) {
    if (v.show_map) return true;  // renderer.rs:428-432: the map view draws everything
    const uint32_t voxel = (rec.z >> 8) & 0xFFu;
    if (v.head_in_water && voxel >= V_WATER_SOURCE && voxel <= V_WATER4) return false;  // Voxel::WATER
    const float dx = (float)((int)rec.x - LOCATION_OFFSET) - v.cx;
    const float dy = (float)((int)rec.y - LOCATION_OFFSET) - v.cy;
    const float dz = (float)(int)(rec.z & 0xFFu) - v.cz;
    if (fabsf(dx) <= VOXEL_PROXIMITY_THRESHOLD && fabsf(dy) <= VOXEL_PROXIMITY_THRESHOLD &&
        fabsf(dz) <= VOXEL_PROXIMITY_THRESHOLD)
        return true;
    const float distance = sqrtf(dot3(dx, dy, dz, dx, dy, dz));
    if (distance > max_distance) return false;
    return dot3(dx, dy, dz, v.lx, v.ly, v.lz) > VOXEL_RENDER_THRESHOLD * distance;
}

__global__ void frame_areas_kernel(const uint2* __restrict__ area_xy, int n, FrameView v,
                                   uint8_t* __restrict__ area_flag, FrameTotals* totals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    if (i < n) {
        const uint2 a = area_xy[i];
        vis = v.show_map || area_visible(a.x, a.y, v, area_render_distance(v));
        area_flag[i] = vis ? 1 : 0;
    }
    const uint32_t cnt = __popc(__ballot_sync(0xFFFFFFFFu, vis));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&totals->n_areas, cnt);
}

// One warp per area.  counts[t * n + i] = drawn records of type t in area i; counts[TYPES * n + i] =
// lamp records of area i (all of them, drawn or not: RenderArea.lights, renderer.rs:60-67).
__global__ void __launch_bounds__(32 * AREAS_PER_CTA)
frame_count_kernel(const uint4* __restrict__ recs, const uint32_t* __restrict__ rec_first, int n, FrameView v,
                   const uint8_t* __restrict__ area_flag, uint32_t* __restrict__ counts, FrameTotals* totals,
                   uint8_t* __restrict__ codes) {
    __shared__ uint32_t hist[AREAS_PER_CTA][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * AREAS_PER_CTA + wid;
    if (i >= n) return;
    hist[wid][lane] = 0;
    __syncwarp();
    uint32_t faces = 0, lamps = 0;
    if (area_flag[i]) {
        const float max_distance = (float)(v.render_size * AREA_SIZE);  // renderer.rs:530
        const uint32_t q0 = rec_first[i], q1 = rec_first[i + 1];
        for (uint32_t q = q0 + lane; q - lane < q1; q += 32) {
            const bool in = q < q1;
            const uint4 rec = in ? recs[q] : make_uint4(0, 0, 0, 0);
            const uint32_t voxel = (rec.z >> 8) & 0xFFu;
            const bool drawn = in && voxel_drawn(rec, v, max_distance);
            lamps += (in && voxel == V_LAMP) ? 1u : 0u;
            faces += drawn ? rec.z >> 24 : 0u;
            // what the scatter pass needs of this record, so that it reads one byte instead of sixteen
            // and repeats none of the arithmetic: voxel type | 0x80 if it is not drawn
            if (in) codes[q] = (uint8_t)(voxel | (drawn ? 0u : 0x80u));
            const uint32_t key = drawn ? voxel : 31u;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            if (drawn && lane == __ffs(peers) - 1) hist[wid][voxel] += __popc(peers);
            __syncwarp();
        }
    }
    if (lane < TYPES) counts[(size_t)lane * n + i] = hist[wid][lane];
    faces = __reduce_add_sync(0xFFFFFFFFu, faces);
    lamps = __reduce_add_sync(0xFFFFFFFFu, lamps);
    if (lane == 0) {
        counts[(size_t)TYPES * n + i] = lamps;
        if (faces) atomicAdd(&totals->n_faces, (unsigned long long)faces);
    }
}

// CTA t scans counts[t * n .. (t + 1) * n) in place (exclusive); the last CTA to finish turns the 28
// row totals into type_first[] and the lamp total.
__global__ void __launch_bounds__(1024) frame_scan_kernel(uint32_t* __restrict__ counts, int n, FrameTotals* totals,
                                                          uint32_t* __restrict__ row_total, uint32_t* done) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t carry;
    __shared__ bool last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* row = counts + (size_t)blockIdx.x * n;
    if (tid == 0) carry = 0;
    __syncthreads();
    constexpr int PER = 16;  // consecutive entries per thread: 16 384 areas are one round
    for (int base = 0; base < n; base += 1024 * PER) {
        const int i0 = base + tid * PER;
        uint32_t c[PER], mine = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            c[k] = i0 + k < n ? row[i0 + k] : 0u;
            mine += c[k];
        }
        uint32_t inc = mine;
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
        uint32_t run = carry + (wid ? s_w[wid - 1] : 0u) + inc - mine;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (i0 + k < n) row[i0 + k] = run;
            run += c[k];
        }
        __syncthreads();
        if (tid == 1023) carry = run;
        __syncthreads();
    }
    if (tid == 0) {
        row_total[blockIdx.x] = carry;
        __threadfence();
        last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && tid == 0) {
        __threadfence();
        uint32_t run = 0;
        for (int t = 0; t < TYPES; ++t) {
            totals->type_first[t] = run;
            run += reinterpret_cast<volatile uint32_t*>(row_total)[t];
        }
        totals->type_first[TYPES] = run;
        totals->n_lamps = reinterpret_cast<volatile uint32_t*>(row_total)[TYPES];
        *done = 0;  // ready for the next frame
    }
}

// The walk of the count kernel over its one-byte codes; every drawn record's index goes to its slot:
// type group, then area order, then record order (deterministic; the reference's order inside a group
// is HashMap order).
__global__ void __launch_bounds__(32 * AREAS_PER_CTA)
frame_scatter_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ rec_first, int n,
                     const uint8_t* __restrict__ area_flag, const uint32_t* __restrict__ counts,
                     FrameTotals* totals, uint32_t* __restrict__ visible) {
    __shared__ uint32_t cursor[AREAS_PER_CTA][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * AREAS_PER_CTA + wid;
    if (i >= n || !area_flag[i]) return;
    if (lane < TYPES) cursor[wid][lane] = totals->type_first[lane] + counts[(size_t)lane * n + i];
    uint32_t lamp_cursor = counts[(size_t)TYPES * n + i];
    __syncwarp();
    const uint32_t q0 = rec_first[i], q1 = rec_first[i + 1];
    for (uint32_t q = q0 + lane; q - lane < q1; q += 32) {
        const bool in = q < q1;
        const uint32_t code = in ? codes[q] : 0xFFu;
        const uint32_t voxel = code & 0x7Fu;
        const bool drawn = in && !(code & 0x80u);
        const uint32_t key = drawn ? voxel : 31u;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (drawn) base = cursor[wid][voxel];
        __syncwarp();
        if (drawn) {
            visible[base + rank] = q;
            if (rank == 0) cursor[wid][voxel] = base + __popc(peers);
        }
        const uint32_t lamp_mask = __ballot_sync(0xFFFFFFFFu, in && voxel == V_LAMP);
        if (in && voxel == V_LAMP) {
            const uint32_t k = lamp_cursor + __popc(lamp_mask & ((1u << lane) - 1u));
            if (k < FRAME_MAX_LAMPS) totals->lamps[k] = q;  // voxel_shader.rs:21: at most 64 are used
        }
        lamp_cursor += __popc(lamp_mask);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------- resident zone
// The same frame over the device-resident mirror of Renderer.meshes (vx_update.cu "retention"): the
// input is, per area of the render zone, the 32 KiB face-mask volume (0 = no mesh entry) and the
// voxel bytes, so nothing has to be re-meshed when the zone slides or an edit lands.  One warp per
// area walks the volume 32 rows at a time (one 16-byte row per lane); rows that hold meshes are then
// taken two at a time, one voxel per lane, in index order.  A drawn voxel is reported as
// zone index << 15 | x + 16 y + 256 z.
template <class Fn>
__device__ __forceinline__ void walk_zone_area(const uint8_t* __restrict__ vol, const uint8_t* __restrict__ vox,
                                               uint32_t ax, uint32_t ay, const FrameView& v, int lane, Fn&& fn) {
    const float max_distance = (float)(v.render_size * AREA_SIZE);  // renderer.rs:530
    const uint4* vol4 = reinterpret_cast<const uint4*>(vol);
    for (int batch = 0; batch < VOXELS_IN_AREA / 16 / 32; ++batch) {
        const uint4 f = vol4[batch * 32 + lane];
        uint32_t active = __ballot_sync(0xFFFFFFFFu, (f.x | f.y | f.z | f.w) != 0u);
        while (active) {
            const int j0 = __ffs(active) - 1;
            active &= active - 1;
            int j1 = -1;
            if (active) {
                j1 = __ffs(active) - 1;
                active &= active - 1;
            }
            const int my_row = lane < 16 ? j0 : j1;
            const uint32_t x = lane & 15;
            uint32_t mask = 0, voxel = 0, local = 0;
            if (my_row >= 0) {
                local = (uint32_t)(batch * 32 + my_row) * 16u + x;
                mask = vol[local];
                voxel = vox[local];
            }
            const bool in = mask != 0u;
            // the record the batch mesher would hold for this voxel (gx, gy, gz | voxel << 8 | ...)
            const uint4 rec = make_uint4(ax * AREA_SIZE + x, ay * AREA_SIZE + ((local >> 4) & 15u),
                                         (local >> 8) | (voxel << 8), 0u);
            const bool drawn = in && voxel_drawn(rec, v, max_distance);
            fn(in, drawn, voxel, mask, local);
        }
    }
}

// prepare_visible_areas iterates over `meshes` (renderer.rs:507-519): an area of the zone list counts
// only if it is resident and meshed.  zone_slot[i] = its pool slot when it is drawn from, else -1.
__global__ void zone_areas_kernel(const uint2* __restrict__ area_xy, int n, AreaMap map,
                                  const uint8_t* __restrict__ meshed, FrameView v, uint8_t* __restrict__ area_flag,
                                  int32_t* __restrict__ zone_slot, FrameTotals* totals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    if (i < n) {
        const uint2 a = area_xy[i];
        const int slot = area_slot(map, a.x, a.y);
        vis = slot >= 0 && meshed[slot] && (v.show_map || area_visible(a.x, a.y, v, area_render_distance(v)));
        area_flag[i] = vis ? 1 : 0;
        zone_slot[i] = vis ? slot : -1;
    }
    const uint32_t cnt = __popc(__ballot_sync(0xFFFFFFFFu, vis));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&totals->n_areas, cnt);
}

__global__ void __launch_bounds__(32 * AREAS_PER_CTA)
zone_count_kernel(const uint2* __restrict__ area_xy, int n, const int32_t* __restrict__ zone_slot,
                  const uint8_t* __restrict__ face_volume, const uint8_t* __restrict__ pool_voxels, FrameView v,
                  uint32_t* __restrict__ counts, FrameTotals* totals) {
    __shared__ uint32_t hist[AREAS_PER_CTA][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * AREAS_PER_CTA + wid;
    if (i >= n) return;
    hist[wid][lane] = 0;
    __syncwarp();
    uint32_t faces = 0, lamps = 0;
    const uint2 a = area_xy[i];
    const int slot = zone_slot[i];
    if (slot >= 0) {
        walk_zone_area(face_volume + (size_t)slot * VOXELS_IN_AREA, pool_voxels + (size_t)slot * VOXELS_IN_AREA, a.x, a.y,
                       v, lane, [&](bool in, bool drawn, uint32_t voxel, uint32_t mask, uint32_t) {
                           lamps += (in && voxel == V_LAMP) ? 1u : 0u;
                           faces += drawn ? __popc(mask) : 0u;
                           const uint32_t key = drawn ? voxel : 31u;
                           const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
                           if (drawn && lane == __ffs(peers) - 1) hist[wid][voxel] += __popc(peers);
                           __syncwarp();
                       });
    }
    if (lane < TYPES) counts[(size_t)lane * n + i] = hist[wid][lane];
    faces = __reduce_add_sync(0xFFFFFFFFu, faces);
    lamps = __reduce_add_sync(0xFFFFFFFFu, lamps);
    if (lane == 0) {
        counts[(size_t)TYPES * n + i] = lamps;
        if (faces) atomicAdd(&totals->n_faces, (unsigned long long)faces);
    }
}

__global__ void __launch_bounds__(32 * AREAS_PER_CTA)
zone_scatter_kernel(const uint2* __restrict__ area_xy, int n, const int32_t* __restrict__ zone_slot,
                    const uint8_t* __restrict__ face_volume, const uint8_t* __restrict__ pool_voxels, FrameView v,
                    const uint32_t* __restrict__ counts, FrameTotals* totals, uint32_t* __restrict__ visible,
                    uint32_t cap_visible) {
    __shared__ uint32_t cursor[AREAS_PER_CTA][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * AREAS_PER_CTA + wid;
    if (i >= n) return;
    const uint2 a = area_xy[i];
    const int slot = zone_slot[i];
    if (slot < 0) return;
    if (lane < TYPES) cursor[wid][lane] = totals->type_first[lane] + counts[(size_t)lane * n + i];
    uint32_t lamp_cursor = counts[(size_t)TYPES * n + i];
    __syncwarp();
    walk_zone_area(face_volume + (size_t)slot * VOXELS_IN_AREA, pool_voxels + (size_t)slot * VOXELS_IN_AREA, a.x, a.y, v,
                   lane, [&](bool in, bool drawn, uint32_t voxel, uint32_t, uint32_t local) {
                       const uint32_t code = ((uint32_t)i << 15) | local;
                       const uint32_t key = drawn ? voxel : 31u;
                       const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
                       const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                       uint32_t base = 0;
                       if (drawn) base = cursor[wid][voxel];
                       __syncwarp();
                       if (drawn) {
                           if (base + rank < cap_visible) visible[base + rank] = code;  // the host repeats the frame with a larger list
                           if (rank == 0) cursor[wid][voxel] = base + __popc(peers);
                       }
                       const uint32_t lamp_mask = __ballot_sync(0xFFFFFFFFu, in && voxel == V_LAMP);
                       if (in && voxel == V_LAMP) {
                           const uint32_t k = lamp_cursor + __popc(lamp_mask & ((1u << lane) - 1u));
                           if (k < FRAME_MAX_LAMPS) totals->lamps[k] = code;
                       }
                       lamp_cursor += __popc(lamp_mask);
                       __syncwarp();
                   });
}

}  // namespace

void launch_frame_visible(const uint4* recs, const uint32_t* rec_first, const uint2* area_xy, int n,
                          const FrameView& view, uint8_t* area_flag, uint32_t* counts, uint32_t* row_total,
                          uint32_t* done, FrameTotals* totals, uint32_t* visible, uint8_t* codes, cudaStream_t stream) {
    if (n <= 0) return;
    frame_areas_kernel<<<(n + 255) / 256, 256, 0, stream>>>(area_xy, n, view, area_flag, totals);
    const int ctas = (n + AREAS_PER_CTA - 1) / AREAS_PER_CTA;
    frame_count_kernel<<<ctas, 32 * AREAS_PER_CTA, 0, stream>>>(recs, rec_first, n, view, area_flag, counts, totals, codes);
    frame_scan_kernel<<<TYPES + 1, 1024, 0, stream>>>(counts, n, totals, row_total, done);
    frame_scatter_kernel<<<ctas, 32 * AREAS_PER_CTA, 0, stream>>>(codes, rec_first, n, area_flag, counts, totals, visible);
}

}  // namespace vx

namespace vx {
void launch_frame_visible_zone(const uint2* area_xy, int n, AreaMap map, const uint8_t* meshed,
                               const uint8_t* face_volume, const uint8_t* pool_voxels, const FrameView& view,
                               uint8_t* area_flag, int32_t* zone_slot, uint32_t* counts, uint32_t* row_total,
                               uint32_t* done, FrameTotals* totals, uint32_t* visible, uint32_t cap_visible,
                               cudaStream_t stream) {
    if (n <= 0) return;
    zone_areas_kernel<<<(n + 255) / 256, 256, 0, stream>>>(area_xy, n, map, meshed, view, area_flag, zone_slot, totals);
    const int ctas = (n + AREAS_PER_CTA - 1) / AREAS_PER_CTA;
    zone_count_kernel<<<ctas, 32 * AREAS_PER_CTA, 0, stream>>>(area_xy, n, zone_slot, face_volume, pool_voxels, view,
                                                               counts, totals);
    frame_scan_kernel<<<TYPES + 1, 1024, 0, stream>>>(counts, n, totals, row_total, done);
    zone_scatter_kernel<<<ctas, 32 * AREAS_PER_CTA, 0, stream>>>(area_xy, n, zone_slot, face_volume, pool_voxels, view,
                                                                 counts, totals, visible, cap_visible);
}
}  // namespace vx
