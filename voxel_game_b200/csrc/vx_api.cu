// vx_api.cu -- the C ABI of include/voxel_cuda.h: context, resident-area pool, area hash map,
// kernel sequencing, host<->device copies.  No torch, no CPU compute path: a context cannot be
// created without a CUDA device, and every voxel / height / vertex is produced by a kernel.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>  // types and prototypes of the virtual-memory API; the entry points are fetched at run time

#include "../../include/voxel_cuda.h"
#include "vx_kernels.hpp"

using namespace vx;

namespace {

// ---- growable pool arrays -----------------------------------------------------------------------
// The resident-area pool (voxels, heights, planes, slot_xy, face masks) grows while a world streams in.
// Growing by allocate-copy-free needs twice the pool for a moment (10 GiB at config 5), a device sync,
// and moves every pointer.  Instead each array reserves a range of VIRTUAL addresses once (64 GiB of
// voxels = 2 M areas, nothing is committed) and physical memory is mapped behind what is in use, chunk
// by chunk (CUDA virtual memory management: cuMemAddressReserve / cuMemCreate / cuMemMap): growth
// copies nothing, waits for nothing, and the base pointers never change.  The driver entry points are
// resolved through the runtime (cudaGetDriverEntryPoint), so the library does not link libcuda.
struct VmApi {
    decltype(&cuMemGetAllocationGranularity) granularity = nullptr;
    decltype(&cuMemAddressReserve) reserve = nullptr;
    decltype(&cuMemAddressFree) address_free = nullptr;
    decltype(&cuMemCreate) create = nullptr;
    decltype(&cuMemRelease) release = nullptr;
    decltype(&cuMemMap) map = nullptr;
    decltype(&cuMemUnmap) unmap = nullptr;
    decltype(&cuMemSetAccess) set_access = nullptr;
    bool ok = false;
    bool load() {
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess &&
                   q == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        ok = get("cuMemGetAllocationGranularity", (void**)&granularity) && get("cuMemAddressReserve", (void**)&reserve) &&
             get("cuMemAddressFree", (void**)&address_free) && get("cuMemCreate", (void**)&create) &&
             get("cuMemRelease", (void**)&release) && get("cuMemMap", (void**)&map) && get("cuMemUnmap", (void**)&unmap) &&
             get("cuMemSetAccess", (void**)&set_access);
        if (!ok) cudaGetLastError();
        return ok;
    }
};

struct VmArray {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, chunk = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    template <class T> T* as() const { return reinterpret_cast<T*>(base); }
    // chunk_hint: physical memory is committed in pieces of about this size (a multiple of the granularity)
    bool reserve(const VmApi& api, int device, size_t max_bytes, size_t chunk_hint) {
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        size_t gran = 0;
        if (api.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return false;
        chunk = ((std::max(chunk_hint, gran) + gran - 1) / gran) * gran;
        reserved = ((max_bytes + chunk - 1) / chunk) * chunk;
        if (api.reserve(&base, reserved, 0, 0, 0) != CUDA_SUCCESS) {
            base = 0;
            reserved = 0;
            return false;
        }
        return true;
    }
    // at least `bytes` are backed by physical memory afterwards; new memory is not initialised
    cudaError_t grow(const VmApi& api, int device, size_t bytes) {
        if (bytes > reserved) return cudaErrorMemoryAllocation;
        while (mapped < bytes) {
            CUmemAllocationProp prop = {};
            prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
            prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            prop.location.id = device;
            CUmemGenericAllocationHandle h;
            if (api.create(&h, chunk, &prop, 0) != CUDA_SUCCESS) return cudaErrorMemoryAllocation;
            if (api.map(base + mapped, chunk, 0, h, 0) != CUDA_SUCCESS) {
                api.release(h);
                return cudaErrorMemoryAllocation;
            }
            CUmemAccessDesc access = {};
            access.location = prop.location;
            access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            if (api.set_access(base + mapped, chunk, &access, 1) != CUDA_SUCCESS) {
                api.unmap(base + mapped, chunk);
                api.release(h);
                return cudaErrorMemoryAllocation;
            }
            handles.push_back(h);
            mapped += chunk;
        }
        return cudaSuccess;
    }
    void unmap_all(const VmApi& api) {  // the reservation stays
        for (size_t i = 0; i < handles.size(); ++i) {
            api.unmap(base + i * chunk, chunk);
            api.release(handles[i]);
        }
        handles.clear();
        mapped = 0;
    }
    void destroy(const VmApi& api) {
        if (!base) return;
        unmap_all(api);
        api.address_free(base, reserved);
        base = 0;
        reserved = 0;
    }
};
constexpr size_t VM_MAX_AREAS = size_t(1) << 21;  // 2 M areas: 64 GiB of voxels, more than a B200 holds

struct DevBuf {  // grow-only device scratch buffer (contents not preserved on growth)
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {  // retry with the exact size before giving up
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

inline uint64_t area_key(uint32_t ax, uint32_t ay) { return ((uint64_t)ax << 32) | ay; }

// AreaLocation -> pool slot.  Open addressing with linear probing over {ax, ay, slot, used}
// entries, load factor <= 1/2, backward-shift deletion.  The entry array is byte-for-byte the
// device-side AreaMap table, so publishing a change is one memcpy.
struct HostAreaMap {
    std::vector<uint4> e;
    uint32_t mask = 0;
    size_t count = 0;
    // entries written since the device copy was last brought up to date (upload_map): a handful of
    // changes travel as {index, entry} patches, a rebuilt or heavily changed table as a whole
    std::vector<uint32_t> touched;
    bool all_touched = true;
    void touch(uint32_t h) {
        if (all_touched) return;
        if (touched.size() >= 4096) {
            all_touched = true;
            touched.clear();
        } else {
            touched.push_back(h);
        }
    }

    HostAreaMap() { e.assign(64, make_uint4(0, 0, 0, 0)); mask = 63; }
    int find(uint32_t ax, uint32_t ay) const {
        uint32_t h = area_hash(ax, ay) & mask;
        while (e[h].w) {
            if (e[h].x == ax && e[h].y == ay) return (int)e[h].z;
            h = (h + 1) & mask;
        }
        return -1;
    }
    void grow() {
        std::vector<uint4> old;
        old.swap(e);
        e.assign(old.size() * 2, make_uint4(0, 0, 0, 0));
        mask = (uint32_t)e.size() - 1;
        all_touched = true;
        touched.clear();
        for (const uint4& v : old)
            if (v.w) place(v.x, v.y, v.z);
    }
    void place(uint32_t ax, uint32_t ay, uint32_t slot) {
        uint32_t h = area_hash(ax, ay) & mask;
        while (e[h].w) h = (h + 1) & mask;
        e[h] = make_uint4(ax, ay, slot, 1u);
        touch(h);
    }
    void insert(uint32_t ax, uint32_t ay, int slot) {  // key must be absent
        if (2 * (count + 1) > e.size()) grow();
        place(ax, ay, (uint32_t)slot);
        ++count;
    }
    int erase(uint32_t ax, uint32_t ay) {  // returns the slot, or -1
        uint32_t h = area_hash(ax, ay) & mask;
        while (e[h].w && !(e[h].x == ax && e[h].y == ay)) h = (h + 1) & mask;
        if (!e[h].w) return -1;
        const int slot = (int)e[h].z;
        uint32_t hole = h, j = h;
        while (true) {  // backward-shift: keep every remaining key reachable from its home bucket
            j = (j + 1) & mask;
            if (!e[j].w) break;
            const uint32_t home = area_hash(e[j].x, e[j].y) & mask;
            if (((j - home) & mask) >= ((j - hole) & mask)) {
                e[hole] = e[j];
                touch(hole);
                hole = j;
            }
        }
        e[hole] = make_uint4(0, 0, 0, 0);
        touch(hole);
        --count;
        return slot;
    }
};

enum TimerId { TM_GEN_COLUMNS, TM_GEN_CAVES, TM_GEN_FINISH, TM_MESH_COUNT, TM_MESH_EMIT, TM_EDIT,
               TM_HEIGHTMAP, TM_LIGHT, TM_VISIBLE, TM_CODEC, TM_COUNT };

}  // namespace

struct vx_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // device -> host copies that need not queue behind later kernels: vx_encode_read waits for the
    // encoder (codec_ev), not for a vx_mesh enqueued after it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t codec_ev = nullptr;
    // deferred reads (vx_set_deferred_reads): vx_encode_read / vx_mesh_read_compact leave their copies on
    // copy_stream and return; files_ev / packed_ev mark the end of the last such copy, and whatever would
    // overwrite the device buffers they read waits for them on the device
    bool deferred_reads = false, files_pending = false, packed_pending = false, heights_pending = false;
    cudaEvent_t files_ev = nullptr, packed_ev = nullptr, pack_done_ev = nullptr, heights_ev = nullptr, heights_staged_ev = nullptr;
    // vx_apply_edits that does not wait: the batch is staged in pinned memory of the context's own
    PinnedBuf h_edits;
    cudaEvent_t edits_ev = nullptr;
    bool edits_pending = false;
    std::recursive_mutex mu;  // recursive: vx_lock / vx_unlock bracket several calls of one thread
    std::string last_error;

    bool has_seed = false;
    bool has_order = false;  // vx_set_gradient_order: gradient-table permutations (libnoise calibration)
    uint8_t order2[24] = {}, order3[12] = {};
    NoiseTables tables{};
    NoiseImage* d_noise_image = nullptr;  // lookup tables of the current seed (vx_tables.hpp)

    // resident-area pool: slot s owns voxels[s*32768..], heights[s*256..], planes[s*4096 u16..]
    int capacity = 0;
    uint8_t* d_voxels = nullptr;
    uint8_t* d_heights = nullptr;
    uint16_t* d_planes = nullptr;
    uint2* d_slot_xy = nullptr;
    // device mirror of Renderer.meshes (only while mesh retention is on, vx_update.cu): per slot one
    // face-mask byte per voxel (0 = no mesh entry) and a "has a RenderArea" flag; the volume of a slot
    // that is not meshed is all zero
    bool retention = false;
    uint8_t* d_face = nullptr;
    uint8_t* d_meshed = nullptr;
    // the six pool arrays above as growable virtual-memory ranges (VmArray); vmm == false (the driver
    // lacks the API): plain cudaMalloc arrays that grow by allocate-copy-free
    VmApi vm;
    bool vmm = false;
    VmArray vm_voxels, vm_heights, vm_planes, vm_xy, vm_face, vm_meshed;
    std::vector<uint64_t> slot_key;    // per slot; ~0 = free
    std::vector<int> free_slots;
    HostAreaMap index;
    uint64_t index_version = 0;  // bumped whenever the resident set changes
    std::vector<uint32_t> last_gen_xy;  // area list + slots of the previous vx_generate
    std::vector<int> last_gen_slots;
    uint64_t last_gen_version = ~0ull;
    std::vector<uint32_t> last_mesh_xy;  // area list resolved in d_jobs by the previous vx_mesh
    uint64_t last_mesh_version = ~0ull;

    // device copy of `index` (see AreaMap)
    uint32_t map_mask = 0;
    DevBuf d_map, d_map_patch;
    PinnedBuf h_map;             // staging of the table / of the patches on their way to the device
    cudaEvent_t map_ev = nullptr;
    bool map_pending = false;
    bool map_dirty = true;

    // scratch
    PinnedBuf h_stage;   // jobs / lists going to the device
    PinnedBuf h_small;   // small results coming back
    bool gen_counters_pending = false;  // d_counters holds the work counters of a timed vx_generate
    cudaEvent_t stage_ev = nullptr;  // h_stage has been consumed by the copy enqueued from it
    uint32_t* h_totals = nullptr;    // mapped pinned words the mesh scan writes for the host (vx_mesh)
    cudaEvent_t totals_ev = nullptr;
    bool stage_pending = false;
    DevBuf d_jobs, d_gen_jobs, d_list, d_missing, d_cave_items, d_counters, d_rec_count, d_face_count;
    DevBuf d_gen_counters;  // cave work list size + work accounting of vx_generate; no other call touches it
    DevBuf d_hash_keys, d_hash_vals, d_dirty_flags, d_dirty_slots, d_edits, d_tmp;
    DevBuf d_heightmap, d_light, d_slot_to_job, d_finished, d_warp_totals, d_mesh_nbr, d_heights_stage;
    DevBuf d_mesh_xy, d_frame_flags, d_frame_counts, d_frame_small, d_frame_visible, d_frame_codes;  // per-frame visibility (K7)
    DevBuf d_codec_sizes, d_codec_offsets, d_codec_bytes, d_codec_status, d_codec_kinds;  // area save files (K8 / K9)
    // incremental updates (vx_update_locations) and the compact record transport
    DevBuf d_upd_xyz, d_upd_slot, d_upd_info, d_upd_rec_count, d_upd_face_count, d_upd_rec_first, d_upd_face_first;
    DevBuf d_upd_recs, d_upd_vertices, d_upd_indices, d_packed, d_zone_xy, d_zone_slot;
    int upd_n = -1;                // locations of the last vx_update_locations
    vx_mesh_info upd_info{};
    int frame_areas = 0;           // areas the result of the last vx_visible / vx_visible_zone covers
    int codec_n = -1;              // areas of the last vx_encode_areas
    uint64_t codec_total = 0;
    bool frame_ready = false;      // d_frame_visible / d_frame_flags hold the result of vx_visible for this mesh
    uint64_t frame_n_visible = 0;
    bool heightmap_init = false;

    // last mesh
    DevBuf d_rec_first, d_face_first, d_recs, d_vertices, d_indices;
    int mesh_n = -1;
    vx_mesh_info mesh_info{};

    // timing
    bool timing = false;
    cudaEvent_t ev[TM_COUNT][2] = {};
    cudaEvent_t ev_start[TM_COUNT] = {};  // ev[id][0], or the end event of the timer it follows directly
    int ev_parent[TM_COUNT] = {};         // that timer (-1: none)
    bool ev_used[TM_COUNT] = {};
    vx_timings timings{};
};

namespace {

int fail_cuda(vx_ctx* c, cudaError_t e, const char* what) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    c->last_error = buf;
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? VX_ERR_OOM : VX_ERR_CUDA;
}
int fail(vx_ctx* c, int status, const char* what) {
    c->last_error = what;
    return status;
}

#define VX_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);   \
    } while (0)

float* timer_slot(vx_ctx* c, int id) {
    float* dst[TM_COUNT] = {&c->timings.gen_columns_ms, &c->timings.gen_caves_ms,
                            &c->timings.gen_finish_ms,  &c->timings.mesh_count_ms,
                            &c->timings.mesh_emit_ms,   &c->timings.edit_ms,
                            &c->timings.heightmap_ms,   &c->timings.light_ms,
                            &c->timings.visible_ms,     &c->timings.codec_ms};
    return dst[id];
}

// adds the elapsed time of a recorded pair to the accumulators (the pair must have completed)
void collect_timer(vx_ctx* c, int id) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_start[id], c->ev[id][1]) == cudaSuccess) *timer_slot(c, id) += ms;
    c->ev_used[id] = false;
}

// Before the events of timer `id` are recorded again: wait for its pending measurement and for the
// pending measurements that start at its end event, and add them to the accumulators.
void settle_timer(vx_ctx* c, int id) {
    for (int j = 0; j < TM_COUNT; ++j)
        if (j != id && c->ev_used[j] && c->ev_parent[j] == id) settle_timer(c, j);
    if (c->ev_used[id]) {
        cudaEventSynchronize(c->ev[id][1]);
        collect_timer(c, id);
    }
}

// Timings accumulate from one vx_get_timings to the next.  A pair recorded by a call that returned
// without waiting (vx_generate without host outputs) stays pending until the next synchronisation.
// `follows` = a timer whose scope closed immediately before this one opens (nothing else went
// into the stream in between): its end event is this timer's start, one record instead of two --
// every record between two kernels keeps the second from being launched ahead.
struct ScopedTimer {
    vx_ctx* c;
    int id;
    ScopedTimer(vx_ctx* ctx, int which, int follows = -1) : c(ctx), id(which) {
        if (c->timing) {
            settle_timer(c, id);  // whatever is still pending from a call that did not wait
            if (follows >= 0 && c->ev_used[follows]) {
                c->ev_start[id] = c->ev[follows][1];
                c->ev_parent[id] = follows;
            } else {
                cudaEventRecord(c->ev[id][0], c->stream);
                c->ev_start[id] = c->ev[id][0];
                c->ev_parent[id] = -1;
            }
            c->ev_used[id] = true;
        }
    }
    ~ScopedTimer() {
        if (c->timing) cudaEventRecord(c->ev[id][1], c->stream);
    }
};

// call after the stream has been synchronised
void collect_timers(vx_ctx* c) {
    for (int i = 0; i < TM_COUNT; ++i)
        if (c->ev_used[i]) collect_timer(c, i);
}

// h_stage may still be the source of a copy enqueued by a call that did not wait
int stage_ready(vx_ctx* ctx) {
    if (ctx->stage_pending) {
        VX_CUDA(cudaEventSynchronize(ctx->stage_ev));
        ctx->stage_pending = false;
    }
    return VX_OK;
}

int sync_stream(vx_ctx* ctx) {
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    VX_CUDA(cudaGetLastError());
    ctx->stage_pending = false;
    ctx->map_pending = false;
    collect_timers(ctx);
    return VX_OK;
}

// ---- pool ---------------------------------------------------------------------------------
// free_slots is kept sorted in DESCENDING order: pop_back hands out the smallest free slot, so the fresh
// slots of a batch ascend and a batch lands contiguously where it can.
// Slots [lo, hi) -- all larger than any slot the pool had -- become free.
void add_free_range(vx_ctx* ctx, int lo, int hi) {
    std::vector<int> fresh;
    fresh.reserve((size_t)(hi - lo) + ctx->free_slots.size());
    for (int s = hi - 1; s >= lo; --s) fresh.push_back(s);
    fresh.insert(fresh.end(), ctx->free_slots.begin(), ctx->free_slots.end());
    ctx->free_slots.swap(fresh);
}

// `freed` (any order) joins the free list: one merge instead of sorting the whole list per call
void add_free_slots(vx_ctx* ctx, std::vector<int>& freed) {
    if (freed.empty()) return;
    if (std::is_sorted(freed.begin(), freed.end())) std::reverse(freed.begin(), freed.end());  // the usual case
    else std::sort(freed.begin(), freed.end(), std::greater<int>());
    std::vector<int> merged(ctx->free_slots.size() + freed.size());
    std::merge(ctx->free_slots.begin(), ctx->free_slots.end(), freed.begin(), freed.end(), merged.begin(), std::greater<int>());
    ctx->free_slots.swap(merged);
}

int grow_pool(vx_ctx* ctx, int min_capacity) {
    if (min_capacity <= ctx->capacity) return VX_OK;
    int new_cap = std::max({min_capacity, ctx->capacity * 2, 256});
    if (ctx->vmm) {
        // map more physical memory behind the arrays: no copy, no wait, no pointer changes.  Growth is
        // geometric only up to 64 Ki areas a step, the rest of the GPU is not claimed ahead of need.
        new_cap = std::max(min_capacity, std::min(new_cap, ctx->capacity + 65536));
        if ((size_t)new_cap > VM_MAX_AREAS) return fail(ctx, VX_ERR_OOM, "more than 2 M resident areas");
        const size_t c = (size_t)new_cap;
        cudaError_t e = ctx->vm_voxels.grow(ctx->vm, ctx->device, c * VOXELS_IN_AREA);
        if (e == cudaSuccess) e = ctx->vm_heights.grow(ctx->vm, ctx->device, c * COLUMNS_IN_AREA);
        if (e == cudaSuccess) e = ctx->vm_planes.grow(ctx->vm, ctx->device, c * 4096 * sizeof(uint16_t));
        if (e == cudaSuccess) e = ctx->vm_xy.grow(ctx->vm, ctx->device, c * sizeof(uint2));
        if (e == cudaSuccess && ctx->retention) {
            const size_t old_face = (size_t)ctx->capacity * VOXELS_IN_AREA;
            e = ctx->vm_face.grow(ctx->vm, ctx->device, c * VOXELS_IN_AREA);
            if (e == cudaSuccess) e = ctx->vm_meshed.grow(ctx->vm, ctx->device, c);
            // slots that have never been meshed hold an empty mesh
            if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_face + old_face, 0, c * VOXELS_IN_AREA - old_face, ctx->stream);
            if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_meshed + ctx->capacity, 0, c - (size_t)ctx->capacity, ctx->stream);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, VX_ERR_OOM, "pool growth: out of device memory");  // what was mapped stays mapped: harmless
        }
        ctx->slot_key.resize(new_cap, ~0ull);
        add_free_range(ctx, ctx->capacity, new_cap);
        ctx->capacity = new_cap;
        return VX_OK;
    }
    uint8_t *nv = nullptr, *nh = nullptr, *nf = nullptr, *nm = nullptr;
    uint16_t* np = nullptr;
    uint2* nxy = nullptr;
    auto free_new = [&]() {
        cudaFree(nv); cudaFree(nh); cudaFree(np); cudaFree(nxy); cudaFree(nf); cudaFree(nm);
        nv = nh = nf = nm = nullptr; np = nullptr; nxy = nullptr;
    };
    auto alloc_all = [&](int cap) -> cudaError_t {
        cudaError_t e;
        if ((e = cudaMalloc(&nv, (size_t)cap * VOXELS_IN_AREA)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&nh, (size_t)cap * COLUMNS_IN_AREA)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&np, (size_t)cap * 4096 * sizeof(uint16_t))) != cudaSuccess) return e;
        if (ctx->retention) {
            if ((e = cudaMalloc(&nf, (size_t)cap * VOXELS_IN_AREA)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&nm, (size_t)cap)) != cudaSuccess) return e;
        }
        return cudaMalloc(&nxy, (size_t)cap * sizeof(uint2));
    };
    cudaError_t e = alloc_all(new_cap);
    if (e != cudaSuccess && new_cap > min_capacity) {
        cudaGetLastError();
        free_new();
        new_cap = min_capacity;
        e = alloc_all(new_cap);
    }
    if (e != cudaSuccess) {
        free_new();
        return fail_cuda(ctx, e, "pool allocation");
    }
    if (ctx->retention) {  // slots that have never been meshed hold an empty mesh
        cudaError_t ce = cudaMemsetAsync(nf, 0, (size_t)new_cap * VOXELS_IN_AREA, ctx->stream);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(nm, 0, (size_t)new_cap, ctx->stream);
        if (ce != cudaSuccess) {
            free_new();
            return fail_cuda(ctx, ce, "pool allocation (mesh retention)");
        }
    }
    if (ctx->capacity > 0) {
        const size_t c = (size_t)ctx->capacity;
        cudaError_t ce = cudaMemcpyAsync(nv, ctx->d_voxels, c * VOXELS_IN_AREA, cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(nh, ctx->d_heights, c * COLUMNS_IN_AREA, cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(np, ctx->d_planes, c * 4096 * sizeof(uint16_t), cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(nxy, ctx->d_slot_xy, c * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess && ctx->retention) ce = cudaMemcpyAsync(nf, ctx->d_face, c * VOXELS_IN_AREA, cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess && ctx->retention) ce = cudaMemcpyAsync(nm, ctx->d_meshed, c, cudaMemcpyDeviceToDevice, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) {  // the old pool stays in place
            free_new();
            return fail_cuda(ctx, ce, "pool growth copy");
        }
        cudaFree(ctx->d_voxels); cudaFree(ctx->d_heights); cudaFree(ctx->d_planes); cudaFree(ctx->d_slot_xy);
        cudaFree(ctx->d_face); cudaFree(ctx->d_meshed);
    }
    ctx->d_voxels = nv; ctx->d_heights = nh; ctx->d_planes = np; ctx->d_slot_xy = nxy;
    ctx->d_face = nf; ctx->d_meshed = nm;
    ctx->slot_key.resize(new_cap, ~0ull);
    add_free_range(ctx, ctx->capacity, new_cap);
    ctx->capacity = new_cap;
    return VX_OK;
}

// slots for the listed areas (existing slot if resident, else a fresh one). Fresh slots of one call
// are handed out in ascending order so a batch lands contiguously in the pool.
int acquire_slots(vx_ctx* ctx, const uint32_t* area_xy, int n, std::vector<int>& slots) {
    slots.resize(n);
    const int used = ctx->capacity - (int)ctx->free_slots.size();
    if ((int)ctx->free_slots.size() < n) {
        // upper bound: every listed area may need a fresh slot (cheap to over-reserve: slots only)
        int missing = 0;
        for (int i = 0; i < n; ++i)
            if (ctx->index.find(area_xy[2 * i], area_xy[2 * i + 1]) < 0) ++missing;
        if ((int)ctx->free_slots.size() < missing) {
            const int st = grow_pool(ctx, used + missing);
            if (st != VX_OK) return st;
        }
    }
    for (int i = 0; i < n; ++i) {
        const uint32_t ax = area_xy[2 * i], ay = area_xy[2 * i + 1];
        int s = ctx->index.find(ax, ay);
        if (s < 0) {
            s = ctx->free_slots.back();
            ctx->free_slots.pop_back();
            ctx->index.insert(ax, ay, s);
            ctx->slot_key[s] = area_key(ax, ay);
            ctx->map_dirty = true;
            ++ctx->index_version;
        }
        slots[i] = s;
    }
    return VX_OK;
}

// Brings the device copy of the area map up to date, in stream order and without waiting: the table
// (or, after a few changes, just the changed entries) goes through a pinned staging buffer of its
// own; map_ev tells when the previous upload has finished reading that buffer.
int upload_map(vx_ctx* ctx) {
    if (!ctx->map_dirty) return VX_OK;
    HostAreaMap& m = ctx->index;
    const size_t bytes = m.e.size() * sizeof(uint4);
    const bool whole = m.all_touched || ctx->d_map.cap < bytes || ctx->map_mask != m.mask;
    if (ctx->map_pending) {
        VX_CUDA(cudaEventSynchronize(ctx->map_ev));
        ctx->map_pending = false;
    }
    if (whole) {
        if (ctx->d_map.cap < bytes) {  // the old table may still be read by kernels in flight
            VX_CUDA(cudaStreamSynchronize(ctx->stream));
            VX_CUDA(ctx->d_map.ensure(bytes));
        }
        VX_CUDA(ctx->h_map.ensure(bytes));
        std::memcpy(ctx->h_map.p, m.e.data(), bytes);
        VX_CUDA(cudaMemcpyAsync(ctx->d_map.p, ctx->h_map.p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else if (!m.touched.empty()) {
        const size_t k = m.touched.size();
        VX_CUDA(ctx->h_map.ensure(k * 32));
        VX_CUDA(ctx->d_map_patch.ensure(k * 32));
        uint4* stage = ctx->h_map.as<uint4>();
        for (size_t i = 0; i < k; ++i) {  // {index, -, -, -}, {entry}: later patches of one entry repeat its final value
            stage[2 * i] = make_uint4(m.touched[i], 0u, 0u, 0u);
            stage[2 * i + 1] = m.e[m.touched[i]];
        }
        VX_CUDA(cudaMemcpyAsync(ctx->d_map_patch.p, stage, k * 32, cudaMemcpyHostToDevice, ctx->stream));
        launch_map_patch(ctx->d_map_patch.as<uint4>(), (int)k, ctx->d_map.as<uint4>(), ctx->stream);
        VX_CUDA(cudaGetLastError());
    }
    VX_CUDA(cudaEventRecord(ctx->map_ev, ctx->stream));
    ctx->map_pending = true;
    m.touched.clear();
    m.all_touched = false;
    ctx->map_mask = m.mask;
    ctx->map_dirty = false;
    return VX_OK;
}

AreaMap device_map(const vx_ctx* ctx) {
    AreaMap m;
    m.entries = ctx->d_map.as<uint4>();
    m.mask = ctx->map_mask;
    return m;
}

int host_slot(const vx_ctx* ctx, uint32_t ax, uint32_t ay) { return ctx->index.find(ax, ay); }

// jobs {ax, ay, slot, 0} -> device (through pinned staging; the stream is idle between calls)
int upload_jobs(vx_ctx* ctx, const uint32_t* area_xy, const std::vector<int>& slots, int n) {
    ctx->last_mesh_version = ~0ull;  // d_jobs no longer holds the list of the last vx_mesh
    if (int st = stage_ready(ctx)) return st;
    VX_CUDA(ctx->h_stage.ensure((size_t)n * sizeof(uint4)));
    VX_CUDA(ctx->d_jobs.ensure((size_t)n * sizeof(uint4)));
    uint4* j = ctx->h_stage.as<uint4>();
    for (int i = 0; i < n; ++i) j[i] = make_uint4(area_xy[2 * i], area_xy[2 * i + 1], (uint32_t)slots[i], 0u);
    VX_CUDA(cudaMemcpyAsync(ctx->d_jobs.p, j, (size_t)n * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
    return VX_OK;
}

// copy per-slot records (stride bytes each) of the listed slots to/from a packed host array, one
// cudaMemcpyAsync per run of consecutive slots
int copy_slots(vx_ctx* ctx, const std::vector<int>& slots, int n, uint8_t* dev_base, size_t stride,
               uint8_t* host, bool to_host, bool host_is_device = false) {
    int i = 0;
    while (i < n) {
        int j = i + 1;
        while (j < n && slots[j] == slots[j - 1] + 1) ++j;
        uint8_t* d = dev_base + (size_t)slots[i] * stride;
        uint8_t* h = host + (size_t)i * stride;
        const size_t bytes = (size_t)(j - i) * stride;
        if (host_is_device) VX_CUDA(cudaMemcpyAsync(to_host ? h : d, to_host ? d : h, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        else if (to_host) VX_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        else VX_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
        i = j;
    }
    return VX_OK;
}

// K1 -> K1b -> K2 for areas already holding slots; asynchronous on the stream
int generate_async(vx_ctx* ctx, const uint32_t* area_xy, const std::vector<int>& slots, int n,
                   bool jobs_on_device = false) {
    if (!jobs_on_device) {  // jobs {ax, ay, slot, 0} -> d_gen_jobs through pinned staging
        int st = stage_ready(ctx);
        if (st != VX_OK) return st;
        VX_CUDA(ctx->h_stage.ensure((size_t)n * sizeof(uint4)));
        VX_CUDA(ctx->d_gen_jobs.ensure((size_t)n * sizeof(uint4)));
        uint4* j = ctx->h_stage.as<uint4>();
        for (int i = 0; i < n; ++i) j[i] = make_uint4(area_xy[2 * i], area_xy[2 * i + 1], (uint32_t)slots[i], 0u);
        VX_CUDA(cudaMemcpyAsync(ctx->d_gen_jobs.p, j, (size_t)n * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
        VX_CUDA(cudaEventRecord(ctx->stage_ev, ctx->stream));
        ctx->stage_pending = true;
        ctx->last_gen_version = ~0ull;  // whatever list d_gen_jobs held before is gone
    }
    VX_CUDA(ctx->d_cave_items.ensure((size_t)n * COLUMNS_IN_AREA * sizeof(CaveItem)));
    VX_CUDA(ctx->d_finished.ensure((size_t)n * sizeof(uint32_t)));  // K2 work list
    VX_CUDA(ctx->d_gen_counters.ensure(64));
    VX_CUDA(cudaMemsetAsync(ctx->d_gen_counters.p, 0, 64, ctx->stream));
    const uint4* jobs = ctx->d_gen_jobs.as<uint4>();
    uint32_t* cave_count = ctx->d_gen_counters.as<uint32_t>();
    unsigned long long* stats = ctx->d_gen_counters.as<unsigned long long>() + 2;  // bytes 16..55
    {
        ScopedTimer t(ctx, TM_GEN_COLUMNS);
        launch_gen_columns(ctx->tables, jobs, n, ctx->d_voxels, ctx->d_heights, ctx->d_slot_xy,
                           ctx->d_cave_items.as<CaveItem>(), cave_count, stats, ctx->d_planes,
                           ctx->d_finished.as<uint32_t>(), ctx->stream);
    }
    {
        ScopedTimer t(ctx, TM_GEN_CAVES, TM_GEN_COLUMNS);
        launch_gen_caves(ctx->tables, ctx->d_cave_items.as<CaveItem>(), cave_count, ctx->d_voxels, ctx->d_heights,
                         ctx->d_slot_xy, stats, ctx->sm_count, ctx->stream);
    }
    {
        ScopedTimer t(ctx, TM_GEN_FINISH, TM_GEN_CAVES);
        launch_gen_finish(jobs, ctx->d_finished.as<uint32_t>(), cave_count + 1, n, ctx->d_voxels, ctx->d_heights,
                          ctx->d_planes, ctx->tables.seed, ctx->sm_count, ctx->stream);
    }
    ctx->timings.gen_launches += 3;
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

// mesh retention: the listed slots lose their RenderArea (face volume zero, flag cleared); one
// memset per run of consecutive slots, in stream order
int clear_mesh_slots(vx_ctx* ctx, std::vector<int>& slots) {
    std::sort(slots.begin(), slots.end());
    size_t i = 0;
    while (i < slots.size()) {
        size_t j = i + 1;
        while (j < slots.size() && slots[j] == slots[j - 1] + 1) ++j;
        VX_CUDA(cudaMemsetAsync(ctx->d_face + (size_t)slots[i] * VOXELS_IN_AREA, 0, (j - i) * VOXELS_IN_AREA, ctx->stream));
        VX_CUDA(cudaMemsetAsync(ctx->d_meshed + slots[i], 0, j - i, ctx->stream));
        i = j;
    }
    return VX_OK;
}

}  // namespace

// =========================================================================================== ABI
namespace {

// Every area a simulator loop can read for these locations (the location's own column and its four
// sides) is resident -- checked up front, so that a call either runs whole or changes nothing.
int sim_check_resident(vx_ctx* ctx, const uint32_t* xyz, int n, const char* who) {
    for (int i = 0; i < n; ++i) {
        const uint32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (z >= (uint32_t)AREA_HEIGHT) return fail(ctx, VX_ERR_INVALID, (std::string(who) + ": z >= 128").c_str());
        const uint32_t xs[3] = {x, x + 1, x - 1}, ys[3] = {y, y + 1, y - 1};
        for (int k = 0; k < 3; ++k)
            if (host_slot(ctx, xs[k] / AREA_SIZE, y / AREA_SIZE) < 0 || host_slot(ctx, x / AREA_SIZE, ys[k] / AREA_SIZE) < 0)
                return fail(ctx, VX_ERR_NOT_LOADED, (std::string(who) + ": an area next to a location is not resident").c_str());
    }
    return VX_OK;
}

// Runs one of the three ordered-list kernels: `in_bytes` of input go to the device, two lists
// (entries of a_words / 3 words) and optionally n keep bytes come back.
struct SimCall {
    const void* in0; size_t in0_bytes;
    const void* in1; size_t in1_bytes;
    uint32_t* a; int cap_a; int a_words; int* n_a;
    uint32_t* b; int cap_b; int* n_b;
    uint8_t* keep; int n_keep;
};

template <typename Launch>
int run_sim(vx_ctx* ctx, const SimCall& c, Launch launch) {
    int st;
    if ((st = upload_map(ctx)) != VX_OK) return st;
    if ((st = sync_stream(ctx)) != VX_OK) return st;  // d_edits may be in use by a call that did not wait
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t off_in1 = up16(c.in0_bytes), off_a = off_in1 + up16(c.in1_bytes);
    const size_t off_b = off_a + up16((size_t)c.cap_a * c.a_words * 4), off_keep = off_b + up16((size_t)c.cap_b * 12);
    VX_CUDA(ctx->d_edits.ensure(off_keep + up16((size_t)c.n_keep) + 16));
    VX_CUDA(ctx->d_counters.ensure(64));
    uint8_t* base = ctx->d_edits.as<uint8_t>();
    VX_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(base, c.in0, c.in0_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (c.in1_bytes) VX_CUDA(cudaMemcpyAsync(base + off_in1, c.in1, c.in1_bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* counters = ctx->d_counters.as<uint32_t>();  // [0] entries of a, [1] entries of b, [2] status
    {
        ScopedTimer t(ctx, TM_EDIT);
        launch(base, base + off_in1, reinterpret_cast<uint32_t*>(base + off_a), reinterpret_cast<uint32_t*>(base + off_b),
               base + off_keep, counters, reinterpret_cast<int32_t*>(counters + 2));
    }
    ctx->timings.other_launches += 1;
    VX_CUDA(cudaGetLastError());
    VX_CUDA(ctx->h_small.ensure(16));
    uint32_t* hs = ctx->h_small.as<uint32_t>();
    VX_CUDA(cudaMemcpyAsync(hs, counters, 12, cudaMemcpyDeviceToHost, ctx->stream));
    if ((st = sync_stream(ctx)) != VX_OK) return st;
    collect_timers(ctx);
    const uint32_t na = hs[0], nb = hs[1];
    const int32_t status = (int32_t)hs[2];
    // whatever was applied before a failure stays applied and is reported, like a panic half way through
    // the reference's loop would leave it
    if (na) VX_CUDA(cudaMemcpyAsync(c.a, base + off_a, (size_t)na * c.a_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (nb) VX_CUDA(cudaMemcpyAsync(c.b, base + off_b, (size_t)nb * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (c.n_keep) VX_CUDA(cudaMemcpyAsync(c.keep, base + off_keep, (size_t)c.n_keep, cudaMemcpyDeviceToHost, ctx->stream));
    if ((st = sync_stream(ctx)) != VX_OK) return st;
    *c.n_a = (int)na;
    *c.n_b = (int)nb;
    if (status != 0)
        return fail(ctx, status, status == VX_ERR_NOT_LOADED ? "simulator step reached an area that is not resident"
                                                             : "simulator step: an output list is too small or a location is out of range");
    return VX_OK;
}

}  // namespace

extern "C" {

int vx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return VX_ERR_NO_DEVICE;
    }
    return n;
}

const char* vx_strerror(int status) {
    switch (status) {
    case VX_OK: return "ok";
    case VX_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case VX_ERR_CUDA: return "CUDA error";
    case VX_ERR_INVALID: return "invalid argument";
    case VX_ERR_NO_SEED: return "vx_set_seed has not been called";
    case VX_ERR_NOT_LOADED: return "area not resident";
    case VX_ERR_OOM: return "out of memory";
    case VX_ERR_NO_MESH: return "no mesh available";
    default: return "unknown status";
    }
}

int vx_create(int device, vx_ctx** out) {
    if (!out) return VX_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return VX_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return VX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return VX_ERR_NO_DEVICE;
    }
    vx_ctx* ctx = new vx_ctx();
    ctx->device = device;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return VX_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->codec_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->files_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->packed_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->pack_done_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->heights_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->heights_staged_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->edits_ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return VX_ERR_CUDA;
    }
    for (int i = 0; i < TM_COUNT; ++i)
        for (int k = 0; k < 2; ++k) cudaEventCreate(&ctx->ev[i][k]);
    cudaEventCreateWithFlags(&ctx->stage_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->totals_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->map_ev, cudaEventDisableTiming);
    if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_totals), 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        vx_destroy(ctx);
        return VX_ERR_CUDA;
    }
    // the pool arrays as virtual-memory ranges (see VmArray); without the API: cudaMalloc arrays
    if (ctx->vm.load()) {
        const size_t big = size_t(64) << 20, small = size_t(2) << 20;
        ctx->vmm = ctx->vm_voxels.reserve(ctx->vm, device, VM_MAX_AREAS * VOXELS_IN_AREA, big) &&
                   ctx->vm_heights.reserve(ctx->vm, device, VM_MAX_AREAS * COLUMNS_IN_AREA, small) &&
                   ctx->vm_planes.reserve(ctx->vm, device, VM_MAX_AREAS * 4096 * sizeof(uint16_t), big / 4) &&
                   ctx->vm_xy.reserve(ctx->vm, device, VM_MAX_AREAS * sizeof(uint2), small) &&
                   ctx->vm_face.reserve(ctx->vm, device, VM_MAX_AREAS * VOXELS_IN_AREA, big) &&
                   ctx->vm_meshed.reserve(ctx->vm, device, VM_MAX_AREAS, small);
        if (ctx->vmm) {
            ctx->d_voxels = ctx->vm_voxels.as<uint8_t>();
            ctx->d_heights = ctx->vm_heights.as<uint8_t>();
            ctx->d_planes = ctx->vm_planes.as<uint16_t>();
            ctx->d_slot_xy = ctx->vm_xy.as<uint2>();
            ctx->d_face = ctx->vm_face.as<uint8_t>();
            ctx->d_meshed = ctx->vm_meshed.as<uint8_t>();
        } else {
            for (VmArray* a : {&ctx->vm_voxels, &ctx->vm_heights, &ctx->vm_planes, &ctx->vm_xy, &ctx->vm_face, &ctx->vm_meshed})
                a->destroy(ctx->vm);
        }
    }
    *out = ctx;
    return VX_OK;
}

void vx_destroy(vx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);  // deferred reads still in flight
    if (ctx->vmm) {
        for (VmArray* a : {&ctx->vm_voxels, &ctx->vm_heights, &ctx->vm_planes, &ctx->vm_xy, &ctx->vm_face, &ctx->vm_meshed})
            a->destroy(ctx->vm);
    } else {
        cudaFree(ctx->d_voxels); cudaFree(ctx->d_heights); cudaFree(ctx->d_planes); cudaFree(ctx->d_slot_xy);
        cudaFree(ctx->d_face); cudaFree(ctx->d_meshed);
    }
    cudaFree(ctx->d_noise_image);
    DevBuf* bufs[] = {&ctx->d_map, &ctx->d_map_patch, &ctx->d_jobs, &ctx->d_gen_jobs, &ctx->d_list, &ctx->d_missing, &ctx->d_cave_items, &ctx->d_counters, &ctx->d_gen_counters,
                      &ctx->d_rec_count, &ctx->d_face_count, &ctx->d_hash_keys, &ctx->d_hash_vals,
                      &ctx->d_dirty_flags, &ctx->d_dirty_slots, &ctx->d_edits, &ctx->d_tmp,
                      &ctx->d_heightmap, &ctx->d_light, &ctx->d_slot_to_job, &ctx->d_finished, &ctx->d_warp_totals, &ctx->d_mesh_nbr, &ctx->d_heights_stage, &ctx->d_mesh_xy, &ctx->d_frame_flags,
                      &ctx->d_frame_counts, &ctx->d_frame_small, &ctx->d_frame_visible, &ctx->d_codec_sizes,
                      &ctx->d_frame_codes, &ctx->d_codec_offsets, &ctx->d_codec_bytes, &ctx->d_codec_status, &ctx->d_codec_kinds, &ctx->d_rec_first, &ctx->d_face_first, &ctx->d_recs,
                      &ctx->d_vertices, &ctx->d_indices, &ctx->d_upd_xyz, &ctx->d_upd_slot, &ctx->d_upd_info,
                      &ctx->d_upd_rec_count, &ctx->d_upd_face_count, &ctx->d_upd_rec_first, &ctx->d_upd_face_first,
                      &ctx->d_upd_recs, &ctx->d_upd_vertices, &ctx->d_upd_indices, &ctx->d_packed, &ctx->d_zone_xy,
                      &ctx->d_zone_slot};
    for (DevBuf* b : bufs) b->release();
    ctx->h_stage.release();
    ctx->h_small.release();
    cudaEventDestroy(ctx->stage_ev);
    cudaEventDestroy(ctx->totals_ev);
    cudaEventDestroy(ctx->map_ev);
    ctx->h_map.release();
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    for (int i = 0; i < TM_COUNT; ++i)
        for (int k = 0; k < 2; ++k) cudaEventDestroy(ctx->ev[i][k]);
    cudaEventDestroy(ctx->codec_ev);
    cudaEventDestroy(ctx->files_ev);
    cudaEventDestroy(ctx->packed_ev);
    cudaEventDestroy(ctx->pack_done_ev);
    cudaEventDestroy(ctx->heights_ev);
    cudaEventDestroy(ctx->heights_staged_ev);
    cudaEventDestroy(ctx->edits_ev);
    ctx->h_edits.release();
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* vx_last_error(vx_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

size_t vx_copy_last_error(vx_ctx* ctx, char* buf, size_t cap) {
    if (!ctx) return 0;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (buf && cap) {
        const size_t k = ctx->last_error.size() < cap - 1 ? ctx->last_error.size() : cap - 1;
        memcpy(buf, ctx->last_error.data(), k);
        buf[k] = 0;
    }
    return ctx->last_error.size();
}

int vx_lock(vx_ctx* ctx) {
    if (!ctx) return VX_ERR_INVALID;
    ctx->mu.lock();
    return VX_OK;
}

int vx_unlock(vx_ctx* ctx) {
    if (!ctx) return VX_ERR_INVALID;
    ctx->mu.unlock();
    return VX_OK;
}

int vx_set_stream(vx_ctx* ctx, void* cuda_stream) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    sync_stream(ctx);  // nothing stays in flight (or pending in the timers) on the stream being left
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return VX_OK;
}

int vx_sync(vx_ctx* ctx) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    return sync_stream(ctx);
}

int vx_set_timing(vx_ctx* ctx, int enabled) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    ctx->timing = enabled != 0;
    return VX_OK;
}

int vx_get_timings(vx_ctx* ctx, vx_timings* out) {
    if (!ctx || !out) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    bool pending = ctx->gen_counters_pending;
    for (int i = 0; i < TM_COUNT; ++i) pending = pending || ctx->ev_used[i];
    if (pending) {  // a call that did not wait (vx_generate without host outputs) is still in flight
        VX_CUDA(cudaSetDevice(ctx->device));
        int st = sync_stream(ctx);
        if (st != VX_OK) return st;
    }
    if (ctx->gen_counters_pending) {  // of the most recent vx_generate
        unsigned long long n[6];
        VX_CUDA(cudaMemcpy(n, ctx->d_gen_counters.p, sizeof n, cudaMemcpyDeviceToHost));
        ctx->timings.cave_columns = (uint32_t)n[0];
        ctx->timings.simplex2_evals = n[2];
        ctx->timings.cave_voxels = n[4];
        ctx->timings.simplex3_evals = n[3] + n[5];  // surface noise + cave octaves actually evaluated
        ctx->timings.cave_evals = n[5];
        ctx->gen_counters_pending = false;
    }
    *out = ctx->timings;
    ctx->timings = vx_timings{};
    return VX_OK;
}

void* vx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void vx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

uint64_t vx_hash_world_name(const char* world_name) {
    return world_name ? vx::hash_world_name(world_name) : 0;
}

int vx_set_seed(vx_ctx* ctx, uint64_t seed) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!ctx->d_noise_image) VX_CUDA(cudaMalloc(&ctx->d_noise_image, sizeof(NoiseImage)));
    build_noise_tables(seed, &ctx->tables);
    NoiseImage image;
    build_noise_image(ctx->tables, &image, ctx->has_order ? ctx->order2 : nullptr, ctx->has_order ? ctx->order3 : nullptr);
    VX_CUDA(cudaMemcpy(ctx->d_noise_image, &image, sizeof image, cudaMemcpyHostToDevice));
    ctx->tables.image = ctx->d_noise_image;
    ctx->has_seed = true;
    return VX_OK;
}

int vx_noise_variant(void) { return VX_NOISE_VARIANT; }

int vx_set_gradient_order(vx_ctx* ctx, const uint8_t* order2, const uint8_t* order3) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    bool seen2[24] = {}, seen3[12] = {};
    for (int i = 0; i < 24; ++i) {
        const int v = order2 ? order2[i] : i;
        if (v >= 24 || seen2[v]) return fail(ctx, VX_ERR_INVALID, "vx_set_gradient_order: order2 is not a permutation of 0..23");
        seen2[v] = true;
        ctx->order2[i] = (uint8_t)v;
    }
    for (int i = 0; i < 12; ++i) {
        const int v = order3 ? order3[i] : i;
        if (v >= 12 || seen3[v]) return fail(ctx, VX_ERR_INVALID, "vx_set_gradient_order: order3 is not a permutation of 0..11");
        seen3[v] = true;
        ctx->order3[i] = (uint8_t)v;
    }
    ctx->has_order = order2 != nullptr || order3 != nullptr;
    if (!ctx->has_seed) return VX_OK;
    // the tables of the current seed again, in the new order; resident areas keep their voxels
    ctx->last_gen_version = ~0ull;
    return vx_set_seed(ctx, ctx->tables.seed);
}

int vx_generate(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* voxels_out,
                uint8_t* max_height_out) {
    if (!ctx || n < 0 || (n > 0 && !area_xy)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!ctx->has_seed) return fail(ctx, VX_ERR_NO_SEED, "vx_generate before vx_set_seed");
    VX_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return VX_OK;
    // Regenerating the list of the previous call with an unchanged resident set (world
    // pre-generation sweeps, benchmarks) reuses its slot assignment: no hashing on the host.
    std::vector<int>& slots = ctx->last_gen_slots;
    const bool same_list = ctx->last_gen_version == ctx->index_version && (int)ctx->last_gen_xy.size() == 2 * n &&
                           std::memcmp(ctx->last_gen_xy.data(), area_xy, (size_t)n * 8) == 0;
    int st;
    if (!same_list) {
        ctx->last_gen_version = ~0ull;  // the cache describes nothing until this call has succeeded
        if ((st = acquire_slots(ctx, area_xy, n, slots)) != VX_OK) return st;
        ctx->last_gen_xy.assign(area_xy, area_xy + 2 * (size_t)n);
    }
    if ((st = generate_async(ctx, area_xy, slots, n, same_list)) != VX_OK) {
        ctx->last_gen_version = ~0ull;
        return st;
    }
    ctx->last_gen_version = ctx->index_version;  // d_gen_jobs now holds exactly this list
    if (ctx->deferred_reads && max_height_out && !voxels_out) {
        // Deferred read of the heights: gathered into a staging buffer on the main stream (device to
        // device, microseconds), copied to the host from there on the copy stream -- later calls may
        // change the pool's heights at once, only the staging buffer has to outlive the copy.
        const size_t bytes = (size_t)n * COLUMNS_IN_AREA;
        if (ctx->heights_pending) VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->heights_ev, 0));
        VX_CUDA(ctx->d_heights_stage.ensure(bytes));
        if ((st = copy_slots(ctx, slots, n, ctx->d_heights, COLUMNS_IN_AREA, ctx->d_heights_stage.as<uint8_t>(), true, true)) != VX_OK) return st;
        VX_CUDA(cudaEventRecord(ctx->heights_staged_ev, ctx->stream));
        VX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->heights_staged_ev, 0));
        VX_CUDA(cudaMemcpyAsync(max_height_out, ctx->d_heights_stage.p, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        VX_CUDA(cudaEventRecord(ctx->heights_ev, ctx->copy_stream));
        ctx->heights_pending = true;
        if (ctx->timing) ctx->gen_counters_pending = true;
        return VX_OK;
    }
    if (voxels_out && (st = copy_slots(ctx, slots, n, ctx->d_voxels, VOXELS_IN_AREA, voxels_out, true)) != VX_OK) return st;
    if (max_height_out && (st = copy_slots(ctx, slots, n, ctx->d_heights, COLUMNS_IN_AREA, max_height_out, true)) != VX_OK) return st;
    // work accounting of this call (the roofline numerators): fetched when the timings are read
    if (ctx->timing) ctx->gen_counters_pending = true;
    // Without host outputs there is nothing to wait for: the areas are resident in stream order,
    // and whatever reads or changes them next is enqueued behind the kernels.  Errors of the
    // kernels then surface in that later call (or in vx_sync).
    if (!voxels_out && !max_height_out) return VX_OK;
    return sync_stream(ctx);
}

int vx_upload(vx_ctx* ctx, const uint32_t* area_xy, int n, const uint8_t* voxels,
              uint8_t* max_height_out) {
    if (!ctx || n < 0 || (n > 0 && (!area_xy || !voxels))) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return VX_OK;
    std::vector<int> slots;
    int st = acquire_slots(ctx, area_xy, n, slots);
    if (st != VX_OK) return st;
    if ((st = copy_slots(ctx, slots, n, ctx->d_voxels, VOXELS_IN_AREA, const_cast<uint8_t*>(voxels), false)) != VX_OK) return st;
    if ((st = upload_jobs(ctx, area_xy, slots, n)) != VX_OK) return st;
    launch_column_heights(ctx->d_jobs.as<uint4>(), n, ctx->d_voxels, ctx->d_heights, ctx->d_slot_xy, ctx->d_planes,
                          ctx->stream);
    ctx->timings.other_launches += 1;
    VX_CUDA(cudaGetLastError());
    if (max_height_out && (st = copy_slots(ctx, slots, n, ctx->d_heights, COLUMNS_IN_AREA, max_height_out, true)) != VX_OK) return st;
    return sync_stream(ctx);
}

int vx_read_areas(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* voxels_out,
                  uint8_t* max_height_out) {
    if (!ctx || n < 0 || (n > 0 && !area_xy)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    std::vector<int> slots(n);
    for (int i = 0; i < n; ++i) {
        slots[i] = host_slot(ctx, area_xy[2 * i], area_xy[2 * i + 1]);
        if (slots[i] < 0) return fail(ctx, VX_ERR_NOT_LOADED, "vx_read_areas: area not resident");
    }
    int st;
    if (voxels_out && (st = copy_slots(ctx, slots, n, ctx->d_voxels, VOXELS_IN_AREA, voxels_out, true)) != VX_OK) return st;
    if (max_height_out && (st = copy_slots(ctx, slots, n, ctx->d_heights, COLUMNS_IN_AREA, max_height_out, true)) != VX_OK) return st;
    return sync_stream(ctx);
}

int vx_unload(vx_ctx* ctx, const uint32_t* area_xy, int n) {
    if (!ctx || n < 0 || (n > 0 && !area_xy)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    bool any = false;
    std::vector<int> freed;
    for (int i = 0; i < n; ++i) {
        const int s = ctx->index.erase(area_xy[2 * i], area_xy[2 * i + 1]);
        if (s < 0) continue;
        ctx->slot_key[s] = ~0ull;
        freed.push_back(s);
        any = true;
    }
    if (any) {
        if (ctx->retention) {  // a slot that is handed out again starts without a mesh
            VX_CUDA(cudaSetDevice(ctx->device));
            if (int st = clear_mesh_slots(ctx, freed)) return st;
        }
        ctx->map_dirty = true;
        ++ctx->index_version;
        add_free_slots(ctx, freed);  // hand-out order stays ascending, so later batches stay contiguous where possible
    }
    return VX_OK;
}

int vx_pool_info(vx_ctx* ctx, uint64_t* capacity_areas, uint64_t* committed_bytes, int* virtual_memory) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (capacity_areas) *capacity_areas = (uint64_t)ctx->capacity;
    if (virtual_memory) *virtual_memory = ctx->vmm ? 1 : 0;
    if (committed_bytes) {
        if (ctx->vmm) {
            *committed_bytes = ctx->vm_voxels.mapped + ctx->vm_heights.mapped + ctx->vm_planes.mapped + ctx->vm_xy.mapped +
                               ctx->vm_face.mapped + ctx->vm_meshed.mapped;
        } else {
            const uint64_t per_area = VOXELS_IN_AREA + COLUMNS_IN_AREA + 8192 + 8 + (ctx->retention ? VOXELS_IN_AREA + 1 : 0);
            *committed_bytes = (uint64_t)ctx->capacity * per_area;
        }
    }
    return VX_OK;
}

int vx_resident_count(vx_ctx* ctx) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return (int)ctx->index.count;
}

int vx_column_heights(vx_ctx* ctx, const uint8_t* voxels, int n, uint8_t* max_height_out) {
    if (!ctx || n < 0 || (n > 0 && (!voxels || !max_height_out))) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return VX_OK;
    // temporary areas outside the resident pool: [voxels n*32768][heights n*256]
    const size_t vbytes = (size_t)n * VOXELS_IN_AREA, hbytes = (size_t)n * COLUMNS_IN_AREA;
    VX_CUDA(ctx->d_tmp.ensure(vbytes + hbytes));
    uint8_t* dv = ctx->d_tmp.as<uint8_t>();
    uint8_t* dh = dv + vbytes;
    VX_CUDA(cudaMemcpyAsync(dv, voxels, vbytes, cudaMemcpyHostToDevice, ctx->stream));
    if (int st = stage_ready(ctx)) return st;
    ctx->last_mesh_version = ~0ull;  // d_jobs is overwritten: it no longer resolves the list of the last vx_mesh
    VX_CUDA(ctx->h_stage.ensure((size_t)n * sizeof(uint4)));
    VX_CUDA(ctx->d_jobs.ensure((size_t)n * sizeof(uint4)));
    uint4* j = ctx->h_stage.as<uint4>();
    for (int i = 0; i < n; ++i) j[i] = make_uint4(0, 0, (uint32_t)i, 0);
    VX_CUDA(cudaMemcpyAsync(ctx->d_jobs.p, j, (size_t)n * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
    launch_column_heights(ctx->d_jobs.as<uint4>(), n, dv, dh, nullptr, nullptr, ctx->stream);
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(max_height_out, dh, hbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_stream(ctx);
}

int vx_mesh(vx_ctx* ctx, const uint32_t* area_xy, int n, vx_mesh_info* info) {
    if (!ctx || n < 0 || (n > 0 && !area_xy)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    ctx->mesh_n = -1;
    ctx->frame_ready = false;
    ctx->mesh_info = vx_mesh_info{};
    int st;
    if (n > 0) {
        if (ctx->capacity == 0) return fail(ctx, VX_ERR_NOT_LOADED, "vx_mesh: area not resident");
        // rec_first of the previous mesh may still be crossing to the host (deferred read)
        if (ctx->packed_pending) VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->packed_ev, 0));
        // the area list goes to the device as is; slots, neighbour residency and the set of bit
        // planes to refresh are resolved there (mesh_prepare_kernel), not in a host loop
        const size_t xy_bytes = (size_t)n * sizeof(uint2);
        if ((st = stage_ready(ctx)) != VX_OK) return st;
        VX_CUDA(ctx->h_stage.ensure(xy_bytes));
        VX_CUDA(ctx->h_small.ensure(64 + (size_t)4 * n * sizeof(uint2)));
        VX_CUDA(ctx->d_list.ensure(xy_bytes));
        VX_CUDA(ctx->d_jobs.ensure((size_t)n * sizeof(uint4)));
        VX_CUDA(ctx->d_mesh_nbr.ensure((size_t)n * sizeof(uint4)));
        VX_CUDA(ctx->d_missing.ensure((size_t)4 * n * sizeof(uint2)));
        VX_CUDA(ctx->d_counters.ensure(64));
        VX_CUDA(ctx->d_rec_count.ensure((size_t)n * 4));
        VX_CUDA(ctx->d_face_count.ensure((size_t)n * 4));
        VX_CUDA(ctx->d_warp_totals.ensure((size_t)n * 64 * 4));
        VX_CUDA(ctx->d_rec_first.ensure((size_t)(n + 1) * 4));
        VX_CUDA(ctx->d_face_first.ensure((size_t)(n + 1) * 4));
        // same list as the previous vx_mesh and an unchanged resident set: the resolved jobs are
        // still on the device (re-meshing after edits, benchmarks)
        bool jobs_ready = ctx->last_mesh_version == ctx->index_version && ctx->last_mesh_xy.size() == 2 * (size_t)n &&
                          std::memcmp(ctx->last_mesh_xy.data(), area_xy, xy_bytes) == 0;
        if (!jobs_ready) {
            std::memcpy(ctx->h_stage.p, area_xy, xy_bytes);
            VX_CUDA(cudaMemcpyAsync(ctx->d_list.p, ctx->h_stage.p, xy_bytes, cudaMemcpyHostToDevice, ctx->stream));
            VX_CUDA(ctx->d_mesh_xy.ensure(xy_bytes));  // the list of this mesh, kept for vx_visible
            VX_CUDA(cudaMemcpyAsync(ctx->d_mesh_xy.p, ctx->d_list.p, xy_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->last_mesh_xy.assign(area_xy, area_xy + 2 * (size_t)n);
            ctx->last_mesh_version = ~0ull;
        }
        uint32_t* counters = ctx->d_counters.as<uint32_t>();
        const volatile uint32_t* back = ctx->h_totals;  // [0] missing self, [1] missing nbrs, [2] recs, [3] faces
        uint64_t n_recs = 0, n_faces = 0;
        auto arena_caps = [&](uint32_t& cap_faces, uint32_t& cap_recs) {
            cap_faces = (uint32_t)std::min<uint64_t>(std::min(ctx->d_vertices.cap / 160, ctx->d_indices.cap / 12), 0xFFFFFFFFull);
            cap_recs = (uint32_t)std::min<uint64_t>(ctx->d_recs.cap / 16, 0xFFFFFFFFull);
        };
        auto launch_emit = [&](uint32_t cap_faces, uint32_t cap_recs) {
            MeshOut out;
            out.rec_first = ctx->d_rec_first.as<uint32_t>();
            out.face_first = ctx->d_face_first.as<uint32_t>();
            out.recs = ctx->d_recs.as<uint4>();
            out.vertices = ctx->d_vertices.as<uint4>();
            out.indices = ctx->d_indices.as<uint32_t>();
            ScopedTimer t(ctx, TM_MESH_EMIT, TM_MESH_COUNT);
            launch_mesh_emit(ctx->d_jobs.as<uint4>(), ctx->d_mesh_nbr.as<uint4>(), n, ctx->d_voxels, ctx->d_planes,
                             ctx->d_warp_totals.as<uint32_t>(), out, cap_faces, cap_recs, ctx->stream);
            ctx->timings.mesh_launches += 1;
        };
        for (int attempt = 0;; ++attempt) {
            if ((st = upload_map(ctx)) != VX_OK) return st;
            const AreaMap map = device_map(ctx);
            VX_CUDA(cudaMemsetAsync(counters, 0, 8, ctx->stream));
            {
                ScopedTimer t(ctx, TM_MESH_COUNT);
                if (!jobs_ready) {
                    launch_mesh_prepare(ctx->d_list.as<uint2>(), n, map, ctx->d_jobs.as<uint4>(),
                                        ctx->d_mesh_nbr.as<uint4>(), counters,
                                        ctx->d_missing.as<uint2>(), 4u * (uint32_t)n, ctx->stream);
                    ctx->timings.mesh_launches += 1;
                }
                launch_mesh_count(ctx->d_jobs.as<uint4>(), ctx->d_mesh_nbr.as<uint4>(), n, ctx->d_voxels, ctx->d_planes,
                                  ctx->d_rec_count.as<uint32_t>(), ctx->d_face_count.as<uint32_t>(),
                                  ctx->d_warp_totals.as<uint32_t>(), ctx->stream);
                launch_mesh_offsets(ctx->d_rec_count.as<uint32_t>(), ctx->d_face_count.as<uint32_t>(), n,
                                    ctx->d_rec_first.as<uint32_t>(), ctx->d_face_first.as<uint32_t>(), counters,
                                    ctx->h_totals, ctx->stream);
            }
            ctx->timings.mesh_launches += 2;
            // Emit goes out right behind the scan, before the host knows the totals: it backs off
            // by itself if the stream does not fit the arena kept from earlier calls.  The host waits
            // for the scan only; in the common case it returns while emit is still running, and
            // whatever the caller enqueues next (vx_mesh_read, the next vx_generate) lines up behind it.
            VX_CUDA(cudaEventRecord(ctx->totals_ev, ctx->stream));
            uint32_t cap_faces, cap_recs;
            arena_caps(cap_faces, cap_recs);
            if (cap_faces && cap_recs) launch_emit(cap_faces, cap_recs);
            VX_CUDA(cudaGetLastError());
            VX_CUDA(cudaEventSynchronize(ctx->totals_ev));
            if (back[0] != 0) {
                sync_stream(ctx);
                return fail(ctx, VX_ERR_NOT_LOADED, "vx_mesh: area not resident");
            }
            if (back[1] == 0) {
                ctx->last_mesh_version = ctx->index_version;  // d_jobs resolves exactly this list
                n_recs = back[2];
                n_faces = back[3];
                if (n_faces > cap_faces || n_recs > cap_recs || !(cap_faces && cap_recs)) {
                    if ((st = sync_stream(ctx)) != VX_OK) return st;  // the arena is about to be replaced
                    VX_CUDA(ctx->d_recs.ensure(std::max<size_t>(16, n_recs * 16)));
                    VX_CUDA(ctx->d_vertices.ensure(std::max<size_t>(16, n_faces * 160)));
                    VX_CUDA(ctx->d_indices.ensure(std::max<size_t>(16, n_faces * 12)));
                    arena_caps(cap_faces, cap_recs);
                    launch_emit(cap_faces, cap_recs);
                    VX_CUDA(cudaGetLastError());
                    if ((st = sync_stream(ctx)) != VX_OK) return st;
                }
                break;
            }
            VX_CUDA(cudaStreamSynchronize(ctx->stream));
            VX_CUDA(cudaGetLastError());
            collect_timers(ctx);
            // edge neighbours that are not resident are generated, like world.get_with_cache ->
            // load_area (world.rs:157-174), then the pass above is repeated
            if (attempt > 0) return fail(ctx, VX_ERR_CUDA, "vx_mesh: neighbours still missing after generation");
            if (!ctx->has_seed) return fail(ctx, VX_ERR_NO_SEED, "vx_mesh needs to generate edge neighbours but no seed is set");
            const uint32_t n_missing = back[1];
            const uint32_t nm = std::min<uint32_t>(n_missing, 4u * (uint32_t)n);
            uint2* miss = reinterpret_cast<uint2*>(ctx->h_small.as<uint32_t>() + 16);
            VX_CUDA(cudaMemcpyAsync(miss, ctx->d_missing.p, (size_t)nm * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
            VX_CUDA(cudaStreamSynchronize(ctx->stream));
            std::vector<uint64_t> keys(nm);
            for (uint32_t i = 0; i < nm; ++i) keys[i] = area_key(miss[i].x, miss[i].y);
            std::sort(keys.begin(), keys.end());
            keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
            std::vector<uint32_t> mxy(2 * keys.size());
            for (size_t i = 0; i < keys.size(); ++i) {
                mxy[2 * i] = (uint32_t)(keys[i] >> 32);
                mxy[2 * i + 1] = (uint32_t)keys[i];
            }
            std::vector<int> mslots;
            if ((st = acquire_slots(ctx, mxy.data(), (int)keys.size(), mslots)) != VX_OK) return st;
            if ((st = generate_async(ctx, mxy.data(), mslots, (int)keys.size())) != VX_OK) return st;
            VX_CUDA(cudaStreamSynchronize(ctx->stream));
            collect_timers(ctx);
            jobs_ready = false;  // the resident set changed: resolve again
        }
        ctx->mesh_info.n_faces = n_faces;
        ctx->mesh_info.n_recs = n_recs;
        if (ctx->retention) {  // RenderArea of every listed area := the records just emitted
            launch_mesh_retain(ctx->d_jobs.as<uint4>(), n, ctx->d_rec_first.as<uint32_t>(), ctx->d_recs.as<uint4>(),
                               ctx->d_face, ctx->d_meshed, ctx->stream);
            ctx->timings.mesh_launches += 1;
            VX_CUDA(cudaGetLastError());
        }
    }
    ctx->mesh_info.vertex_bytes = ctx->mesh_info.n_faces * 160;
    ctx->mesh_info.index_bytes = ctx->mesh_info.n_faces * 12;
    ctx->mesh_info.rec_bytes = ctx->mesh_info.n_recs * 16;
    ctx->mesh_n = n;
    if (info) *info = ctx->mesh_info;
    return VX_OK;
}

int vx_mesh_read(vx_ctx* ctx, uint32_t* rec_first, uint32_t* face_first, vx_voxel_rec* recs,
                 void* vertices, uint16_t* indices) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (ctx->mesh_n < 0) return fail(ctx, VX_ERR_NO_MESH, "vx_mesh_read without vx_mesh");
    VX_CUDA(cudaSetDevice(ctx->device));
    const int n = ctx->mesh_n;
    if (n == 0) {
        if (rec_first) rec_first[0] = 0;
        if (face_first) face_first[0] = 0;
        return VX_OK;
    }
    const vx_mesh_info& mi = ctx->mesh_info;
    if (rec_first) VX_CUDA(cudaMemcpyAsync(rec_first, ctx->d_rec_first.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (face_first) VX_CUDA(cudaMemcpyAsync(face_first, ctx->d_face_first.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (recs && mi.rec_bytes) VX_CUDA(cudaMemcpyAsync(recs, ctx->d_recs.p, mi.rec_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (vertices && mi.vertex_bytes) VX_CUDA(cudaMemcpyAsync(vertices, ctx->d_vertices.p, mi.vertex_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (indices && mi.index_bytes) VX_CUDA(cudaMemcpyAsync(indices, ctx->d_indices.p, mi.index_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_encode_areas(vx_ctx* ctx, const uint32_t* area_xy, int n, uint64_t* total_bytes) {
    if (!ctx || n < 0 || (n > 0 && !area_xy) || !total_bytes) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    *total_bytes = 0;
    ctx->codec_n = -1;
    ctx->codec_total = 0;
    if (n == 0) {
        ctx->codec_n = 0;
        return VX_OK;
    }
    std::vector<int> slots(n);
    for (int i = 0; i < n; ++i) {
        slots[i] = host_slot(ctx, area_xy[2 * i], area_xy[2 * i + 1]);
        if (slots[i] < 0) return fail(ctx, VX_ERR_NOT_LOADED, "vx_encode_areas: area not resident");
    }
    // the files of the previous call may still be crossing to the host (deferred read)
    if (ctx->files_pending) VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->files_ev, 0));
    int st;
    if ((st = upload_jobs(ctx, area_xy, slots, n)) != VX_OK) return st;
    VX_CUDA(ctx->d_codec_sizes.ensure((size_t)n * 4));
    VX_CUDA(ctx->d_codec_offsets.ensure((size_t)(n + 1) * 8));
    VX_CUDA(ctx->d_codec_kinds.ensure((size_t)n * (VOXELS_IN_AREA / 16)));
    VX_CUDA(ctx->h_small.ensure(64));
    unsigned long long* offsets = ctx->d_codec_offsets.as<unsigned long long>();
    {
        ScopedTimer t(ctx, TM_CODEC);
        launch_codec_sizes(ctx->d_jobs.as<uint4>(), n, ctx->d_voxels, ctx->d_codec_kinds.as<uint8_t>(),
                           ctx->d_codec_sizes.as<uint32_t>(), offsets, ctx->stream);
    }
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(ctx->h_small.p, offsets + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    // The emit pass goes out right behind the size pass, before the host knows the total, when the
    // buffer kept from earlier calls may do: it backs off by itself if the files do not fit.  The host
    // waits for the total only.  (With the timers on the passes are issued one after the other.)
    const size_t cap_before = ctx->d_codec_bytes.cap;
#ifdef VX_ENCODE_NO_SPECULATION  // A/B switch (profiles/r02_ab_encode_spec.log)
    const bool speculative = false;
#else
    const bool speculative = !ctx->timing && cap_before > 32;
#endif
    if (speculative) {
        VX_CUDA(cudaEventRecord(ctx->totals_ev, ctx->stream));
        launch_codec_emit(ctx->d_jobs.as<uint4>(), n, ctx->d_voxels, ctx->d_codec_kinds.as<uint8_t>(), offsets,
                          ctx->d_codec_bytes.as<uint8_t>(), cap_before - 32, ctx->stream);
        VX_CUDA(cudaGetLastError());
        VX_CUDA(cudaEventSynchronize(ctx->totals_ev));
    } else {
        if ((st = sync_stream(ctx)) != VX_OK) return st;
        collect_timers(ctx);
    }
    const uint64_t total = *ctx->h_small.as<uint64_t>();
    if (!speculative || total + 32 > cap_before) {
        if (speculative && (st = sync_stream(ctx)) != VX_OK) return st;  // the buffer is about to be replaced
        VX_CUDA(ctx->d_codec_bytes.ensure(total + 32));  // the decoder stages whole 16-byte pieces: slack past the last file
        ScopedTimer t(ctx, TM_CODEC);
        launch_codec_emit(ctx->d_jobs.as<uint4>(), n, ctx->d_voxels, ctx->d_codec_kinds.as<uint8_t>(), offsets,
                          ctx->d_codec_bytes.as<uint8_t>(), ctx->d_codec_bytes.cap, ctx->stream);
    }
    ctx->timings.other_launches += 3;
    VX_CUDA(cudaGetLastError());
    // the sizes are known; the files are being written: vx_encode_read waits for this event only, so a
    // vx_mesh issued in between runs while the files cross to the host
    VX_CUDA(cudaEventRecord(ctx->codec_ev, ctx->stream));
    ctx->codec_n = n;
    ctx->codec_total = total;
    *total_bytes = total;
    return VX_OK;
}

int vx_encode_read(vx_ctx* ctx, uint8_t* files, uint64_t* offsets) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (ctx->codec_n < 0) return fail(ctx, VX_ERR_INVALID, "vx_encode_read without vx_encode_areas");
    VX_CUDA(cudaSetDevice(ctx->device));
    if (ctx->codec_n == 0) {
        if (offsets) offsets[0] = 0;
        return VX_OK;
    }
    VX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->codec_ev, 0));
    if (files && ctx->codec_total)
        VX_CUDA(cudaMemcpyAsync(files, ctx->d_codec_bytes.p, ctx->codec_total, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (offsets)
        VX_CUDA(cudaMemcpyAsync(offsets, ctx->d_codec_offsets.p, (size_t)(ctx->codec_n + 1) * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (ctx->deferred_reads) {  // valid after vx_wait_reads
        VX_CUDA(cudaEventRecord(ctx->files_ev, ctx->copy_stream));
        ctx->files_pending = true;
        return VX_OK;
    }
    VX_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return VX_OK;
}

int vx_decode_areas(vx_ctx* ctx, const uint32_t* area_xy, int n, const uint8_t* files,
                    const uint64_t* offsets, uint8_t* ok, uint8_t* max_height_out) {
    if (!ctx || n < 0 || (n > 0 && (!area_xy || !offsets || !ok))) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return VX_OK;
    for (int i = 0; i < n; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, VX_ERR_INVALID, "vx_decode_areas: offsets must not decrease");
    const uint64_t total = offsets[n] - offsets[0];
    if (total && !files) return fail(ctx, VX_ERR_INVALID, "vx_decode_areas: files is NULL");
    std::vector<int> slots;
    int st = acquire_slots(ctx, area_xy, n, slots);
    if (st != VX_OK) return st;
    if ((st = upload_jobs(ctx, area_xy, slots, n)) != VX_OK) return st;
    VX_CUDA(ctx->d_codec_bytes.ensure(total + 32));  // the decoder stages whole 16-byte pieces: slack past the last file
    VX_CUDA(ctx->d_codec_offsets.ensure((size_t)(n + 1) * 8));
    VX_CUDA(ctx->d_codec_status.ensure((size_t)n));
    ctx->codec_n = -1;  // d_codec_bytes no longer holds the result of vx_encode_areas
    if (ctx->files_pending) VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->files_ev, 0));
    // offsets relative to the first file, as uploaded
    std::vector<uint64_t> rel(n + 1);
    for (int i = 0; i <= n; ++i) rel[i] = offsets[i] - offsets[0];
    if (total) VX_CUDA(cudaMemcpyAsync(ctx->d_codec_bytes.p, files + offsets[0], total, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(ctx->d_codec_offsets.p, rel.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        ScopedTimer t(ctx, TM_CODEC);
        launch_codec_decode(ctx->d_jobs.as<uint4>(), n, ctx->d_codec_bytes.as<uint8_t>(),
                            ctx->d_codec_offsets.as<unsigned long long>(), ctx->d_voxels,
                            ctx->d_codec_status.as<uint8_t>(), ctx->stream);
    }
    launch_column_heights(ctx->d_jobs.as<uint4>(), n, ctx->d_voxels, ctx->d_heights, ctx->d_slot_xy, ctx->d_planes,
                          ctx->stream);
    ctx->timings.other_launches += 2;
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(ok, ctx->d_codec_status.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (max_height_out && (st = copy_slots(ctx, slots, n, ctx->d_heights, COLUMNS_IN_AREA, max_height_out, true)) != VX_OK) return st;
    if ((st = sync_stream(ctx)) != VX_OK) return st;  // rel[] was read by the copy above: keep it alive until here
    collect_timers(ctx);
    bool any = false;
    for (int i = 0; i < n; ++i) {
        if (ok[i]) continue;
        const int s = ctx->index.erase(area_xy[2 * i], area_xy[2 * i + 1]);  // not resident after all
        if (s < 0) continue;
        ctx->slot_key[s] = ~0ull;
        ctx->free_slots.push_back(s);
        any = true;
    }
    if (any) {
        ctx->map_dirty = true;
        ++ctx->index_version;
        std::sort(ctx->free_slots.begin(), ctx->free_slots.end(), std::greater<int>());
    }
    return VX_OK;
}

int vx_visible(vx_ctx* ctx, const vx_view* view, vx_visible_info* info) {
    if (!ctx || !view || !info) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (ctx->mesh_n < 0) return fail(ctx, VX_ERR_NO_MESH, "vx_visible without vx_mesh");
    VX_CUDA(cudaSetDevice(ctx->device));
    *info = vx_visible_info{};
    ctx->frame_ready = false;
    ctx->frame_n_visible = 0;
    const int n = ctx->mesh_n;
    if (n == 0) {
        ctx->frame_areas = 0;
        ctx->frame_ready = true;
        return VX_OK;
    }
    const uint64_t n_recs = ctx->mesh_info.n_recs;
    VX_CUDA(ctx->d_frame_flags.ensure((size_t)n));
    VX_CUDA(ctx->d_frame_counts.ensure((size_t)28 * n * 4));
    VX_CUDA(ctx->d_frame_visible.ensure(std::max<size_t>(16, n_recs * 4)));
    VX_CUDA(ctx->d_frame_codes.ensure(std::max<size_t>(16, n_recs)));
    const bool fresh = ctx->d_frame_small.p == nullptr;
    VX_CUDA(ctx->d_frame_small.ensure(sizeof(FrameTotals) + 32 * 4));
    VX_CUDA(ctx->h_small.ensure(sizeof(FrameTotals)));
    FrameTotals* totals = ctx->d_frame_small.as<FrameTotals>();
    uint32_t* row_total = reinterpret_cast<uint32_t*>(totals + 1);
    uint32_t* done = row_total + 28;
    if (fresh) VX_CUDA(cudaMemsetAsync(ctx->d_frame_small.p, 0, sizeof(FrameTotals) + 32 * 4, ctx->stream));
    VX_CUDA(cudaMemsetAsync(totals, 0, 16, ctx->stream));  // n_faces, n_areas, n_lamps
    FrameView fv;
    fv.cx = view->cam_pos[0]; fv.cy = view->cam_pos[1]; fv.cz = view->cam_pos[2];
    fv.target_z = view->cam_target_z;
    fv.lx = view->look[0]; fv.ly = view->look[1]; fv.lz = view->look[2];
    fv.render_size = view->render_size;
    fv.head_in_water = view->head_in_water;
    fv.show_map = view->show_map;
    {
        ScopedTimer t(ctx, TM_VISIBLE);
        launch_frame_visible(ctx->d_recs.as<uint4>(), ctx->d_rec_first.as<uint32_t>(), ctx->d_mesh_xy.as<uint2>(), n, fv,
                             ctx->d_frame_flags.as<uint8_t>(), ctx->d_frame_counts.as<uint32_t>(), row_total, done,
                             totals, ctx->d_frame_visible.as<uint32_t>(), ctx->d_frame_codes.as<uint8_t>(), ctx->stream);
        ctx->timings.other_launches += 4;
    }
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(ctx->h_small.p, totals, sizeof(FrameTotals), cudaMemcpyDeviceToHost, ctx->stream));
    int st;
    if ((st = sync_stream(ctx)) != VX_OK) return st;
    collect_timers(ctx);
    const FrameTotals* h = ctx->h_small.as<FrameTotals>();
    info->n_visible = h->type_first[27];
    info->n_faces = h->n_faces;
    info->n_areas = h->n_areas;
    info->n_lamps = h->n_lamps;
    std::memcpy(info->type_first, h->type_first, sizeof info->type_first);
    const uint32_t nl = std::min<uint32_t>(h->n_lamps, VX_MAX_LIGHTS);
    std::memcpy(info->lamps, h->lamps, nl * sizeof(uint32_t));
    ctx->frame_n_visible = info->n_visible;
    ctx->frame_areas = n;
    ctx->frame_ready = true;
    return VX_OK;
}

int vx_visible_read(vx_ctx* ctx, uint32_t* rec_index, uint8_t* area_visible) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!ctx->frame_ready) return fail(ctx, VX_ERR_NO_MESH, "vx_visible_read without vx_visible");
    VX_CUDA(cudaSetDevice(ctx->device));
    if (ctx->frame_areas == 0) return VX_OK;
    if (rec_index && ctx->frame_n_visible)
        VX_CUDA(cudaMemcpyAsync(rec_index, ctx->d_frame_visible.p, ctx->frame_n_visible * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (area_visible)
        VX_CUDA(cudaMemcpyAsync(area_visible, ctx->d_frame_flags.p, (size_t)ctx->frame_areas, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_apply_edits(vx_ctx* ctx, const vx_edit* edits, int n, uint32_t* dirty_area_xy, int* n_dirty) {
    if (!ctx || n < 0 || (n > 0 && !edits) || (dirty_area_xy && !n_dirty)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    if (n_dirty) *n_dirty = 0;
    if (n == 0) return VX_OK;
    for (int i = 0; i < n; ++i) {  // all-or-nothing validation (reference: world_actions.rs guards)
        if (edits[i].gz >= (uint32_t)AREA_HEIGHT || edits[i].voxel >= (uint32_t)V_VARIANTS)
            return fail(ctx, VX_ERR_INVALID, "vx_apply_edits: z >= 128 or voxel id >= 27");
        if (host_slot(ctx, edits[i].gx / AREA_SIZE, edits[i].gy / AREA_SIZE) < 0)
            return fail(ctx, VX_ERR_NOT_LOADED, "vx_apply_edits: edited area not resident");
    }
    int st;
    if ((st = upload_map(ctx)) != VX_OK) return st;
    uint32_t cap = 1024;
    while (cap < (uint32_t)n * 2u) cap <<= 1;
    VX_CUDA(ctx->d_hash_keys.ensure((size_t)cap * 8));
    VX_CUDA(ctx->d_hash_vals.ensure((size_t)cap * 4));
    VX_CUDA(ctx->d_dirty_flags.ensure((size_t)ctx->capacity * 4));
    VX_CUDA(ctx->d_dirty_slots.ensure((size_t)ctx->capacity * 4));
    VX_CUDA(ctx->d_counters.ensure(64));
    VX_CUDA(ctx->d_edits.ensure((size_t)n * sizeof(uint4)));
    VX_CUDA(cudaMemsetAsync(ctx->d_hash_keys.p, 0xFF, (size_t)cap * 8, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_hash_vals.p, 0, (size_t)cap * 4, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_dirty_flags.p, 0, (size_t)ctx->capacity * 4, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
    static_assert(sizeof(vx_edit) == sizeof(uint4), "vx_edit layout");
    const void* edits_src = edits;
    if (!dirty_area_xy) {  // the call will not wait: the caller's array must not be read after it returns
        if (ctx->edits_pending) {
            VX_CUDA(cudaEventSynchronize(ctx->edits_ev));
            ctx->edits_pending = false;
        }
        VX_CUDA(ctx->h_edits.ensure((size_t)n * sizeof(uint4)));
        std::memcpy(ctx->h_edits.p, edits, (size_t)n * sizeof(uint4));
        edits_src = ctx->h_edits.p;
    }
    VX_CUDA(cudaMemcpyAsync(ctx->d_edits.p, edits_src, (size_t)n * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
    if (!dirty_area_xy) {
        VX_CUDA(cudaEventRecord(ctx->edits_ev, ctx->stream));
        ctx->edits_pending = true;
    }
    uint32_t* counters = ctx->d_counters.as<uint32_t>();
    {
        ScopedTimer t(ctx, TM_EDIT);
        launch_apply_edits(ctx->d_edits.as<uint4>(), n, device_map(ctx), ctx->d_voxels, ctx->d_heights, ctx->d_planes,
                           ctx->d_hash_keys.as<unsigned long long>(), ctx->d_hash_vals.as<uint32_t>(), cap,
                           ctx->d_dirty_flags.as<uint32_t>(), ctx->d_dirty_slots.as<uint32_t>(),
                           counters /* dirty_count */, reinterpret_cast<int32_t*>(counters + 1), ctx->stream);
    }
    ctx->timings.other_launches += 3;
    VX_CUDA(cudaGetLastError());
    // No dirty list wanted (the caller follows up with vx_update_locations, which needs none): everything
    // was validated on the host, so the call returns with the kernels enqueued, like vx_generate.
    if (!dirty_area_xy) return VX_OK;
    // the counters and as many dirty slots as the batch can have produced, in one wait
    const size_t max_dirty = std::min<size_t>((size_t)3 * n, (size_t)ctx->capacity);
    VX_CUDA(ctx->h_small.ensure(16 + max_dirty * 4));
    uint32_t* hs = ctx->h_small.as<uint32_t>();
    VX_CUDA(cudaMemcpyAsync(hs, counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(hs + 4, ctx->d_dirty_slots.p, max_dirty * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if ((st = sync_stream(ctx)) != VX_OK) return st;
    const uint32_t nd = hs[0];
    std::vector<uint32_t> ds(hs + 4, hs + 4 + nd);
    std::sort(ds.begin(), ds.end());  // deterministic order
    for (uint32_t i = 0; i < nd; ++i) {
        const uint64_t key = ctx->slot_key[ds[i]];
        dirty_area_xy[2 * i] = (uint32_t)(key >> 32);
        dirty_area_xy[2 * i + 1] = (uint32_t)key;
    }
    *n_dirty = (int)nd;
    return VX_OK;
}

int vx_explode(vx_ctx* ctx, const float* positions, int n, uint32_t* removed_xyz, int* n_removed,
               uint32_t* bomb_xyz, int* n_bombs, uint32_t* dirty_area_xy, int* n_dirty) {
    if (!ctx || n < 0 || n > 32 || !n_removed || !n_bombs || !n_dirty ||
        (n > 0 && (!positions || !removed_xyz || !bomb_xyz || !dirty_area_xy)))
        return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    *n_removed = *n_bombs = *n_dirty = 0;
    if (n == 0) return VX_OK;
    for (int i = 0; i < n; ++i) {  // all-or-nothing: every area a sphere can reach is resident
        const float px = positions[3 * i], py = positions[3 * i + 1], pz = positions[3 * i + 2];
        if (!(std::fabs(px) < 1.0e6f && std::fabs(py) < 1.0e6f && std::fabs(pz) < 1.0e6f))
            return fail(ctx, VX_ERR_INVALID, "vx_explode: position out of range");
        const int cx = (int)std::floor(px), cy = (int)std::floor(py);
        for (int dx = -5; dx <= 5; dx += 10)
            for (int dy = -5; dy <= 5; dy += 10) {
                const uint32_t gx = (uint32_t)(cx + dx + LOCATION_OFFSET), gy = (uint32_t)(cy + dy + LOCATION_OFFSET);
                if (host_slot(ctx, gx / AREA_SIZE, gy / AREA_SIZE) < 0)
                    return fail(ctx, VX_ERR_NOT_LOADED, "vx_explode: an area within reach of an explosion is not resident");
            }
    }
    int st;
    if ((st = upload_map(ctx)) != VX_OK) return st;
    const size_t cand = (size_t)n * EXPLOSION_CANDIDATES;
    uint32_t cap = 1024;
    while (cap < cand * 2u) cap <<= 1;
    VX_CUDA(ctx->d_hash_keys.ensure((size_t)cap * 8));
    VX_CUDA(ctx->d_hash_vals.ensure((size_t)cap * 4));
    VX_CUDA(ctx->d_dirty_flags.ensure((size_t)ctx->capacity * 4));
    VX_CUDA(ctx->d_dirty_slots.ensure((size_t)ctx->capacity * 4));
    VX_CUDA(ctx->d_counters.ensure(64));
    // scratch: positions | flags | removed | bombs
    const size_t pos_bytes = ((size_t)n * 12 + 15) & ~(size_t)15, flag_bytes = (cand + 15) & ~(size_t)15;
    VX_CUDA(ctx->d_edits.ensure(pos_bytes + flag_bytes + 2 * cand * 12));
    uint8_t* base = ctx->d_edits.as<uint8_t>();
    float* d_pos = reinterpret_cast<float*>(base);
    uint8_t* d_flags = base + pos_bytes;
    uint32_t* d_removed = reinterpret_cast<uint32_t*>(base + pos_bytes + flag_bytes);
    uint32_t* d_bombs = d_removed + 3 * cand;
    VX_CUDA(cudaMemsetAsync(ctx->d_hash_keys.p, 0xFF, (size_t)cap * 8, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_hash_vals.p, 0, (size_t)cap * 4, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_dirty_flags.p, 0, (size_t)ctx->capacity * 4, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(d_pos, positions, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* counters = ctx->d_counters.as<uint32_t>();  // [0] dirty, [1] status, [2] removed, [3] bombs
    {
        ScopedTimer t(ctx, TM_EDIT);
        launch_explode(d_pos, n, device_map(ctx), ctx->d_voxels, ctx->d_heights, ctx->d_planes,
                       ctx->d_hash_keys.as<unsigned long long>(), ctx->d_hash_vals.as<uint32_t>(), cap,
                       ctx->d_dirty_flags.as<uint32_t>(), ctx->d_dirty_slots.as<uint32_t>(), counters,
                       reinterpret_cast<int32_t*>(counters + 1), d_flags, d_removed, d_bombs, counters + 2, ctx->stream);
    }
    ctx->timings.other_launches += 3;
    VX_CUDA(cudaGetLastError());
    VX_CUDA(ctx->h_small.ensure(16 + (size_t)ctx->capacity * 4));
    uint32_t* hs = ctx->h_small.as<uint32_t>();
    VX_CUDA(cudaMemcpyAsync(hs, counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t nd = hs[0], nr = hs[2], nb = hs[3];
    if (nd) VX_CUDA(cudaMemcpyAsync(hs + 4, ctx->d_dirty_slots.p, (size_t)nd * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (nr) VX_CUDA(cudaMemcpyAsync(removed_xyz, d_removed, (size_t)nr * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (nb) VX_CUDA(cudaMemcpyAsync(bomb_xyz, d_bombs, (size_t)nb * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if ((st = sync_stream(ctx)) != VX_OK) return st;
    collect_timers(ctx);
    std::vector<uint32_t> ds(hs + 4, hs + 4 + nd);
    std::sort(ds.begin(), ds.end());  // deterministic order
    for (uint32_t i = 0; i < nd; ++i) {
        const uint64_t key = ctx->slot_key[ds[i]];
        dirty_area_xy[2 * i] = (uint32_t)(key >> 32);
        dirty_area_xy[2 * i + 1] = (uint32_t)key;
    }
    *n_dirty = (int)nd;
    *n_removed = (int)nr;
    *n_bombs = (int)nb;
    return VX_OK;
}

int vx_heightmap(vx_ctx* ctx, const uint32_t* area_xy, int n, uint32_t cam_area_x,
                 uint32_t cam_area_y, uint8_t* rgba) {
    if (!ctx || n < 0 || (n > 0 && !area_xy) || !rgba) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    const size_t img_bytes = (size_t)VX_HEIGHTMAP_SIZE * VX_HEIGHTMAP_SIZE * 4;
    int st;
    if ((st = upload_map(ctx)) != VX_OK) return st;
    if (!ctx->heightmap_init) {
        VX_CUDA(ctx->d_heightmap.ensure(img_bytes));
        VX_CUDA(cudaMemsetAsync(ctx->d_heightmap.p, 0xFF, img_bytes, ctx->stream));  // Image::gen_image_color(WHITE)
        ctx->heightmap_init = true;
    }
    if (n > 0) {
        if (int st = stage_ready(ctx)) return st;
        VX_CUDA(ctx->h_stage.ensure((size_t)n * 8));
        VX_CUDA(ctx->d_list.ensure((size_t)n * 8));
        std::memcpy(ctx->h_stage.p, area_xy, (size_t)n * 8);
        VX_CUDA(cudaMemcpyAsync(ctx->d_list.p, ctx->h_stage.p, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        ScopedTimer t(ctx, TM_HEIGHTMAP);
        launch_heightmap(ctx->d_list.as<uint2>(), n, device_map(ctx), ctx->d_heights, cam_area_x, cam_area_y,
                         ctx->d_heightmap.as<uint32_t>(), ctx->stream);
        ctx->timings.other_launches += 1;
    }
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(rgba, ctx->d_heightmap.p, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_stream(ctx);
}

int vx_light(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* light_out, int* iterations) {
    if (!ctx || n < 0 || (n > 0 && (!area_xy || !light_out))) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    if (iterations) *iterations = 0;
    if (n == 0) return VX_OK;
    std::vector<int> slots(n);
    for (int i = 0; i < n; ++i) {
        slots[i] = host_slot(ctx, area_xy[2 * i], area_xy[2 * i + 1]);
        if (slots[i] < 0) return fail(ctx, VX_ERR_NOT_LOADED, "vx_light: area not resident");
    }
    int st;
    if ((st = upload_map(ctx)) != VX_OK) return st;
    if ((st = upload_jobs(ctx, area_xy, slots, n)) != VX_OK) return st;
    const size_t bytes = (size_t)n * VOXELS_IN_AREA;
    VX_CUDA(ctx->d_light.ensure(bytes));
    VX_CUDA(ctx->d_slot_to_job.ensure((size_t)ctx->capacity * sizeof(int32_t)));
    VX_CUDA(ctx->d_counters.ensure(64));
    VX_CUDA(ctx->h_small.ensure(64));
    VX_CUDA(cudaMemsetAsync(ctx->d_slot_to_job.p, 0xFF, (size_t)ctx->capacity * sizeof(int32_t), ctx->stream));
    uint32_t* flag = ctx->d_counters.as<uint32_t>();
    uint32_t* host_flag = ctx->h_small.as<uint32_t>();
    const uint4* jobs = ctx->d_jobs.as<uint4>();
    int iters = 0;
    {
        ScopedTimer t(ctx, TM_LIGHT);
        launch_light_init(jobs, n, ctx->d_voxels, ctx->d_heights, ctx->d_light.as<uint8_t>(),
                          ctx->d_slot_to_job.as<int32_t>(), ctx->stream);
        ctx->timings.other_launches += 2;
        // levels travel at most 15 voxels, i.e. across at most one area border per launch pair;
        // the bound below is generous and only guards against a logic error
        // Two launches per look at the flag: the last launch of a run only confirms that nothing moves
        // any more, so the host would otherwise wait once per launch for an answer that is "go on" every
        // time but the last.
        while (iters < 64) {
            VX_CUDA(cudaMemsetAsync(flag, 0, 8, ctx->stream));
            for (int k = 0; k < 2; ++k) {
                launch_light_relax(jobs, n, ctx->d_voxels, device_map(ctx), ctx->d_slot_to_job.as<int32_t>(),
                                   ctx->d_light.as<uint8_t>(), flag + k, ctx->stream);
                ctx->timings.other_launches += 1;
            }
            iters += 2;
            VX_CUDA(cudaMemcpyAsync(host_flag, flag, 8, cudaMemcpyDeviceToHost, ctx->stream));
            VX_CUDA(cudaStreamSynchronize(ctx->stream));
            if (host_flag[1] == 0) {  // the second launch of the pair changed nothing: the fixed point
                if (host_flag[0] == 0) --iters;  // ... and it was already reached before the pair's second launch
                break;
            }
        }
    }
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(light_out, ctx->d_light.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (iterations) *iterations = iters;
    return sync_stream(ctx);
}

int vx_set_mesh_retention(vx_ctx* ctx, int enabled) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    const bool on = enabled != 0;
    if (on == ctx->retention) return VX_OK;
    int st;
    if ((st = sync_stream(ctx)) != VX_OK) return st;
    if (ctx->vmm) {
        if (on && ctx->capacity > 0) {
            cudaError_t e = ctx->vm_face.grow(ctx->vm, ctx->device, (size_t)ctx->capacity * VOXELS_IN_AREA);
            if (e == cudaSuccess) e = ctx->vm_meshed.grow(ctx->vm, ctx->device, (size_t)ctx->capacity);
            if (e != cudaSuccess) {
                cudaGetLastError();
                ctx->vm_face.unmap_all(ctx->vm);
                ctx->vm_meshed.unmap_all(ctx->vm);
                return fail(ctx, VX_ERR_OOM, "mesh retention allocation");
            }
            VX_CUDA(cudaMemsetAsync(ctx->d_face, 0, (size_t)ctx->capacity * VOXELS_IN_AREA, ctx->stream));
            VX_CUDA(cudaMemsetAsync(ctx->d_meshed, 0, (size_t)ctx->capacity, ctx->stream));
        } else if (!on) {  // give the physical memory back, keep the address range
            ctx->vm_face.unmap_all(ctx->vm);
            ctx->vm_meshed.unmap_all(ctx->vm);
        }
    } else if (on) {
        if (ctx->capacity > 0) {
            VX_CUDA(cudaMalloc(&ctx->d_face, (size_t)ctx->capacity * VOXELS_IN_AREA));
            cudaError_t e = cudaMalloc(&ctx->d_meshed, (size_t)ctx->capacity);
            if (e != cudaSuccess) {
                cudaFree(ctx->d_face);
                ctx->d_face = nullptr;
                return fail_cuda(ctx, e, "mesh retention allocation");
            }
            VX_CUDA(cudaMemsetAsync(ctx->d_face, 0, (size_t)ctx->capacity * VOXELS_IN_AREA, ctx->stream));
            VX_CUDA(cudaMemsetAsync(ctx->d_meshed, 0, (size_t)ctx->capacity, ctx->stream));
        }
    } else {
        cudaFree(ctx->d_face);
        cudaFree(ctx->d_meshed);
        ctx->d_face = ctx->d_meshed = nullptr;
    }
    ctx->retention = on;
    return VX_OK;
}

int vx_mesh_unload(vx_ctx* ctx, const uint32_t* area_xy, int n) {
    if (!ctx || n < 0 || (n > 0 && !area_xy)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!ctx->retention || n == 0) return VX_OK;  // nothing is kept on the device
    VX_CUDA(cudaSetDevice(ctx->device));
    std::vector<int> slots;
    for (int i = 0; i < n; ++i) {
        const int s = host_slot(ctx, area_xy[2 * i], area_xy[2 * i + 1]);
        if (s >= 0) slots.push_back(s);
    }
    return clear_mesh_slots(ctx, slots);
}

int vx_read_face_masks(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* masks_out, uint8_t* meshed_out) {
    if (!ctx || n < 0 || (n > 0 && !area_xy)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!ctx->retention) return fail(ctx, VX_ERR_NO_MESH, "vx_read_face_masks: mesh retention is off");
    VX_CUDA(cudaSetDevice(ctx->device));
    std::vector<int> slots(n);
    for (int i = 0; i < n; ++i) {
        slots[i] = host_slot(ctx, area_xy[2 * i], area_xy[2 * i + 1]);
        if (slots[i] < 0) return fail(ctx, VX_ERR_NOT_LOADED, "vx_read_face_masks: area not resident");
    }
    int st;
    if (masks_out && (st = copy_slots(ctx, slots, n, ctx->d_face, VOXELS_IN_AREA, masks_out, true)) != VX_OK) return st;
    if (meshed_out && (st = copy_slots(ctx, slots, n, ctx->d_meshed, 1, meshed_out, true)) != VX_OK) return st;
    return sync_stream(ctx);
}

int vx_update_locations(vx_ctx* ctx, const uint32_t* xyz, int n, vx_mesh_info* info) {
    if (!ctx || n < 0 || (n > 0 && !xyz)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    ctx->upd_n = -1;
    ctx->upd_info = vx_mesh_info{};
    if (info) *info = ctx->upd_info;
    if (n == 0) {
        ctx->upd_n = 0;
        return VX_OK;
    }
    for (int i = 0; i < n; ++i)
        if (xyz[3 * i + 2] >= (uint32_t)AREA_HEIGHT) return fail(ctx, VX_ERR_INVALID, "vx_update_locations: z >= 128");
    if (ctx->capacity == 0) {  // nothing is loaded: get_without_loading -> None for every location
        ctx->upd_n = n;
        return VX_OK;
    }
    int st;
    const uint32_t n_cand = 7u * (uint32_t)n, n_blocks = (n_cand + 255u) / 256u;
    uint32_t cap = 1024;
    while (cap < n_cand * 2u) cap <<= 1;
    VX_CUDA(ctx->d_hash_keys.ensure((size_t)cap * 8));
    VX_CUDA(ctx->d_hash_vals.ensure((size_t)cap * 4));
    VX_CUDA(ctx->d_counters.ensure(64));
    VX_CUDA(ctx->d_upd_xyz.ensure((size_t)n * 12));
    VX_CUDA(ctx->d_upd_slot.ensure((size_t)n_cand * 4));
    VX_CUDA(ctx->d_upd_info.ensure((size_t)n_cand * 2));
    VX_CUDA(ctx->d_upd_rec_count.ensure((size_t)(n_blocks + 16) * 4));
    VX_CUDA(ctx->d_upd_face_count.ensure((size_t)(n_blocks + 16) * 4));
    VX_CUDA(ctx->d_upd_rec_first.ensure((size_t)(n_blocks + 17) * 4));
    VX_CUDA(ctx->d_upd_face_first.ensure((size_t)(n_blocks + 17) * 4));
    if ((st = upload_map(ctx)) != VX_OK) return st;
    VX_CUDA(cudaMemsetAsync(ctx->d_hash_keys.p, 0xFF, (size_t)cap * 8, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_hash_vals.p, 0xFF, (size_t)cap * 4, ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
    VX_CUDA(cudaMemcpyAsync(ctx->d_upd_xyz.p, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t* d_xyz = ctx->d_upd_xyz.as<uint32_t>();
    uint32_t* counters = ctx->d_counters.as<uint32_t>();
    ScopedTimer t(ctx, TM_EDIT);
    // which voxels get re-meshed is decided on the areas resident NOW (get_without_loading) ...
    launch_update_claim(d_xyz, n, device_map(ctx), ctx->d_hash_keys.as<unsigned long long>(),
                        ctx->d_hash_vals.as<uint32_t>(), cap, ctx->d_upd_slot.as<int32_t>(), ctx->stream);
    // ... while their face tests may reach one voxel further into areas that are not resident: those
    // are generated first, as world.get_with_cache -> load_area does (world.rs:157-174).  Only
    // locations within two voxels of an area border can need that.
    {
        std::vector<uint64_t> missing;
        uint64_t last_key = ~0ull;
        bool last_resident = false;
        auto resident = [&](uint32_t ax, uint32_t ay) {
            const uint64_t k = area_key(ax, ay);
            if (k != last_key) {
                last_key = k;
                last_resident = host_slot(ctx, ax, ay) >= 0;
            }
            return last_resident;
        };
        static const int CAND[5][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}};
        for (int i = 0; i < n; ++i) {
            const uint32_t gx = xyz[3 * i], gy = xyz[3 * i + 1];
            const uint32_t lx = gx % AREA_SIZE, ly = gy % AREA_SIZE;
            if (lx >= 2 && lx <= 13 && ly >= 2 && ly <= 13) continue;
            for (const auto& c : CAND) {
                const uint32_t vx_ = gx + (uint32_t)c[0], vy_ = gy + (uint32_t)c[1];
                if (!resident(vx_ / AREA_SIZE, vy_ / AREA_SIZE)) continue;  // the candidate itself is skipped
                for (const auto& d : CAND) {
                    if (d[0] == 0 && d[1] == 0) continue;
                    const uint32_t ax = (vx_ + (uint32_t)d[0]) / AREA_SIZE, ay = (vy_ + (uint32_t)d[1]) / AREA_SIZE;
                    if (!resident(ax, ay)) missing.push_back(area_key(ax, ay));
                }
            }
        }
        if (!missing.empty()) {
            if (!ctx->has_seed) return fail(ctx, VX_ERR_NO_SEED, "vx_update_locations needs to generate a neighbouring area but no seed is set");
            std::sort(missing.begin(), missing.end());
            missing.erase(std::unique(missing.begin(), missing.end()), missing.end());
            std::vector<uint32_t> mxy(2 * missing.size());
            for (size_t i = 0; i < missing.size(); ++i) {
                mxy[2 * i] = (uint32_t)(missing[i] >> 32);
                mxy[2 * i + 1] = (uint32_t)missing[i];
            }
            std::vector<int> mslots;
            if ((st = acquire_slots(ctx, mxy.data(), (int)missing.size(), mslots)) != VX_OK) return st;
            if ((st = generate_async(ctx, mxy.data(), mslots, (int)missing.size())) != VX_OK) return st;
            if ((st = upload_map(ctx)) != VX_OK) return st;
        }
    }
    launch_update_count(d_xyz, n, device_map(ctx), ctx->d_voxels, ctx->d_hash_keys.as<unsigned long long>(),
                        ctx->d_hash_vals.as<uint32_t>(), cap, ctx->d_upd_slot.as<int32_t>(),
                        ctx->d_upd_info.as<uint16_t>(), ctx->d_upd_rec_count.as<uint32_t>(),
                        ctx->d_upd_face_count.as<uint32_t>(), reinterpret_cast<int32_t*>(counters + 1), ctx->stream);
    launch_mesh_offsets(ctx->d_upd_rec_count.as<uint32_t>(), ctx->d_upd_face_count.as<uint32_t>(), (int)n_blocks,
                        ctx->d_upd_rec_first.as<uint32_t>(), ctx->d_upd_face_first.as<uint32_t>(), counters,
                        ctx->h_totals, ctx->stream);
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaEventRecord(ctx->totals_ev, ctx->stream));
    VX_CUDA(cudaEventSynchronize(ctx->totals_ev));
    const volatile uint32_t* back = ctx->h_totals;  // [1] status, [2] records, [3] faces
    if (back[1] != 0) {
        sync_stream(ctx);
        return fail(ctx, VX_ERR_NOT_LOADED, "vx_update_locations: a neighbouring area is not resident");
    }
    const uint64_t n_recs = back[2], n_faces = back[3];
    VX_CUDA(ctx->d_upd_recs.ensure(std::max<size_t>(16, n_recs * 16)));
    VX_CUDA(ctx->d_upd_vertices.ensure(std::max<size_t>(16, n_faces * 160)));
    VX_CUDA(ctx->d_upd_indices.ensure(std::max<size_t>(16, n_faces * 12)));
    MeshOut out;
    out.rec_first = ctx->d_upd_rec_first.as<uint32_t>();
    out.face_first = ctx->d_upd_face_first.as<uint32_t>();
    out.recs = ctx->d_upd_recs.as<uint4>();
    out.vertices = ctx->d_upd_vertices.as<uint4>();
    out.indices = ctx->d_upd_indices.as<uint32_t>();
    launch_update_emit(d_xyz, n, ctx->d_upd_slot.as<int32_t>(), ctx->d_upd_info.as<uint16_t>(), out.rec_first,
                       out.face_first, out, ctx->retention ? ctx->d_face : nullptr, ctx->d_meshed, ctx->stream);
    ctx->timings.other_launches += 4;
    VX_CUDA(cudaGetLastError());
    ctx->upd_info.n_recs = n_recs;
    ctx->upd_info.n_faces = n_faces;
    ctx->upd_info.rec_bytes = n_recs * 16;
    ctx->upd_info.vertex_bytes = n_faces * 160;
    ctx->upd_info.index_bytes = n_faces * 12;
    ctx->upd_n = n;
    if (info) *info = ctx->upd_info;
    return VX_OK;
}

int vx_update_read(vx_ctx* ctx, vx_voxel_rec* recs, void* vertices, uint16_t* indices) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (ctx->upd_n < 0) return fail(ctx, VX_ERR_NO_MESH, "vx_update_read without vx_update_locations");
    VX_CUDA(cudaSetDevice(ctx->device));
    const vx_mesh_info& mi = ctx->upd_info;
    if (recs && mi.rec_bytes) VX_CUDA(cudaMemcpyAsync(recs, ctx->d_upd_recs.p, mi.rec_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (vertices && mi.vertex_bytes) VX_CUDA(cudaMemcpyAsync(vertices, ctx->d_upd_vertices.p, mi.vertex_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (indices && mi.index_bytes) VX_CUDA(cudaMemcpyAsync(indices, ctx->d_upd_indices.p, mi.index_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_stream(ctx);
}

int vx_mesh_read_compact(vx_ctx* ctx, uint32_t* rec_first, uint32_t* packed_recs) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (ctx->mesh_n < 0) return fail(ctx, VX_ERR_NO_MESH, "vx_mesh_read_compact without vx_mesh");
    VX_CUDA(cudaSetDevice(ctx->device));
    const int n = ctx->mesh_n;
    if (n == 0) {
        if (rec_first) rec_first[0] = 0;
        return VX_OK;
    }
    const uint64_t n_recs = ctx->mesh_info.n_recs;
    if (ctx->deferred_reads) {
        // the pack kernel stays on the main stream, the copies go to the copy stream behind it: whatever the
        // caller enqueues next (the next vx_generate) runs while they cross to the host
        if (packed_recs && n_recs) {
            if (ctx->packed_pending) VX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->packed_ev, 0));  // d_packed is reused
            VX_CUDA(ctx->d_packed.ensure(n_recs * 4));
            launch_mesh_pack(ctx->d_recs.as<uint4>(), n_recs, ctx->d_packed.as<uint32_t>(), ctx->stream);
            ctx->timings.mesh_launches += 1;
            VX_CUDA(cudaGetLastError());
        }
        VX_CUDA(cudaEventRecord(ctx->pack_done_ev, ctx->stream));
        VX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->pack_done_ev, 0));
        if (rec_first) VX_CUDA(cudaMemcpyAsync(rec_first, ctx->d_rec_first.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (packed_recs && n_recs)
            VX_CUDA(cudaMemcpyAsync(packed_recs, ctx->d_packed.p, n_recs * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        VX_CUDA(cudaEventRecord(ctx->packed_ev, ctx->copy_stream));
        ctx->packed_pending = true;
        return VX_OK;
    }
    if (rec_first) VX_CUDA(cudaMemcpyAsync(rec_first, ctx->d_rec_first.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (packed_recs && n_recs) {
        VX_CUDA(ctx->d_packed.ensure(n_recs * 4));
        launch_mesh_pack(ctx->d_recs.as<uint4>(), n_recs, ctx->d_packed.as<uint32_t>(), ctx->stream);
        ctx->timings.mesh_launches += 1;
        VX_CUDA(cudaGetLastError());
        VX_CUDA(cudaMemcpyAsync(packed_recs, ctx->d_packed.p, n_recs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_set_deferred_reads(vx_ctx* ctx, int enabled) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!enabled) {
        const int st = vx_wait_reads(ctx);
        if (st != VX_OK) return st;
    }
    ctx->deferred_reads = enabled != 0;
    return VX_OK;
}

int vx_wait_reads(vx_ctx* ctx) {
    if (!ctx) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    if (ctx->files_pending) VX_CUDA(cudaEventSynchronize(ctx->files_ev));
    if (ctx->packed_pending) VX_CUDA(cudaEventSynchronize(ctx->packed_ev));
    if (ctx->heights_pending) VX_CUDA(cudaEventSynchronize(ctx->heights_ev));
    ctx->files_pending = ctx->packed_pending = ctx->heights_pending = false;
    VX_CUDA(cudaGetLastError());
    return VX_OK;
}

int vx_visible_zone(vx_ctx* ctx, const uint32_t* area_xy, int n, const vx_view* view, vx_visible_info* info) {
    if (!ctx || !view || !info || n < 0 || (n > 0 && !area_xy) || n > (1 << 17)) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!ctx->retention) return fail(ctx, VX_ERR_NO_MESH, "vx_visible_zone: mesh retention is off");
    VX_CUDA(cudaSetDevice(ctx->device));
    *info = vx_visible_info{};
    ctx->frame_ready = false;
    ctx->frame_n_visible = 0;
    ctx->frame_areas = n;
    if (n == 0 || ctx->capacity == 0) {
        ctx->frame_areas = 0;
        ctx->frame_ready = true;
        return VX_OK;
    }
    int st;
    if ((st = upload_map(ctx)) != VX_OK) return st;
    if ((st = stage_ready(ctx)) != VX_OK) return st;
    const size_t xy_bytes = (size_t)n * sizeof(uint2);
    VX_CUDA(ctx->h_stage.ensure(xy_bytes));
    VX_CUDA(ctx->d_zone_xy.ensure(xy_bytes));
    VX_CUDA(ctx->d_zone_slot.ensure((size_t)n * 4));
    VX_CUDA(ctx->d_frame_flags.ensure((size_t)n));
    VX_CUDA(ctx->d_frame_counts.ensure((size_t)28 * n * 4));
    // every meshed voxel of every listed area may be drawn (map view): 4 B each; grown on demand below
    const bool fresh = ctx->d_frame_small.p == nullptr;
    VX_CUDA(ctx->d_frame_small.ensure(sizeof(FrameTotals) + 32 * 4));
    VX_CUDA(ctx->h_small.ensure(sizeof(FrameTotals)));
    FrameTotals* totals = ctx->d_frame_small.as<FrameTotals>();
    uint32_t* row_total = reinterpret_cast<uint32_t*>(totals + 1);
    uint32_t* done = row_total + 28;
    if (fresh) VX_CUDA(cudaMemsetAsync(ctx->d_frame_small.p, 0, sizeof(FrameTotals) + 32 * 4, ctx->stream));
    std::memcpy(ctx->h_stage.p, area_xy, xy_bytes);
    VX_CUDA(cudaMemcpyAsync(ctx->d_zone_xy.p, ctx->h_stage.p, xy_bytes, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(cudaEventRecord(ctx->stage_ev, ctx->stream));
    ctx->stage_pending = true;
    FrameView fv;
    fv.cx = view->cam_pos[0]; fv.cy = view->cam_pos[1]; fv.cz = view->cam_pos[2];
    fv.target_z = view->cam_target_z;
    fv.lx = view->look[0]; fv.ly = view->look[1]; fv.lz = view->look[2];
    fv.render_size = view->render_size;
    fv.head_in_water = view->head_in_water;
    fv.show_map = view->show_map;
    // the index list can hold one entry per voxel of the zone in the worst case; start from what the
    // last frames needed and repeat the frame once if it turns out too small
    size_t cap_visible = std::max<size_t>(ctx->d_frame_visible.cap / 4, (size_t)n * 1024);
    const FrameTotals* h = ctx->h_small.as<FrameTotals>();
    for (int attempt = 0; attempt < 2; ++attempt) {
        VX_CUDA(ctx->d_frame_visible.ensure(cap_visible * 4));
        cap_visible = ctx->d_frame_visible.cap / 4;
        VX_CUDA(cudaMemsetAsync(totals, 0, 16, ctx->stream));  // n_faces, n_areas, n_lamps
        {
            ScopedTimer t(ctx, TM_VISIBLE);
            launch_frame_visible_zone(ctx->d_zone_xy.as<uint2>(), n, device_map(ctx), ctx->d_meshed, ctx->d_face,
                                      ctx->d_voxels, fv, ctx->d_frame_flags.as<uint8_t>(),
                                      ctx->d_zone_slot.as<int32_t>(), ctx->d_frame_counts.as<uint32_t>(), row_total,
                                      done, totals, ctx->d_frame_visible.as<uint32_t>(), (uint32_t)std::min<size_t>(cap_visible, 0xFFFFFFFFull),
                                      ctx->stream);
            ctx->timings.other_launches += 4;
        }
        VX_CUDA(cudaGetLastError());
        VX_CUDA(cudaMemcpyAsync(ctx->h_small.p, totals, sizeof(FrameTotals), cudaMemcpyDeviceToHost, ctx->stream));
        if ((st = sync_stream(ctx)) != VX_OK) return st;
        if (h->type_first[27] <= cap_visible) break;
        if (attempt == 1) return fail(ctx, VX_ERR_CUDA, "vx_visible_zone: index list still too small");
        cap_visible = h->type_first[27];
    }
    info->n_visible = h->type_first[27];
    info->n_faces = h->n_faces;
    info->n_areas = h->n_areas;
    info->n_lamps = h->n_lamps;
    std::memcpy(info->type_first, h->type_first, sizeof info->type_first);
    const uint32_t nl = std::min<uint32_t>(h->n_lamps, VX_MAX_LIGHTS);
    std::memcpy(info->lamps, h->lamps, nl * sizeof(uint32_t));
    ctx->frame_n_visible = info->n_visible;
    ctx->frame_ready = true;
    return VX_OK;
}

int vx_water_tick(vx_ctx* ctx, const uint32_t* check_xyz, int n, uint32_t* changed_xyzv, int* n_changed,
                  uint32_t* next_xyz, int* n_next) {
    if (!ctx || n < 0 || !n_changed || !n_next || (n > 0 && (!check_xyz || !changed_xyzv || !next_xyz))) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    *n_changed = *n_next = 0;
    if (n == 0) return VX_OK;
    int st = sim_check_resident(ctx, check_xyz, n, "vx_water_tick");
    if (st != VX_OK) return st;
    SimCall c{check_xyz, (size_t)n * 12, nullptr, 0, changed_xyzv, 4 * n, 4, n_changed, next_xyz, 7 * n, n_next, nullptr, 0};
    return run_sim(ctx, c, [&](uint8_t* in0, uint8_t*, uint32_t* a, uint32_t* b, uint8_t*, uint32_t* counts, int32_t* status) {
        launch_water_tick(reinterpret_cast<const uint32_t*>(in0), n, device_map(ctx), ctx->d_voxels, ctx->d_heights,
                          ctx->d_planes, a, (uint32_t)(4 * n), b, (uint32_t)(7 * n), counts, status, ctx->stream);
    });
}

int vx_falling_check(vx_ctx* ctx, const uint32_t* check_xyz, int n, uint32_t* spawned_xyzv, int cap_spawned,
                     int* n_spawned, uint32_t* water_next_xyz, int cap_next, int* n_next) {
    if (!ctx || n < 0 || cap_spawned < 0 || cap_next < 0 || !n_spawned || !n_next ||
        (n > 0 && (!check_xyz || !spawned_xyzv || !water_next_xyz)))
        return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    *n_spawned = *n_next = 0;
    if (n == 0) return VX_OK;
    int st = sim_check_resident(ctx, check_xyz, n, "vx_falling_check");
    if (st != VX_OK) return st;
    SimCall c{check_xyz, (size_t)n * 12, nullptr, 0, spawned_xyzv, cap_spawned, 4, n_spawned, water_next_xyz, cap_next, n_next,
              nullptr, 0};
    return run_sim(ctx, c, [&](uint8_t* in0, uint8_t*, uint32_t* a, uint32_t* b, uint8_t*, uint32_t* counts, int32_t* status) {
        launch_falling_check(reinterpret_cast<const uint32_t*>(in0), n, device_map(ctx), ctx->d_voxels, ctx->d_heights,
                             ctx->d_planes, a, (uint32_t)cap_spawned, b, (uint32_t)cap_next, counts, status, ctx->stream);
    });
}

int vx_falling_land(vx_ctx* ctx, const float* positions, const uint8_t* voxel_types, int n, uint8_t* keep,
                    uint32_t* placed_xyzv, int* n_placed, uint32_t* water_next_xyz, int* n_next) {
    if (!ctx || n < 0 || !n_placed || !n_next || (n > 0 && (!positions || !voxel_types || !keep || !placed_xyzv || !water_next_xyz)))
        return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    *n_placed = *n_next = 0;
    if (n == 0) return VX_OK;
    for (int i = 0; i < n; ++i) {
        const float px = positions[3 * i], py = positions[3 * i + 1], pz = positions[3 * i + 2];
        if (!(std::fabs(px) < 1.0e6f && std::fabs(py) < 1.0e6f && std::fabs(pz) < 1.0e6f))
            return fail(ctx, VX_ERR_INVALID, "vx_falling_land: position out of range");
        if (voxel_types[i] >= V_VARIANTS) return fail(ctx, VX_ERR_INVALID, "vx_falling_land: voxel id >= 27");
        const uint32_t gx = (uint32_t)((int)std::round(px) + LOCATION_OFFSET), gy = (uint32_t)((int)std::round(py) + LOCATION_OFFSET);
        if (host_slot(ctx, gx / AREA_SIZE, gy / AREA_SIZE) < 0)
            return fail(ctx, VX_ERR_NOT_LOADED, "vx_falling_land: the area under a falling voxel is not resident");
    }
    SimCall c{positions, (size_t)n * 12, voxel_types, (size_t)n, placed_xyzv, n, 4, n_placed, water_next_xyz, 7 * n, n_next, keep, n};
    return run_sim(ctx, c, [&](uint8_t* in0, uint8_t* in1, uint32_t* a, uint32_t* b, uint8_t* d_keep, uint32_t* counts, int32_t* status) {
        launch_falling_land(reinterpret_cast<const float*>(in0), in1, n, device_map(ctx), ctx->d_voxels, ctx->d_heights,
                            ctx->d_planes, d_keep, a, (uint32_t)n, b, (uint32_t)(7 * n), counts, status, ctx->stream);
    });
}

int vx_cave_columns(vx_ctx* ctx, const uint32_t* area_xy, int n, uint32_t* cave_columns_out, uint32_t* cave_voxels_out) {
    if (!ctx || n < 0 || (n > 0 && (!area_xy || !cave_columns_out))) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (!ctx->has_seed) return fail(ctx, VX_ERR_NO_SEED, "vx_cave_columns before vx_set_seed");
    VX_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return VX_OK;
    int st = sync_stream(ctx);  // d_tmp may be in use by a call that did not wait
    if (st != VX_OK) return st;
    VX_CUDA(ctx->d_tmp.ensure((size_t)n * 16));
    uint2* d_xy = ctx->d_tmp.as<uint2>();
    uint32_t* d_columns = reinterpret_cast<uint32_t*>(d_xy + n);
    uint32_t* d_voxels = d_columns + n;
    VX_CUDA(cudaMemcpyAsync(d_xy, area_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    launch_cave_columns(ctx->tables, d_xy, n, d_columns, d_voxels, ctx->stream);
    ctx->timings.other_launches += 1;
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(cave_columns_out, d_columns, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (cave_voxels_out)
        VX_CUDA(cudaMemcpyAsync(cave_voxels_out, d_voxels, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_stream(ctx);
}

int vx_diag_fp64_rate(vx_ctx* ctx, double* gops_per_s, float* ms_out) {
    if (!ctx || !gops_per_s) return VX_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    VX_CUDA(cudaSetDevice(ctx->device));
    VX_CUDA(ctx->d_tmp.ensure(64));
    const int blocks = ctx->sm_count * 8, iters = 1 << 15;
    cudaEvent_t a, b;
    VX_CUDA(cudaEventCreate(&a));
    VX_CUDA(cudaEventCreate(&b));
    launch_fp64_rate(ctx->d_tmp.as<double>(), blocks, iters / 8, ctx->stream);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, ctx->stream);
        launch_fp64_rate(ctx->d_tmp.as<double>(), blocks, iters, ctx->stream);
        cudaEventRecord(b, ctx->stream);
        VX_CUDA(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        VX_CUDA(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    const double ops = (double)blocks * 256.0 * 16.0 * (double)iters;
    *gops_per_s = ops / (best * 1e-3) / 1e9;
    if (ms_out) *ms_out = best;
    return VX_OK;
}

}  // extern "C"
