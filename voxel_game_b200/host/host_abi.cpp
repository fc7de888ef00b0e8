// host_abi.cpp -- C entry points of the C++ host mirror (voxel_game.hpp), so that the tests and
// bench.py can drive the SAME host-side code a compiled-code host would use: the compact-transport
// expander and Renderer::load_all_blocking / update_locations with their per-voxel Mesh maps
// (renderer.rs:45-73,197-219,468-472).  Built into voxel_game_b200/libvoxel_host.so by
// voxel_game_b200/host/Makefile; links against libvoxel_cuda.so.  No CPU compute path: every mesh
// comes from the device through the C ABI.
#include <chrono>
#include <memory>

#include "voxel_game.hpp"

using namespace voxel_game;

namespace {
struct HostRenderer {
    std::shared_ptr<Pipeline> pipeline;
    World world;
    Renderer renderer;
    std::string error;
    explicit HostRenderer(vx_ctx* ctx) : pipeline(std::make_shared<Pipeline>(ctx)), world(pipeline) {}
};
double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
std::vector<AreaLocation> areas_of(const uint32_t* area_xy, int n) {
    std::vector<AreaLocation> out(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) out[static_cast<size_t>(i)] = AreaLocation{area_xy[2 * i], area_xy[2 * i + 1]};
    return out;
}
}  // namespace

extern "C" {

// expand_records: the flat stream of vx_mesh_read rebuilt from vx_mesh_read_compact's output
void vxh_expand_records(const uint32_t* area_xy, int n, const uint32_t* rec_first, const uint32_t* packed,
                        uint32_t* face_first, vx_voxel_rec* recs, void* vertices, uint16_t* indices, int threads) {
    expand_records(area_xy, n, rec_first, packed, face_first, recs, static_cast<Vertex*>(vertices), indices,
                   static_cast<unsigned>(threads < 0 ? 0 : threads));
}

void* vxh_renderer_create(vx_ctx* ctx) { return ctx ? new HostRenderer(ctx) : nullptr; }
void vxh_renderer_destroy(void* r) { delete static_cast<HostRenderer*>(r); }
const char* vxh_last_error(void* r) { return static_cast<HostRenderer*>(r)->error.c_str(); }

// Renderer::load_all_blocking over n areas; transport 0 = raw streams, 1 = compact records expanded
// on the host.  ms_out[0] = whole call, [1] = device + transfer part (vx_mesh + read), as wall clock.
int vxh_renderer_load_all_blocking(void* r, const uint32_t* area_xy, int n, int transport, int threads,
                                   double* ms_out) {
    auto* h = static_cast<HostRenderer*>(r);
    try {
        h->renderer.transport = transport ? Transport::Compact : Transport::Raw;
        h->renderer.threads = static_cast<unsigned>(threads < 0 ? 0 : threads);
        const double t0 = now_ms();
        h->renderer.load_all_blocking(h->world, areas_of(area_xy, n));
        if (ms_out) ms_out[0] = now_ms() - t0;
        return VX_OK;
    } catch (const VxError& e) {
        h->error = e.what();
        return e.status;
    }
}

// Renderer::update_location for n internal locations (after vx_apply_edits)
int vxh_renderer_update_locations(void* r, const uint32_t* xyz, int n) {
    auto* h = static_cast<HostRenderer*>(r);
    try {
        std::vector<InternalLocation> locs(static_cast<size_t>(n));
        for (int i = 0; i < n; ++i) locs[static_cast<size_t>(i)] = InternalLocation{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        h->renderer.update_locations(h->world, locs);
        return VX_OK;
    } catch (const VxError& e) {
        h->error = e.what();
        return e.status;
    }
}

int vxh_renderer_unload_area(void* r, uint32_t ax, uint32_t ay) {
    auto* h = static_cast<HostRenderer*>(r);
    try {
        h->renderer.unload_area(h->world, AreaLocation{ax, ay});
        return VX_OK;
    } catch (const VxError& e) {
        h->error = e.what();
        return e.status;
    }
}

void vxh_renderer_clear(void* r) { static_cast<HostRenderer*>(r)->renderer.clear(); }
uint64_t vxh_renderer_face_count(void* r) { return static_cast<HostRenderer*>(r)->renderer.get_voxel_face_count(); }
uint64_t vxh_renderer_voxel_count(void* r) { return static_cast<HostRenderer*>(r)->renderer.get_voxel_count(); }
uint64_t vxh_renderer_area_count(void* r) { return static_cast<HostRenderer*>(r)->renderer.meshes().size(); }

// The Renderer.meshes entries of one area as a dense volume (byte x + 16 y + 256 z = face count of
// that voxel's mesh, 0 = no entry) plus its lamp count: what the parity tests compare.
// Returns -1 if the area has no RenderArea.
int vxh_renderer_area_face_counts(void* r, uint32_t ax, uint32_t ay, uint8_t* counts_out, uint32_t* lamps_out) {
    auto* h = static_cast<HostRenderer*>(r);
    const auto it = h->renderer.meshes().find(AreaLocation{ax, ay});
    if (it == h->renderer.meshes().end()) return -1;
    std::memset(counts_out, 0, VOXELS_IN_AREA);
    for (const auto& m : it->second.mesh_map) {
        const LocKey& k = m.first;
        if (m.second.mesh.vertices.size() != 4u * m.second.face_count || m.second.mesh.indices.size() != 6u * m.second.face_count)
            return -2;
        counts_out[(k.x % AREA_SIZE) + (k.y % AREA_SIZE) * AREA_SIZE + k.z * 256u] = m.second.face_count;
    }
    if (lamps_out) *lamps_out = static_cast<uint32_t>(it->second.lights.size());
    return VX_OK;
}

// One voxel's mesh: vertices (face_count * 160 bytes) and indices (face_count * 6) copied out.
// Returns the face count, 0 if the voxel has no mesh.
int vxh_renderer_voxel_mesh(void* r, uint32_t gx, uint32_t gy, uint32_t gz, void* vertices_out, uint16_t* indices_out) {
    auto* h = static_cast<HostRenderer*>(r);
    const auto it = h->renderer.meshes().find(AreaLocation{gx / AREA_SIZE, gy / AREA_SIZE});
    if (it == h->renderer.meshes().end()) return 0;
    const auto m = it->second.mesh_map.find(LocKey{gx, gy, gz});
    if (m == it->second.mesh_map.end()) return 0;
    std::memcpy(vertices_out, m->second.mesh.vertices.data(), m->second.mesh.vertices.size() * sizeof(Vertex));
    std::memcpy(indices_out, m->second.mesh.indices.data(), m->second.mesh.indices.size() * 2);
    return m->second.face_count;
}

}  // extern "C"
