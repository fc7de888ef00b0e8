// voxel_game.hpp -- C++ mirror of the reference's host-side interface for the chunk pipeline, on
// top of the C ABI (include/voxel_cuda.h).  The reference is Rust; the build image has no Rust
// toolchain, so this header plays the role of the Rust shim (rust/voxel_cuda) for compiled-code
// hosts and for tests/cpp/test_host.cpp, which restates the reference's own unit tests against it.
// Same names, argument meaning and error behaviour as the reference:
//   AreaLocation / InternalLocation / Location      model/location.rs:8-96
//   Voxel                                           model/voxel.rs:8-36
//   Area                                            model/area.rs:11-201
//   AreaGenerator::generate_area                    service/area_generation/generator.rs:49-64
//   AreaLoader::load_all_blocking                   service/persistence/world_persistence.rs:85-94
//   World::{load_all_blocking, get, set, get_height} model/world.rs:32-263
//   Renderer::{load_all_blocking, update_locations, get_voxel_face_count}
//                                                   graphics/renderer.rs:239-322,455-472
// No CPU compute path: every method that produces voxels, heights or meshes calls the library and
// throws VxError when it fails.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/voxel_cuda.h"

namespace voxel_game {

constexpr uint32_t AREA_SIZE = 16;     // model/area.rs:7
constexpr uint32_t AREA_HEIGHT = 128;  // model/area.rs:8
constexpr size_t VOXELS_IN_AREA = 32768;
constexpr int32_t LOCATION_OFFSET = 1000000;  // model/location.rs:6

enum class Voxel : uint8_t {  // model/voxel.rs:8-36
    None, Cobblestone, Sand, Grass, Wood, Leaves, Brick, Dirt, Boards, Stone, Clay, Lamp, Trampoline,
    Cactus, WaterSource, WaterDown, Water1, Water2, Water3, Water4, StoneBrick, StonePillar, Snow, Ice,
    Bomb, ActiveBomb, Glass
};

inline bool is_transparent(Voxel v) {  // Voxel::TRANSPARENT, model/voxel.rs:39-49
    return (0x048FC001u >> static_cast<unsigned>(v)) & 1u;
}

struct AreaLocation {
    uint32_t x, y;
    bool operator<(const AreaLocation& o) const { return x != o.x ? x < o.x : y < o.y; }
    bool operator==(const AreaLocation& o) const { return x == o.x && y == o.y; }
};
struct InternalLocation {
    uint32_t x, y, z;
};
struct Location {  // player-facing coordinates, model/location.rs:8-27
    int32_t x, y, z;
    InternalLocation to_internal() const {
        return {static_cast<uint32_t>(x + LOCATION_OFFSET), static_cast<uint32_t>(y + LOCATION_OFFSET),
                static_cast<uint32_t>(z)};
    }
};

struct VxError : std::runtime_error {
    int status;
    VxError(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

// One CUDA device + one world seed.  hash_world_name is computed by the library here (a Rust host
// passes std's DefaultHasher value instead, generator.rs:24-28).
class Pipeline {
public:
    Pipeline(int device, const std::string& world_name) {
        int st = vx_create(device, &ctx_);
        if (st != VX_OK) throw VxError(st, vx_strerror(st));
        check(vx_set_seed(ctx_, vx_hash_world_name(world_name.c_str())));
    }
    ~Pipeline() { vx_destroy(ctx_); }
    Pipeline(const Pipeline&) = delete;
    Pipeline& operator=(const Pipeline&) = delete;
    vx_ctx* raw() const { return ctx_; }
    void check(int st) const {
        if (st != VX_OK) throw VxError(st, vx_last_error(ctx_));
    }

private:
    vx_ctx* ctx_ = nullptr;
};

class Area {  // model/area.rs:11-124
public:
    bool has_changed = true;
    Area(AreaLocation loc, std::vector<uint8_t> voxels, std::vector<uint8_t> max_height)
        : area_location_(loc), voxels_(std::move(voxels)), max_height_(std::move(max_height)) {}
    Voxel get(InternalLocation l) const { return static_cast<Voxel>(voxels_[l.x + l.y * AREA_SIZE + l.z * AREA_SIZE * AREA_SIZE]); }
    uint8_t sample_height(uint32_t x, uint32_t y) const { return max_height_[x + AREA_SIZE * y]; }
    AreaLocation get_area_location() const { return area_location_; }
    uint32_t get_x() const { return area_location_.x; }
    uint32_t get_y() const { return area_location_.y; }
    const std::vector<uint8_t>& voxels() const { return voxels_; }
    const std::vector<uint8_t>& max_height() const { return max_height_; }
    // Area::update_all_column_heights (area.rs:34-44): recomputed on the device
    void update_all_column_heights(const Pipeline& p) {
        p.check(vx_column_heights(p.raw(), voxels_.data(), 1, max_height_.data()));
    }

private:
    friend class World;
    AreaLocation area_location_;
    std::vector<uint8_t> voxels_, max_height_;
};

inline std::vector<uint32_t> flatten(const std::vector<AreaLocation>& areas) {
    std::vector<uint32_t> xy;
    xy.reserve(areas.size() * 2);
    for (const auto& a : areas) {
        xy.push_back(a.x);
        xy.push_back(a.y);
    }
    return xy;
}

struct AreaGenerator {
    // generator.rs:49-64; a single area is a batch of one
    static Area generate_area(const Pipeline& p, AreaLocation loc) {
        std::vector<uint8_t> v(VOXELS_IN_AREA), h(256);
        const uint32_t xy[2] = {loc.x, loc.y};
        p.check(vx_generate(p.raw(), xy, 1, v.data(), h.data()));
        return Area(loc, std::move(v), std::move(h));
    }
};

struct AreaLoader {
    // world_persistence.rs:85-94 with every area a disk miss: ONE launch for the batch
    static std::vector<Area> load_all_blocking(const Pipeline& p, const std::vector<AreaLocation>& areas) {
        const size_t n = areas.size();
        std::vector<uint8_t> v(n * VOXELS_IN_AREA), h(n * 256);
        const auto xy = flatten(areas);
        p.check(vx_generate(p.raw(), xy.data(), static_cast<int>(n), v.data(), h.data()));
        std::vector<Area> out;
        out.reserve(n);
        for (size_t i = 0; i < n; ++i)
            out.emplace_back(areas[i], std::vector<uint8_t>(v.begin() + i * VOXELS_IN_AREA, v.begin() + (i + 1) * VOXELS_IN_AREA),
                             std::vector<uint8_t>(h.begin() + i * 256, h.begin() + (i + 1) * 256));
        return out;
    }
};

class World {  // model/world.rs:16-276 (area cache + get/set)
public:
    explicit World(std::shared_ptr<Pipeline> p) : p_(std::move(p)) {}
    const Pipeline& pipeline() const { return *p_; }
    void load_all_blocking(const std::vector<AreaLocation>& areas) {  // world.rs:247-263
        std::vector<AreaLocation> todo;
        for (const auto& a : areas)
            if (!areas_.count(a)) todo.push_back(a);
        for (auto& area : AreaLoader::load_all_blocking(*p_, todo)) areas_.emplace(area.get_area_location(), std::move(area));
    }
    void load_area(AreaLocation a) {  // world.rs:32-38
        if (!areas_.count(a)) areas_.emplace(a, AreaGenerator::generate_area(*p_, a));
    }
    static AreaLocation convert_global_to_area_location(InternalLocation l) { return {l.x / AREA_SIZE, l.y / AREA_SIZE}; }
    Voxel get(InternalLocation l) {  // world.rs:149-155
        const AreaLocation a = convert_global_to_area_location(l);
        load_area(a);
        return areas_.at(a).get({l.x % AREA_SIZE, l.y % AREA_SIZE, l.z});
    }
    uint8_t get_height(InternalLocation l) {  // world.rs:71-76
        const AreaLocation a = convert_global_to_area_location(l);
        load_area(a);
        return areas_.at(a).sample_height(l.x % AREA_SIZE, l.y % AREA_SIZE);
    }
    // World::set for a batch (world.rs:184-191), array order; returns the areas Renderer must re-mesh.
    // The host copies of the touched areas are refreshed from the device.
    std::vector<AreaLocation> set_batch(const std::vector<std::pair<InternalLocation, Voxel>>& edits) {
        std::vector<vx_edit> e;
        e.reserve(edits.size());
        for (const auto& ed : edits) {
            load_area(convert_global_to_area_location(ed.first));
            e.push_back({ed.first.x, ed.first.y, ed.first.z, static_cast<uint32_t>(ed.second)});
        }
        std::vector<uint32_t> dirty(edits.size() * 6 + 2);
        int nd = 0;
        p_->check(vx_apply_edits(p_->raw(), e.data(), static_cast<int>(e.size()), dirty.data(), &nd));
        std::vector<AreaLocation> out;
        for (int i = 0; i < nd; ++i) out.push_back({dirty[2 * i], dirty[2 * i + 1]});
        for (const auto& a : out) {
            auto it = areas_.find(a);
            if (it == areas_.end()) continue;
            const uint32_t xy[2] = {a.x, a.y};
            p_->check(vx_read_areas(p_->raw(), xy, 1, it->second.voxels_.data(), it->second.max_height_.data()));
            it->second.has_changed = true;
        }
        return out;
    }
    size_t get_loaded_areas_count() const { return areas_.size(); }
    const Area& get_area_without_loading(AreaLocation a) const { return areas_.at(a); }

private:
    std::shared_ptr<Pipeline> p_;
    std::map<AreaLocation, Area> areas_;
};

struct Vertex {  // macroquad::ui::Vertex, repr(C)
    float position[3];
    float uv[2];
    uint8_t color[4];
    float normal[4];
};
static_assert(sizeof(Vertex) == VX_VERTEX_BYTES, "Vertex layout");

struct Mesh {  // macroquad::models::Mesh without the texture handle
    std::vector<Vertex> vertices;
    std::vector<uint16_t> indices;
};
struct MeshInfo {  // renderer.rs:44-45
    uint8_t face_count;
    Voxel voxel;
    Mesh mesh;
};
struct LocKey {
    uint32_t x, y, z;
    bool operator==(const LocKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct LocHash {
    size_t operator()(const LocKey& k) const { return (static_cast<size_t>(k.x) * 73856093u) ^ (static_cast<size_t>(k.y) * 19349663u) ^ (k.z * 83492791u); }
};
struct RenderArea {  // renderer.rs:48-73
    std::unordered_map<LocKey, MeshInfo, LocHash> mesh_map;
    std::vector<LocKey> lights;  // Lamp voxels with at least one face
};

struct Camera3D {  // macroquad::camera::Camera3D, the two fields the visibility pass reads
    float position[3];
    float target[3];
};
struct FrameResult {  // what one frame draws
    size_t visible_areas = 0, faces_visible = 0;  // render_voxels' return value (renderer.rs:414-452)
    std::vector<LocKey> draw_order;                // voxels in draw_mesh order (optimise_render_order)
    std::vector<LocKey> lights;                    // set_lights input: first <= 64 lamps (voxel_shader.rs:176-200)
    size_t lamps_in_view = 0;
};

class Renderer {  // graphics/renderer.rs:97-322,365-472
public:
    // set_voxel_shader_and_find_visible_areas + render_voxels (renderer.rs:365-452) over the areas of
    // the last load_all_blocking / update_locations batch: the per-frame filter runs on the device
    // over the records kept there, the host only receives the indices of the meshes to draw.
    FrameResult render_voxels(World& world, const Camera3D& camera, uint32_t render_size, bool head_in_water,
                              bool show_map) {
        const Pipeline& p = world.pipeline();
        vx_view v{};
        float d[3], len2 = 0.f;
        for (int k = 0; k < 3; ++k) {
            v.cam_pos[k] = camera.position[k];
            d[k] = camera.target[k] - camera.position[k];
        }
        len2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
        const float rcp = 1.0f / std::sqrt(len2);  // Vec3::normalize_or_zero
        const bool ok = std::isfinite(rcp) && rcp > 0.0f;
        for (int k = 0; k < 3; ++k) v.look[k] = ok ? d[k] * rcp : 0.0f;
        v.cam_target_z = camera.target[2];
        v.render_size = render_size;
        v.head_in_water = head_in_water;
        v.show_map = show_map;
        vx_visible_info info{};
        p.check(vx_visible(p.raw(), &v, &info));
        std::vector<uint32_t> idx(info.n_visible);
        p.check(vx_visible_read(p.raw(), idx.data(), nullptr));
        FrameResult f;
        f.visible_areas = info.n_areas;
        f.faces_visible = info.n_faces;
        f.lamps_in_view = info.n_lamps;
        f.draw_order.reserve(idx.size());
        for (uint32_t r : idx) f.draw_order.push_back(last_keys_[r]);
        for (uint32_t k = 0; k < info.n_lamps && k < VX_MAX_LIGHTS; ++k) f.lights.push_back(last_keys_[info.lamps[k]]);
        return f;
    }

    // load_all_blocking (renderer.rs:468-472): skips areas already meshed (:311)
    void load_all_blocking(World& world, const std::vector<AreaLocation>& areas) {
        std::vector<AreaLocation> todo;
        for (const auto& a : areas)
            if (!meshes_.count(a)) todo.push_back(a);
        mesh_batch(world.pipeline(), todo);
    }
    // update_location for a batch of edits (renderer.rs:239-286): re-mesh the dirty areas
    void update_locations(World& world, const std::vector<AreaLocation>& dirty) {
        std::vector<AreaLocation> loaded;
        for (const auto& a : dirty)
            if (meshes_.count(a)) loaded.push_back(a);
        mesh_batch(world.pipeline(), loaded);
    }
    void unload_area(AreaLocation a) { meshes_.erase(a); }
    size_t get_voxel_face_count() const {  // renderer.rs:455-461
        size_t n = 0;
        for (const auto& kv : meshes_)
            for (const auto& m : kv.second.mesh_map) n += m.second.face_count;
        return n;
    }
    const std::map<AreaLocation, RenderArea>& meshes() const { return meshes_; }

private:
    void mesh_batch(const Pipeline& p, const std::vector<AreaLocation>& areas) {
        if (areas.empty()) return;
        const auto xy = flatten(areas);
        const int n = static_cast<int>(areas.size());
        vx_mesh_info info{};
        p.check(vx_mesh(p.raw(), xy.data(), n, &info));
        std::vector<uint32_t> rec_first(n + 1), face_first(n + 1);
        std::vector<vx_voxel_rec> recs(info.n_recs);
        std::vector<Vertex> vertices(info.n_faces * 4);
        std::vector<uint16_t> indices(info.n_faces * 6);
        p.check(vx_mesh_read(p.raw(), rec_first.data(), face_first.data(), recs.data(), vertices.data(), indices.data()));
        last_keys_.assign(info.n_recs, LocKey{0, 0, 0});
        for (int i = 0; i < n; ++i) {
            RenderArea ra;
            for (uint32_t r = rec_first[i]; r < rec_first[i + 1]; ++r) {
                const vx_voxel_rec& rec = recs[r];
                const uint32_t gz = rec.packed & 0xFF, nf = rec.packed >> 24;
                const Voxel voxel = static_cast<Voxel>((rec.packed >> 8) & 0xFF);
                MeshInfo mi{static_cast<uint8_t>(nf), voxel, {}};
                mi.mesh.vertices.assign(vertices.begin() + size_t(rec.first_face) * 4, vertices.begin() + size_t(rec.first_face + nf) * 4);
                mi.mesh.indices.assign(indices.begin() + size_t(rec.first_face) * 6, indices.begin() + size_t(rec.first_face + nf) * 6);
                const LocKey key{rec.gx, rec.gy, gz};
                last_keys_[r] = key;
                if (voxel == Voxel::Lamp) ra.lights.push_back(key);  // RenderArea::insert, renderer.rs:60-67
                ra.mesh_map.emplace(key, std::move(mi));
            }
            meshes_[areas[i]] = std::move(ra);
        }
    }
    std::map<AreaLocation, RenderArea> meshes_;
    std::vector<LocKey> last_keys_;  // record index of the last batch -> voxel location
};

}  // namespace voxel_game
