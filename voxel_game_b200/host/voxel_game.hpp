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
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
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
    static Location from_internal(InternalLocation l) {  // location.rs:19-27
        return {static_cast<int32_t>(l.x) - LOCATION_OFFSET, static_cast<int32_t>(l.y) - LOCATION_OFFSET,
                static_cast<int32_t>(l.z)};
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
    // a context created elsewhere (e.g. by the Python plumbing); not destroyed here
    explicit Pipeline(vx_ctx* borrowed) : ctx_(borrowed), owned_(false) {}
    ~Pipeline() {
        if (owned_) vx_destroy(ctx_);
    }
    Pipeline(const Pipeline&) = delete;
    Pipeline& operator=(const Pipeline&) = delete;
    vx_ctx* raw() const { return ctx_; }
    // vx_set_deferred_reads / vx_wait_reads: vx_encode_read and vx_mesh_read_compact into pinned buffers
    // return before their copies are done (one thread pipelining pre-generation steps)
    void set_deferred_reads(bool on) const { check(vx_set_deferred_reads(ctx_, on ? 1 : 0)); }
    void wait_reads() const { check(vx_wait_reads(ctx_)); }
    void check(int st) const {
        if (st == VX_OK) return;
        char msg[1024] = "";
        vx_copy_last_error(ctx_, msg, sizeof msg);  // copied out under the context's lock
        throw VxError(st, msg);
    }
    // vx_lock ... vx_unlock around a call and the read of its result (vx_mesh -> vx_mesh_read, ...):
    // the Pipeline may be shared by worker threads, the pair must see its own result
    class Transaction {
    public:
        explicit Transaction(const Pipeline& p) : ctx_(p.raw()) { vx_lock(ctx_); }
        ~Transaction() { vx_unlock(ctx_); }
        Transaction(const Transaction&) = delete;
        Transaction& operator=(const Transaction&) = delete;

    private:
        vx_ctx* ctx_;
    };

private:
    vx_ctx* ctx_ = nullptr;
    bool owned_ = true;
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
    // The device changed voxels behind the host copies (simulator ticks): read the touched areas again.
    void refresh(const std::vector<InternalLocation>& changed) {
        std::vector<AreaLocation> seen;
        for (const auto& l : changed) {
            const AreaLocation a = convert_global_to_area_location(l);
            if (std::find(seen.begin(), seen.end(), a) != seen.end()) continue;
            seen.push_back(a);
            auto it = areas_.find(a);
            if (it == areas_.end()) continue;
            const uint32_t xy[2] = {a.x, a.y};
            p_->check(vx_read_areas(p_->raw(), xy, 1, it->second.voxels_.data(), it->second.max_height_.data()));
            it->second.has_changed = true;
        }
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
    void insert(const LocKey& key, MeshInfo&& mi) {  // renderer.rs:60-67
        if (mi.voxel == Voxel::Lamp) {
            bool known = false;
            for (const LocKey& l : lights) known = known || l == key;
            if (!known) lights.push_back(key);
        } else {
            drop_light(key);
        }
        mesh_map[key] = std::move(mi);
    }
    void remove(const LocKey& key) {  // renderer.rs:69-72
        drop_light(key);
        mesh_map.erase(key);
    }

private:
    void drop_light(const LocKey& key) {
        for (size_t i = 0; i < lights.size(); ++i)
            if (lights[i] == key) {
                lights.erase(lights.begin() + static_cast<long>(i));
                return;
            }
    }
};

// ---- host-side expansion of the compact record transport (vx_mesh_read_compact) ---------------
// MeshGenerator::generate_mesh / get_verticies_for_voxel (mesh_generator.rs:109-147,200-498) on the
// host: the four vertices and six indices of every face of a visible voxel are a pure function of
// (location, voxel type, face mask).  The arithmetic is the reference's f32 arithmetic (offset -/+
// 0.5, "offset_z - 0.5 + height_offset", "side_uv.y -= height_offset"; no multiply, so nothing a
// compiler could contract) and the result is byte-identical to the device's vertex stream
// (tests/test_gpu_host.py).  Corner table: per direction four (x side, y side, bottom z, uv corner)
// tuples in the vertex order of the six vec![] blocks of mesh_generator.rs:222-487.
namespace detail {
struct Corner {
    uint8_t px, py, bottom, uv;
};
// push order of renderer.rs:139-180: Left(x+1) Right(x-1) Front(y+1) Back(y-1) Down(z+1) Up(z-1)
constexpr Corner CORNERS[6][4] = {
    {{1, 0, 1, 1}, {1, 0, 0, 2}, {1, 1, 0, 3}, {1, 1, 1, 0}},
    {{0, 0, 0, 3}, {0, 0, 1, 0}, {0, 1, 1, 1}, {0, 1, 0, 2}},
    {{0, 1, 1, 0}, {1, 1, 1, 1}, {1, 1, 0, 2}, {0, 1, 0, 3}},
    {{0, 0, 0, 3}, {1, 0, 0, 2}, {1, 0, 1, 1}, {0, 0, 1, 0}},
    {{0, 0, 1, 0}, {1, 0, 1, 1}, {1, 1, 1, 2}, {0, 1, 1, 3}},
    {{1, 0, 0, 0}, {0, 0, 0, 1}, {0, 1, 0, 2}, {1, 1, 0, 3}},
};
constexpr float NORMALS[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};  // mesh_generator.rs:45-50
// TextureManager::VOXELS_WITH_DIFFERENT_FACES (texture_manager.rs:95-102)
inline bool has_different_faces(Voxel v) {
    return v == Voxel::Grass || v == Voxel::Trampoline || v == Voxel::Wood || v == Voxel::StonePillar ||
           v == Voxel::Bomb || v == Voxel::ActiveBomb;
}
}  // namespace detail

// Appends the faces of one voxel: 4 vertices and 6 indices ([0,1,2,0,2,3] + 4 * ordinal) per set bit
// of face_mask, in bit order.  Returns the number of faces.
inline unsigned expand_voxel(uint32_t gx, uint32_t gy, uint32_t gz, Voxel voxel, uint32_t face_mask, Vertex* v_out,
                             uint16_t* i_out) {
    using namespace detail;
    const float cx = static_cast<float>(static_cast<int32_t>(gx) - LOCATION_OFFSET);  // location.rs:19-27
    const float cy = static_cast<float>(static_cast<int32_t>(gy) - LOCATION_OFFSET);
    const float cz = static_cast<float>(static_cast<int32_t>(gz));
    float ho = 0.0f;  // get_partial_height_offset (mesh_generator.rs:490-498)
    switch (voxel) {
    case Voxel::Water1: ho = 1.0f / 5.0f; break;
    case Voxel::Water2: ho = 2.0f / 5.0f; break;
    case Voxel::Water3: ho = 3.0f / 5.0f; break;
    case Voxel::Water4: ho = 4.0f / 5.0f; break;
    default: break;
    }
    const float xs[2] = {cx - 0.5f, cx + 0.5f}, ys[2] = {cy - 0.5f, cy + 0.5f};
    const float zs[2] = {cz - 0.5f + ho, cz + 0.5f};  // top, bottom
    const bool diff = has_different_faces(voxel);
    unsigned ord = 0;
    for (int d = 0; d < 6; ++d) {
        if (!((face_mask >> d) & 1u)) continue;
        // v of uv corners {0, 1} and {2, 3}: UV_REPEATING 1 / 0; TOP 0.625 / 0.375, SIDE 1 / 0.75,
        // BOTTOM 0.25 / 0 for voxels with different faces (mesh_generator.rs:61-99)
        float vhi = 1.0f, vlo = 0.0f;
        if (diff) {
            vhi = d == 5 ? 0.625f : (d == 4 ? 0.25f : 1.0f);
            vlo = d == 5 ? 0.375f : (d == 4 ? 0.0f : 0.75f);
        }
        if (d < 4 && ho > 0.0f) {  // mesh_generator.rs:214-220, side faces only
            if (vhi > 0.0f) vhi -= ho;
            if (vlo > 0.0f) vlo -= ho;
        }
        for (int k = 0; k < 4; ++k) {
            const Corner c = CORNERS[d][k];
            Vertex& v = v_out[4 * ord + static_cast<unsigned>(k)];
            v.position[0] = xs[c.px];
            v.position[1] = ys[c.py];
            v.position[2] = zs[c.bottom];
            v.uv[0] = (c.uv == 1 || c.uv == 2) ? 1.0f : 0.0f;
            v.uv[1] = c.uv < 2 ? vhi : vlo;
            v.color[0] = v.color[1] = v.color[2] = v.color[3] = 255;  // mesh_generator.rs:43
            v.normal[0] = NORMALS[d][0];
            v.normal[1] = NORMALS[d][1];
            v.normal[2] = NORMALS[d][2];
            v.normal[3] = 0.0f;
        }
        const uint16_t b = static_cast<uint16_t>(4 * ord);
        uint16_t* w = i_out + 6 * ord;
        w[0] = b; w[1] = b + 1; w[2] = b + 2; w[3] = b; w[4] = b + 2; w[5] = b + 3;  // mesh_generator.rs:44
        ++ord;
    }
    return ord;
}

inline unsigned popcount6(uint32_t m) {
    unsigned n = 0;
    for (int d = 0; d < 6; ++d) n += (m >> d) & 1u;
    return n;
}

// Runs fn(first, last) over [0, n) in `threads` contiguous pieces.
template <class Fn> void parallel_ranges(size_t n, unsigned threads, Fn fn) {
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    threads = static_cast<unsigned>(std::min<size_t>(threads, std::max<size_t>(n, 1)));
    if (threads <= 1) {
        fn(size_t(0), n);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back([=] { fn(n * t / threads, n * (t + 1) / threads); });
    for (auto& th : pool) th.join();
}

// The flat stream vx_mesh_read would have delivered, rebuilt from the compact transport: 16-byte
// records (first_face = running face count in call order), vertices, indices.  face_first (n + 1)
// is an output too.  Areas are expanded in parallel.
inline void expand_records(const uint32_t* area_xy, int n, const uint32_t* rec_first, const uint32_t* packed,
                           uint32_t* face_first, vx_voxel_rec* recs, Vertex* vertices, uint16_t* indices,
                           unsigned threads = 0) {
    std::vector<uint32_t> faces(static_cast<size_t>(n) + 1, 0);
    parallel_ranges(static_cast<size_t>(n), threads, [&](size_t a0, size_t a1) {
        for (size_t a = a0; a < a1; ++a) {
            uint32_t f = 0;
            for (uint32_t q = rec_first[a]; q < rec_first[a + 1]; ++q) f += popcount6(VX_PACKED_MASK(packed[q]));
            faces[a + 1] = f;
        }
    });
    for (int a = 0; a < n; ++a) faces[a + 1] += faces[a];
    if (face_first) std::memcpy(face_first, faces.data(), (static_cast<size_t>(n) + 1) * 4);
    parallel_ranges(static_cast<size_t>(n), threads, [&](size_t a0, size_t a1) {
        for (size_t a = a0; a < a1; ++a) {
            const uint32_t gx0 = area_xy[2 * a] * AREA_SIZE, gy0 = area_xy[2 * a + 1] * AREA_SIZE;
            uint32_t f = faces[a];
            for (uint32_t q = rec_first[a]; q < rec_first[a + 1]; ++q) {
                const uint32_t p = packed[q], local = VX_PACKED_LOCAL(p), mask = VX_PACKED_MASK(p);
                const uint32_t gx = gx0 + (local & 15u), gy = gy0 + ((local >> 4) & 15u), gz = local >> 8;
                const unsigned nf = popcount6(mask);
                if (recs) recs[q] = vx_voxel_rec{gx, gy, gz | (VX_PACKED_VOXEL(p) << 8) | (mask << 16) | (nf << 24), f};
                if (vertices) expand_voxel(gx, gy, gz, static_cast<Voxel>(VX_PACKED_VOXEL(p)), mask, vertices + size_t(f) * 4, indices + size_t(f) * 6);
                f += nf;
            }
        }
    });
}

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

enum class Transport {
    Compact,  // 4-byte records over PCIe, vertices rebuilt on the host (expand_voxel)
    Raw       // records + vertex + index streams from the device (vx_mesh_read)
};

class Renderer {  // graphics/renderer.rs:97-322,365-472
public:
    Transport transport = Transport::Compact;
    unsigned threads = 0;  // host threads for the per-voxel Mesh construction (0 = all)

    // set_voxel_shader_and_find_visible_areas + render_voxels (renderer.rs:365-452) over `meshes`:
    // the per-frame filter runs on the device over the resident mirror of the meshes (mesh
    // retention), the host only receives which voxel meshes to draw, in draw order.
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
        std::vector<AreaLocation> zone;
        zone.reserve(meshes_.size());
        for (const auto& kv : meshes_) zone.push_back(kv.first);
        const auto xy = flatten(zone);
        vx_visible_info info{};
        std::vector<uint32_t> codes;
        {
            Pipeline::Transaction pair(p);
            p.check(vx_visible_zone(p.raw(), xy.data(), static_cast<int>(zone.size()), &v, &info));
            codes.resize(info.n_visible);
            p.check(vx_visible_read(p.raw(), codes.data(), nullptr));
        }
        auto key_of = [&](uint32_t code) {
            const AreaLocation a = zone[code >> 15];
            const uint32_t local = code & 0x7FFFu;
            return LocKey{a.x * AREA_SIZE + (local & 15u), a.y * AREA_SIZE + ((local >> 4) & 15u), local >> 8};
        };
        FrameResult f;
        f.visible_areas = info.n_areas;
        f.faces_visible = info.n_faces;
        f.lamps_in_view = info.n_lamps;
        f.draw_order.reserve(codes.size());
        for (uint32_t c : codes) f.draw_order.push_back(key_of(c));
        for (uint32_t k = 0; k < info.n_lamps && k < VX_MAX_LIGHTS; ++k) f.lights.push_back(key_of(info.lamps[k]));
        return f;
    }

    // load_all_blocking (renderer.rs:468-472): skips areas already meshed (:311)
    void load_all_blocking(World& world, const std::vector<AreaLocation>& areas) {
        std::vector<AreaLocation> todo;
        for (const auto& a : areas)
            if (!meshes_.count(a)) todo.push_back(a);
        mesh_batch(world.pipeline(), todo);
    }
    // update_location (renderer.rs:239-286) for the locations world.set has just changed: the voxel
    // and its <= 6 neighbours are re-meshed on the device; meshes are inserted / removed here
    void update_locations(World& world, const std::vector<InternalLocation>& locations) {
        if (locations.empty()) return;
        const Pipeline& p = world.pipeline();
        ensure_retention(p);
        std::vector<uint32_t> xyz;
        xyz.reserve(locations.size() * 3);
        for (const auto& l : locations) {
            xyz.push_back(l.x);
            xyz.push_back(l.y);
            xyz.push_back(l.z);
        }
        vx_mesh_info info{};
        std::vector<vx_voxel_rec> recs;
        std::vector<Vertex> vertices;
        std::vector<uint16_t> indices;
        {
            Pipeline::Transaction pair(p);
            p.check(vx_update_locations(p.raw(), xyz.data(), static_cast<int>(locations.size()), &info));
            recs.resize(info.n_recs);
            vertices.resize(info.n_faces * 4);
            indices.resize(info.n_faces * 6);
            p.check(vx_update_read(p.raw(), recs.data(), vertices.data(), indices.data()));
        }
        for (const vx_voxel_rec& rec : recs) {  // set_voxel_mesh (renderer.rs:197-219)
            const uint32_t gz = rec.packed & 0xFF, nf = rec.packed >> 24;
            const LocKey key{rec.gx, rec.gy, gz};
            RenderArea& ra = meshes_[AreaLocation{rec.gx / AREA_SIZE, rec.gy / AREA_SIZE}];  // created if absent (:208-212)
            if (nf == 0) {
                ra.remove(key);
                continue;
            }
            MeshInfo mi{static_cast<uint8_t>(nf), static_cast<Voxel>((rec.packed >> 8) & 0xFF), {}};
            mi.mesh.vertices.assign(vertices.begin() + size_t(rec.first_face) * 4, vertices.begin() + size_t(rec.first_face + nf) * 4);
            mi.mesh.indices.assign(indices.begin() + size_t(rec.first_face) * 6, indices.begin() + size_t(rec.first_face + nf) * 6);
            ra.insert(key, std::move(mi));
        }
    }
    void update_location(World& world, InternalLocation location) { update_locations(world, {location}); }
    void unload_area(World& world, AreaLocation a) {  // renderer.rs:111-117
        meshes_.erase(a);
        const uint32_t xy[2] = {a.x, a.y};
        world.pipeline().check(vx_mesh_unload(world.pipeline().raw(), xy, 1));
    }
    size_t get_voxel_face_count() const {  // renderer.rs:455-461
        size_t n = 0;
        for (const auto& kv : meshes_)
            for (const auto& m : kv.second.mesh_map) n += m.second.face_count;
        return n;
    }
    size_t get_voxel_count() const {
        size_t n = 0;
        for (const auto& kv : meshes_) n += kv.second.mesh_map.size();
        return n;
    }
    const std::map<AreaLocation, RenderArea>& meshes() const { return meshes_; }
    void clear() { meshes_.clear(); }

private:
    void ensure_retention(const Pipeline& p) {
        if (!retention_on_) {
            p.check(vx_set_mesh_retention(p.raw(), 1));
            retention_on_ = true;
        }
    }
    void mesh_batch(const Pipeline& p, const std::vector<AreaLocation>& areas) {
        if (areas.empty()) return;
        ensure_retention(p);
        const auto xy = flatten(areas);
        const int n = static_cast<int>(areas.size());
        vx_mesh_info info{};
        std::unique_ptr<Pipeline::Transaction> pair(new Pipeline::Transaction(p));  // vx_mesh + its read
        p.check(vx_mesh(p.raw(), xy.data(), n, &info));
        std::vector<uint32_t> rec_first(static_cast<size_t>(n) + 1);
        std::vector<RenderArea> built(static_cast<size_t>(n));
        if (transport == Transport::Compact) {
            std::vector<uint32_t> packed(info.n_recs);
            p.check(vx_mesh_read_compact(p.raw(), rec_first.data(), packed.data()));
            p.check(vx_wait_reads(p.raw()));  // no-op unless the caller turned deferred reads on
            pair.reset();
            parallel_ranges(static_cast<size_t>(n), threads, [&](size_t a0, size_t a1) {
                for (size_t a = a0; a < a1; ++a) {
                    RenderArea& ra = built[a];
                    ra.mesh_map.reserve(rec_first[a + 1] - rec_first[a]);
                    const uint32_t gx0 = areas[a].x * AREA_SIZE, gy0 = areas[a].y * AREA_SIZE;
                    for (uint32_t q = rec_first[a]; q < rec_first[a + 1]; ++q) {
                        const uint32_t pk = packed[q], local = VX_PACKED_LOCAL(pk), mask = VX_PACKED_MASK(pk);
                        const LocKey key{gx0 + (local & 15u), gy0 + ((local >> 4) & 15u), local >> 8};
                        const unsigned nf = popcount6(mask);
                        MeshInfo mi{static_cast<uint8_t>(nf), static_cast<Voxel>(VX_PACKED_VOXEL(pk)), {}};
                        mi.mesh.vertices.resize(4 * nf);
                        mi.mesh.indices.resize(6 * nf);
                        expand_voxel(key.x, key.y, key.z, mi.voxel, mask, mi.mesh.vertices.data(), mi.mesh.indices.data());
                        ra.insert(key, std::move(mi));
                    }
                }
            });
        } else {
            std::vector<uint32_t> face_first(static_cast<size_t>(n) + 1);
            std::vector<vx_voxel_rec> recs(info.n_recs);
            std::vector<Vertex> vertices(info.n_faces * 4);
            std::vector<uint16_t> indices(info.n_faces * 6);
            p.check(vx_mesh_read(p.raw(), rec_first.data(), face_first.data(), recs.data(), vertices.data(), indices.data()));
            pair.reset();
            parallel_ranges(static_cast<size_t>(n), threads, [&](size_t a0, size_t a1) {
                for (size_t a = a0; a < a1; ++a) {
                    RenderArea& ra = built[a];
                    ra.mesh_map.reserve(rec_first[a + 1] - rec_first[a]);
                    for (uint32_t r = rec_first[a]; r < rec_first[a + 1]; ++r) {
                        const vx_voxel_rec& rec = recs[r];
                        const uint32_t gz = rec.packed & 0xFF, nf = rec.packed >> 24;
                        MeshInfo mi{static_cast<uint8_t>(nf), static_cast<Voxel>((rec.packed >> 8) & 0xFF), {}};
                        mi.mesh.vertices.assign(vertices.begin() + size_t(rec.first_face) * 4, vertices.begin() + size_t(rec.first_face + nf) * 4);
                        mi.mesh.indices.assign(indices.begin() + size_t(rec.first_face) * 6, indices.begin() + size_t(rec.first_face + nf) * 6);
                        ra.insert(LocKey{rec.gx, rec.gy, gz}, std::move(mi));
                    }
                }
            });
        }
        for (int i = 0; i < n; ++i) meshes_[areas[static_cast<size_t>(i)]] = std::move(built[static_cast<size_t>(i)]);
    }
    std::map<AreaLocation, RenderArea> meshes_;
    bool retention_on_ = false;
};

// service/physics/water_simulator.rs.  check_locations is a HashSet in the reference and its tick
// iterates it while it mutates the world, so the outcome depends on the iteration order; this mirror
// iterates in insertion order and hands that order to the device (vx_water_tick applies the list one
// location after the other, exactly as simulate_voxels' loop body does).
class WaterSimulator {
public:
    void location_updated(InternalLocation l) {  // :38-53, the reference's own conditions for z - 1 / z + 1
        insert(l);
        insert({l.x + 1, l.y, l.z});
        insert({l.x - 1, l.y, l.z});
        insert({l.x, l.y + 1, l.z});
        insert({l.x, l.y - 1, l.z});
        if (l.z - 1u < AREA_HEIGHT) insert({l.x, l.y, l.z - 1});
        if (l.z > 0) insert({l.x, l.y, l.z + 1});
    }
    // simulate_voxels (:55-77); returns the number of voxels it set
    size_t simulate_voxels(World& world, Renderer& renderer) {
        std::vector<InternalLocation> check;
        check.swap(order_);  // mem::take
        set_.clear();
        if (check.empty()) return 0;
        const Pipeline& p = world.pipeline();
        std::vector<uint32_t> xyz, changed(16 * check.size()), next(21 * check.size());
        for (const auto& l : check) {
            xyz.push_back(l.x);
            xyz.push_back(l.y);
            xyz.push_back(l.z);
        }
        int nc = 0, nn = 0;
        p.check(vx_water_tick(p.raw(), xyz.data(), static_cast<int>(check.size()), changed.data(), &nc, next.data(), &nn));
        std::vector<InternalLocation> touched;
        for (int i = 0; i < nc; ++i) touched.push_back({changed[4 * i], changed[4 * i + 1], changed[4 * i + 2]});
        world.refresh(touched);
        renderer.update_locations(world, touched);  // renderer.update_location after every world.set
        for (int i = 0; i < nn; ++i) insert({next[3 * i], next[3 * i + 1], next[3 * i + 2]});
        return static_cast<size_t>(nc);
    }
    size_t pending() const { return order_.size(); }
    // check_locations.insert: what the device reports as queued by location_updated goes in as it comes
    void insert(InternalLocation l) {
        if (l.z >= AREA_HEIGHT) return;  // the reference would index past the area (z = 127 is never edited)
        if (set_.insert(LocKey{l.x, l.y, l.z}).second) order_.push_back(l);
    }

private:
    std::vector<InternalLocation> order_;
    std::unordered_set<LocKey, LocHash> set_;
};

// service/physics/falling_voxel_simulator.rs without the entity meshes and the drawing
struct SimulatedVoxel {
    Voxel voxel_type;
    float position[3];
    float velocity;
};
class FallingVoxelSimulator {
public:
    // update_voxels (:105-150) for one location: unsupported falling voxels around it (and, recursively,
    // above them) leave the world and become entities
    void update_voxels(World& world, Renderer& renderer, WaterSimulator& water, InternalLocation location) {
        const Pipeline& p = world.pipeline();
        const uint32_t xyz[3] = {location.x, location.y, location.z};
        const int cap = 4096;
        std::vector<uint32_t> spawned(4 * cap), next(21 * cap);
        int ns = 0, nn = 0;
        p.check(vx_falling_check(p.raw(), xyz, 1, spawned.data(), cap, &ns, next.data(), 7 * cap, &nn));
        std::vector<InternalLocation> touched;
        for (int i = 0; i < ns; ++i) {
            const InternalLocation l{spawned[4 * i], spawned[4 * i + 1], spawned[4 * i + 2]};
            touched.push_back(l);
            const Location w = Location::from_internal(l);
            simulated_.push_back({static_cast<Voxel>(spawned[4 * i + 3]),
                                  {static_cast<float>(w.x), static_cast<float>(w.y), static_cast<float>(w.z)}, 0.0f});
        }
        world.refresh(touched);
        renderer.update_locations(world, touched);
        for (int i = 0; i < nn; ++i) water.insert({next[3 * i], next[3 * i + 1], next[3 * i + 2]});  // location_updated, expanded
    }
    // simulate_falling (:152-166): gravity on the host, the retain pass on the device
    void simulate_falling(World& world, Renderer& renderer, WaterSimulator& water, float delta) {
        if (simulated_.empty()) return;
        std::vector<float> pos;
        std::vector<uint8_t> types;
        for (auto& v : simulated_) {
            v.velocity += delta * 0.2f;                       // GRAVITY
            v.velocity = v.velocity < 3.0f ? v.velocity : 3.0f;  // MAX_FALL_SPEED
            v.position[2] += v.velocity;
            pos.insert(pos.end(), v.position, v.position + 3);
            types.push_back(static_cast<uint8_t>(v.voxel_type));
        }
        const Pipeline& p = world.pipeline();
        const int n = static_cast<int>(simulated_.size());
        std::vector<uint8_t> keep(simulated_.size());
        std::vector<uint32_t> placed(4 * simulated_.size()), next(21 * simulated_.size());
        int np = 0, nn = 0;
        p.check(vx_falling_land(p.raw(), pos.data(), types.data(), n, keep.data(), placed.data(), &np, next.data(), &nn));
        std::vector<InternalLocation> touched;
        for (int i = 0; i < np; ++i) touched.push_back({placed[4 * i], placed[4 * i + 1], placed[4 * i + 2]});
        world.refresh(touched);
        renderer.update_locations(world, touched);
        for (int i = 0; i < nn; ++i) water.insert({next[3 * i], next[3 * i + 1], next[3 * i + 2]});
        std::vector<SimulatedVoxel> kept;
        for (size_t i = 0; i < simulated_.size(); ++i)
            if (keep[i]) kept.push_back(simulated_[i]);
        simulated_.swap(kept);
    }
    const std::vector<SimulatedVoxel>& simulated_voxels() const { return simulated_; }

private:
    std::vector<SimulatedVoxel> simulated_;
};

}  // namespace voxel_game
