// test_host.cpp -- the reference's own unit tests for the hot path, restated against the C++ host
// mirror (voxel_game_b200/host/voxel_game.hpp) and therefore against the CUDA path through the C
// ABI.  Needs a GPU; built and run by tests/test_gpu_parity.py::test_cpp_host_mirror.
//   generator.rs:155-220   test_generate_area, _different_locations, _different_seeds, _same_area,
//                          _heights_calculated_correctly
//   world.rs:288-336       test_get_same_location, test_get_renderable_locations_for_area (structural)
//   renderer                face count bookkeeping, edits through World::set + update_location,
//                          one frame of render_voxels (visible areas / voxels, draw order)
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "../../voxel_game_b200/host/voxel_game.hpp"

using namespace voxel_game;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                    \
        }                                                                    \
    } while (0)

static bool areas_differ(const Area& a, const Area& b) {  // generator.rs:222-243
    return a.voxels() != b.voxels() || a.max_height() != b.max_height();
}

int main() {
    auto p = std::make_shared<Pipeline>(0, "test");

    // test_generate_area (generator.rs:155-179)
    Area area = AreaGenerator::generate_area(*p, {123, 456});
    CHECK(area.has_changed && area.get_x() == 123 && area.get_y() == 456);
    for (uint32_t x = 0; x < AREA_SIZE; ++x)
        for (uint32_t y = 0; y < AREA_SIZE; ++y) {
            int max_height = -1;
            bool contains_stone = false;
            for (uint32_t z = 0; z < AREA_HEIGHT; ++z) {
                const Voxel v = area.get({x, y, z});
                if (!is_transparent(v) && max_height < 0) max_height = static_cast<int>(z);
                if (v == Voxel::Stone) contains_stone = true;
            }
            CHECK(max_height >= 0);
            CHECK(contains_stone);
            CHECK(max_height == area.sample_height(x, y));
        }

    // test_generate_area_different_locations / _different_seeds / _same_area (generator.rs:181-200)
    CHECK(areas_differ(area, AreaGenerator::generate_area(*p, {999, 400})));
    {
        Pipeline p1(0, "test1"), p2(0, "test2");
        CHECK(areas_differ(AreaGenerator::generate_area(p1, {123, 456}), AreaGenerator::generate_area(p2, {123, 456})));
    }
    CHECK(!areas_differ(area, AreaGenerator::generate_area(*p, {123, 456})));

    // test_genearate_area_heights_calculated_correctly (generator.rs:202-220)
    for (uint32_t x = 0; x < 10; ++x) {
        Area a = AreaGenerator::generate_area(*p, {x, 123});
        Area b = a;
        b.update_all_column_heights(*p);
        CHECK(!areas_differ(a, b));
    }

    // World: get is stable (world.rs:302-316), batch load == single loads
    World world(p);
    std::vector<AreaLocation> zone;
    for (uint32_t x = 62498; x < 62503; ++x)
        for (uint32_t y = 62498; y < 62503; ++y) zone.push_back({x, y});
    world.load_all_blocking(zone);
    CHECK(world.get_loaded_areas_count() == 25);
    CHECK(!areas_differ(world.get_area_without_loading({62500, 62500}), AreaGenerator::generate_area(*p, {62500, 62500})));
    const InternalLocation probe = Location{3, 5, 100}.to_internal();
    CHECK(world.get(probe) == world.get(probe));

    // Renderer: every mesh entry is a non-None voxel of its area with 1..6 faces, 4 vertices and 6
    // indices per face (renderer.rs:126-219); face count bookkeeping (:455-461)
    Renderer renderer;
    std::vector<AreaLocation> inner;
    for (uint32_t x = 62499; x < 62502; ++x)
        for (uint32_t y = 62499; y < 62502; ++y) inner.push_back({x, y});
    renderer.load_all_blocking(world, inner);
    size_t faces = 0;
    for (const auto& kv : renderer.meshes())
        for (const auto& m : kv.second.mesh_map) {
            const MeshInfo& mi = m.second;
            CHECK(mi.voxel != Voxel::None && mi.face_count >= 1 && mi.face_count <= 6);
            CHECK(mi.mesh.vertices.size() == 4u * mi.face_count && mi.mesh.indices.size() == 6u * mi.face_count);
            CHECK(world.get({m.first.x, m.first.y, m.first.z}) == mi.voxel);
            CHECK(mi.mesh.indices[0] == 0 && mi.mesh.indices[5] == 3);
            faces += mi.face_count;
        }
    CHECK(faces > 0 && faces == renderer.get_voxel_face_count());
    const size_t before = renderer.get_voxel_face_count();
    renderer.load_all_blocking(world, inner);  // already meshed: no-op (renderer.rs:311)
    CHECK(renderer.get_voxel_face_count() == before);

    // edit: place a Lamp in the air near the spawn point, then break it again
    const InternalLocation lamp_at = Location{5, 6, 20}.to_internal();  // interior column of area (62500, 62500)
    CHECK(world.get(lamp_at) == Voxel::None);
    auto dirty = world.set_batch({{lamp_at, Voxel::Lamp}});
    CHECK(dirty.size() == 1 && dirty[0] == (AreaLocation{62500, 62500}));
    CHECK(world.get(lamp_at) == Voxel::Lamp);
    CHECK(world.get_height(lamp_at) == 20);
    renderer.update_location(world, lamp_at);
    CHECK(renderer.get_voxel_face_count() == before + 6);  // a lone block has six faces
    CHECK(renderer.meshes().at({62500, 62500}).lights.size() == 1);
    dirty = world.set_batch({{lamp_at, Voxel::None}});
    renderer.update_location(world, lamp_at);
    CHECK(renderer.get_voxel_face_count() == before);
    CHECK(renderer.meshes().at({62500, 62500}).lights.empty());
    CHECK(world.get_height(lamp_at) > 20);

    // water (water_simulator.rs): in a walled 3 x 3 basin on a stone plate in the air a source fills its
    // four sides with Water1 and the corners with Water2, and everything dries up again once the source
    // is gone (on open ground a stream would run downhill for ever: the front advances as the tail
    // dries); the meshes end as they began.  Falling sand (falling_voxel_simulator.rs): placed in the air
    // it becomes an entity and comes to rest on the ground.
    {
        WaterSimulator water;
        FallingVoxelSimulator falling;
        const Location col{5, 6, 0};
        const uint32_t ground = world.get_height(col.to_internal());  // first opaque z of the column
        const int32_t plate_z = static_cast<int32_t>(ground) - 12, basin_z = plate_z - 1;
        CHECK(plate_z > 2);
        std::vector<std::pair<InternalLocation, Voxel>> build, demolish;
        std::vector<InternalLocation> built;
        for (int dx = -2; dx <= 2; ++dx)
            for (int dy = -2; dy <= 2; ++dy) {
                const InternalLocation floor = Location{5 + dx, 6 + dy, plate_z}.to_internal();
                CHECK(world.get(floor) == Voxel::None);
                build.push_back({floor, Voxel::Stone});
                if (dx == -2 || dx == 2 || dy == -2 || dy == 2) {
                    const InternalLocation wall = Location{5 + dx, 6 + dy, basin_z}.to_internal();
                    CHECK(world.get(wall) == Voxel::None);
                    build.push_back({wall, Voxel::Stone});
                }
            }
        for (const auto& e : build) {
            built.push_back(e.first);
            demolish.push_back({e.first, Voxel::None});
        }
        world.set_batch(build);
        renderer.update_locations(world, built);
        const size_t with_basin = renderer.get_voxel_face_count();
        CHECK(with_basin > before);
        const InternalLocation src = Location{5, 6, basin_z}.to_internal();
        world.set_batch({{src, Voxel::WaterSource}});
        renderer.update_location(world, src);
        water.location_updated(src);
        size_t changed_total = 0;
        for (int tick = 0; tick < 8; ++tick) changed_total += water.simulate_voxels(world, renderer);
        CHECK(changed_total >= 8);
        CHECK(world.get({src.x + 1, src.y, src.z}) == Voxel::Water1 && world.get({src.x, src.y - 1, src.z}) == Voxel::Water1);
        CHECK(world.get({src.x + 1, src.y + 1, src.z}) == Voxel::Water2 && world.get({src.x - 1, src.y - 1, src.z}) == Voxel::Water2);
        CHECK(world.get({src.x + 2, src.y, src.z}) == Voxel::Stone);  // the wall holds
        CHECK(renderer.get_voxel_face_count() > with_basin);
        world.set_batch({{src, Voxel::None}});
        renderer.update_location(world, src);
        water.location_updated(src);
        for (int tick = 0; tick < 40 && water.pending(); ++tick) water.simulate_voxels(world, renderer);
        CHECK(water.pending() == 0);
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) CHECK(world.get({src.x + dx, src.y + dy, src.z}) == Voxel::None);
        CHECK(renderer.get_voxel_face_count() == with_basin);
        world.set_batch(demolish);
        renderer.update_locations(world, built);
        CHECK(renderer.get_voxel_face_count() == before);

        const InternalLocation sand = Location{5, 6, static_cast<int32_t>(ground) - 6}.to_internal();
        world.set_batch({{sand, Voxel::Sand}});
        renderer.update_location(world, sand);
        CHECK(renderer.get_voxel_face_count() == before + 6);
        falling.update_voxels(world, renderer, water, sand);
        CHECK(falling.simulated_voxels().size() == 1 && world.get(sand) == Voxel::None);
        CHECK(renderer.get_voxel_face_count() == before);
        for (int step = 0; step < 200 && !falling.simulated_voxels().empty(); ++step)
            falling.simulate_falling(world, renderer, water, 0.5f);
        CHECK(falling.simulated_voxels().empty());
        CHECK(world.get({sand.x, sand.y, ground - 1}) == Voxel::Sand);
        CHECK(world.get_height(sand) == ground - 1);
        world.set_batch({{InternalLocation{sand.x, sand.y, ground - 1}, Voxel::None}});
        renderer.update_location(world, {sand.x, sand.y, ground - 1});
        CHECK(renderer.get_voxel_face_count() == before);
    }

    // one frame (renderer.rs:365-452): everything in the map view, a strict subset when looking +x
    // from the spawn column, nothing drawn twice, every drawn voxel has a mesh
    // (the frame runs over the resident meshes of the whole zone; nothing is re-meshed for it)
    const float eye = static_cast<float>(world.get_height(Location{8, 8, 0}.to_internal())) - 3.f;  // above the ground
    const FrameResult all = renderer.render_voxels(world, Camera3D{{8.f, 8.f, eye}, {9.f, 8.f, eye}}, 8, false, true);
    CHECK(all.visible_areas == inner.size() && all.faces_visible == renderer.get_voxel_face_count());
    const FrameResult some = renderer.render_voxels(world, Camera3D{{8.f, 8.f, eye}, {9.f, 8.f, eye}}, 8, false, false);
    CHECK(some.visible_areas > 0 && some.visible_areas < all.visible_areas);
    CHECK(some.faces_visible > 0 && some.faces_visible < all.faces_visible);
    size_t some_faces = 0;
    int last_type = -1;
    for (const LocKey& k : some.draw_order) {
        const AreaLocation a{k.x / AREA_SIZE, k.y / AREA_SIZE};
        const MeshInfo& mi = renderer.meshes().at(a).mesh_map.at(k);
        some_faces += mi.face_count;
        CHECK(static_cast<int>(mi.voxel) >= last_type);  // grouped by voxel type
        last_type = static_cast<int>(mi.voxel);
        CHECK(static_cast<float>(static_cast<int>(k.x) - LOCATION_OFFSET) > 8.f - 6.f);  // nothing far behind the camera
    }
    CHECK(some_faces == some.faces_visible);

    // error behaviour: an unknown device has no CPU fallback
    bool threw = false;
    try {
        Pipeline bad(1000, "test");
    } catch (const VxError& e) {
        threw = e.status == VX_ERR_INVALID || e.status == VX_ERR_NO_DEVICE;
    }
    CHECK(threw);

    // both transports build the same Renderer.meshes: vertices rebuilt on the host from the 4-byte
    // records (expand_voxel) equal the device's vertex stream, voxel by voxel
    {
        Renderer raw;
        raw.transport = Transport::Raw;
        raw.load_all_blocking(world, inner);
        Renderer compact;
        compact.threads = 3;
        compact.load_all_blocking(world, inner);
        CHECK(raw.get_voxel_face_count() == compact.get_voxel_face_count() && raw.get_voxel_count() == compact.get_voxel_count());
        for (const auto& kv : raw.meshes()) {
            const RenderArea& other = compact.meshes().at(kv.first);
            CHECK(other.mesh_map.size() == kv.second.mesh_map.size() && other.lights.size() == kv.second.lights.size());
            for (const auto& m : kv.second.mesh_map) {
                const MeshInfo& a = m.second;
                const MeshInfo& b = other.mesh_map.at(m.first);
                CHECK(a.face_count == b.face_count && a.voxel == b.voxel && a.mesh.indices == b.mesh.indices);
                CHECK(a.mesh.vertices.size() == b.mesh.vertices.size());
                CHECK(std::memcmp(a.mesh.vertices.data(), b.mesh.vertices.data(), a.mesh.vertices.size() * sizeof(Vertex)) == 0);
            }
        }
    }
    std::printf("test_host: all checks passed (%zu faces over %zu areas)\n", before, inner.size());
    return 0;
}
