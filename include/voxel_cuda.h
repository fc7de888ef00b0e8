/*
 * voxel_cuda.h -- C ABI of the B200 (sm_100a) chunk pipeline: terrain generation, column heights
 * ("light"), face-culled meshing, batched edits, height-map image.
 *
 * This is the drop-in boundary for the chunk pipeline of IvanDimovSIT/voxel_game v1.7.0.  The
 * reference has no FFI of its own; each entry point below names the Rust function
 * (file:line under the reference's src/) whose body a maintainer replaces with a call to it.
 * INTEGRATION.md shows the Rust `extern "C"` block and the call sites.
 *
 * Conventions: plain pointers and sizes only; every function returns VX_OK (0) or a negative
 * vx_status; no exceptions cross the boundary; all state lives in an opaque vx_ctx bound to ONE
 * CUDA device (one process per GPU, several contexts per process allowed).  A context is
 * internally locked: calls may come from any thread (the reference enters generation from rayon
 * workers, world_persistence.rs:90-93,120-123).  There is NO CPU fallback: without a usable CUDA
 * device vx_create fails with VX_ERR_NO_DEVICE and nothing else can be called.
 *
 * Data layouts (identical to the reference's in-memory layouts):
 *   voxels      32768 bytes per area, index x + 16*y + 256*z, z = 0 is the TOP   (area.rs:58-64)
 *               byte = Voxel discriminant 0..26                                   (voxel.rs:8-36)
 *   max_height  256 bytes per area, index x + 16*y                                (area.rs:29-31)
 *   area_xy     pairs (AreaLocation.x, AreaLocation.y) as u32                     (location.rs:92-96)
 *   vertices    40 bytes each: position f32x3, uv f32x2, color u8x4, normal f32x4
 *               (macroquad::ui::Vertex, repr(C); mesh_generator.rs:225-234), 4 per face
 *   indices     u16, 6 per face: [0,1,2,0,2,3] + 4*k, k = face ordinal inside its voxel
 *               (mesh_generator.rs:44,131-139)
 */
#ifndef VOXEL_CUDA_H
#define VOXEL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_AREA_SIZE 16
#define VX_AREA_HEIGHT 128
#define VX_VOXELS_IN_AREA 32768
#define VX_COLUMNS_IN_AREA 256
#define VX_VERTEX_BYTES 40
#define VX_HEIGHTMAP_SIZE 512

typedef enum {
    VX_OK = 0,
    VX_ERR_NO_DEVICE = -1,    /* no CUDA device / driver: there is no CPU fallback */
    VX_ERR_CUDA = -2,         /* a CUDA call failed; vx_last_error(ctx) has the text */
    VX_ERR_INVALID = -3,      /* bad argument (null pointer, n < 0, voxel id >= 27, z >= 128) */
    VX_ERR_NO_SEED = -4,      /* vx_set_seed has not been called */
    VX_ERR_NOT_LOADED = -5,   /* an area named by the call is not resident in the context */
    VX_ERR_OOM = -6,          /* device or pinned-host allocation failed */
    VX_ERR_NO_MESH = -7       /* vx_mesh_read without a preceding successful vx_mesh */
} vx_status;

typedef struct vx_ctx vx_ctx;

/* One visible voxel: key + value of Renderer's mesh_map entry without the vertex payload
 * (renderer.rs:45-51 `MeshInfo = (u8 face_count, Voxel, Mesh)` keyed by InternalLocation). */
typedef struct {
    uint32_t gx, gy;      /* global InternalLocation x, y (location.rs:44-49) */
    uint32_t packed;      /* gz | voxel << 8 | face_mask << 16 | n_faces << 24 */
    uint32_t first_face;  /* index of the voxel's first face in the call's face stream:
                             vertices[first_face*4 ..][n_faces*4], indices[first_face*6 ..] */
} vx_voxel_rec;
/* face_mask bits, in the push order of renderer.rs:139-180 (so bit order == face order) */
#define VX_FACE_LEFT 1u   /* x+1, normal +x */
#define VX_FACE_RIGHT 2u  /* x-1 */
#define VX_FACE_FRONT 4u  /* y+1 */
#define VX_FACE_BACK 8u   /* y-1 */
#define VX_FACE_DOWN 16u  /* z+1 */
#define VX_FACE_UP 32u    /* z-1 */

/* One block edit = World::set(location, voxel) (world.rs:184-191). */
typedef struct {
    uint32_t gx, gy, gz;
    uint32_t voxel;
} vx_edit;

typedef struct {
    uint64_t n_faces;  /* faces in the stream */
    uint64_t n_recs;   /* visible voxels */
    uint64_t vertex_bytes, index_bytes, rec_bytes;
} vx_mesh_info;

/* Device time of the kernels launched since the previous vx_get_timings (which clears the
 * accumulators), CUDA events on the context's stream.  Only filled while vx_set_timing(ctx, 1).
 * vx_get_timings waits for a vx_generate that returned without waiting (see there). */
typedef struct {
    float gen_columns_ms;  /* K1 */
    float gen_caves_ms;    /* K1b */
    float gen_finish_ms;   /* K2 (trees + column heights) */
    float mesh_count_ms;   /* K4a (+ offsets scan) */
    float mesh_emit_ms;    /* K4b */
    float edit_ms;         /* K5 */
    float heightmap_ms;    /* K6 */
    float light_ms;        /* L-ext */
    float visible_ms;      /* K7 (vx_visible) */
    uint32_t gen_launches, mesh_launches, other_launches; /* kernels launched */
    uint32_t cave_columns;                                /* K1b work items of the last generate */
    float codec_ms;        /* K8 / K9 (vx_encode_areas, vx_decode_areas) */
    /* work accounting of the last vx_generate (roofline numerators): octave-level simplex
     * evaluations executed and voxels that ran the cave noise */
    uint64_t simplex2_evals, simplex3_evals, cave_voxels;
    uint64_t cave_evals; /* of simplex3_evals: cave-noise octaves evaluated (<= 14 per cave voxel:
                            a voxel stops once the remaining octaves cannot change its verdict) */
} vx_timings;

/* ---- lifetime ------------------------------------------------------------------------------ */
int vx_device_count(void);                       /* >= 0, or VX_ERR_NO_DEVICE */
int vx_create(int device, vx_ctx** out);         /* binds to `device`, creates a stream */
void vx_destroy(vx_ctx* ctx);
const char* vx_strerror(int status);
const char* vx_last_error(vx_ctx* ctx);
/* Several calls as one unit.  Every call locks the context for its own duration only, but some
 * results are fetched by a second call (vx_mesh -> vx_mesh_read / vx_mesh_read_compact,
 * vx_encode_areas -> vx_encode_read, vx_visible / vx_visible_zone -> vx_visible_read,
 * vx_update_locations -> vx_update_read; a failed call -> its error text): the second call reads
 * whatever the LAST first call on the context left, so threads that share a context (rayon workers,
 * world_persistence.rs:90-93) bracket such a pair with vx_lock / vx_unlock -- the lock is recursive
 * and held by the calling thread until it unlocks.  The wrappers in rust/, host/ and binding.py do.
 * vx_last_error returns a pointer into the context (valid until the next failing call: use it only
 * under vx_lock); vx_copy_last_error copies the text out under the lock and returns its length. */
int vx_lock(vx_ctx* ctx);
int vx_unlock(vx_ctx* ctx);
size_t vx_copy_last_error(vx_ctx* ctx, char* buf, size_t cap);
/* Run on an external CUDA stream (a cudaStream_t / CUstream handle, e.g. torch's current
 * stream) instead of the context's own; NULL restores the own stream. */
int vx_set_stream(vx_ctx* ctx, void* cuda_stream);
int vx_sync(vx_ctx* ctx);
int vx_set_timing(vx_ctx* ctx, int enabled);
int vx_get_timings(vx_ctx* ctx, vx_timings* out);
/* Pinned host memory for the output buffers (pageable memory works too, only slower). */
void* vx_host_alloc(size_t bytes);
void vx_host_free(void* p);

/* ---- seed: replaces AreaGenerator::new (generator.rs:135-146) and the five noise constructors
 * (terrain_type.rs:15-24, biome_type.rs:19-23, cave_generator.rs:26-34,
 * voxel_type_generator.rs:24-29, lake_generator.rs:21-25).  `seed` = hash_world_name(world_name)
 * (generator.rs:24-28), computed by the Rust side.  Tables are built once per seed instead of
 * once per area. */
int vx_set_seed(vx_ctx* ctx, uint64_t seed);
/* libnoise calibration.  The f64 noise arithmetic follows libnoise 1.1.2 as far as it is known
 * without the crate's source; the choices nothing pins are switches.  Those that
 * change code are compile-time (csrc/vx_tables.hpp VX_NOISE_VARIANT; vx_noise_variant() returns the
 * mask this library was built with, 0 by default), the order of the two gradient tables is data:
 * order2 / order3 are permutations of 0..23 / 0..11 (entry i of the crate's table = entry order[i]
 * of the canonical tables in csrc/vx_tables.cpp), NULL = canonical.  Takes effect for everything
 * generated afterwards (the tables of the current seed are rebuilt).  tests/test_reference_dump.py
 * reports the setting that reproduces a dump of the real crate (rust/parity_dump/). */
int vx_noise_variant(void);
int vx_set_gradient_order(vx_ctx* ctx, const uint8_t* order2, const uint8_t* order3);
/* The same hash, for hosts that are not Rust: SipHash-1-3(k0=k1=0) over name bytes || 0xFF. */
uint64_t vx_hash_world_name(const char* world_name);

/* ---- generation + column heights: replaces AreaGenerator::generate_area (generator.rs:49-64,
 * which ends in Area::update_all_column_heights, area.rs:34-44) for a whole batch, i.e. the body
 * of AreaLoader::load_all_blocking / batch_load (world_persistence.rs:85-110) for disk misses.
 * Areas become resident in the context.  voxels_out (n*32768) / max_height_out (n*256) are host
 * buffers, either may be NULL to leave the result on the device only.  With both NULL the call
 * returns as soon as the kernels are enqueued on the context's stream (world pre-generation keeps
 * the device busy while the host prepares the next call); every later call on the context is
 * ordered behind them, and a failure of the kernels is reported by that call or by vx_sync. */
int vx_generate(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* voxels_out,
                uint8_t* max_height_out);

/* Areas loaded from disk (world_persistence.rs:62-67): upload voxels and compute their column
 * heights on the device, as AreaDTO::into_area does (area.rs:209-219). max_height_out may be NULL. */
int vx_upload(vx_ctx* ctx, const uint32_t* area_xy, int n, const uint8_t* voxels,
              uint8_t* max_height_out);

/* Read resident areas back (World::unload_areas -> store, world.rs:225-240). Outputs may be NULL. */
int vx_read_areas(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* voxels_out,
                  uint8_t* max_height_out);

/* Drop resident areas (World::retain_areas / unload_areas, world.rs:196-240). Unknown areas are
 * ignored. */
int vx_unload(vx_ctx* ctx, const uint32_t* area_xy, int n);
int vx_resident_count(vx_ctx* ctx);
/* The resident-area pool: slots allocated, device memory behind them, and whether the pool arrays are
 * virtual-memory ranges that grow by mapping more physical memory (no copy, stable pointers; DESIGN.md
 * 3) or -- on a driver without that API -- plain allocations that grow by allocate-copy-free.  Any
 * pointer may be NULL. */
int vx_pool_info(vx_ctx* ctx, uint64_t* capacity_areas, uint64_t* committed_bytes, int* virtual_memory);

/* Area::update_all_column_heights (area.rs:34-44) as a stand-alone operator on host data. */
int vx_column_heights(vx_ctx* ctx, const uint8_t* voxels, int n, uint8_t* max_height_out);

/* ---- meshing: replaces Renderer::load_full_area (renderer.rs:310-322) for a batch, i.e. the body
 * of Renderer::load_all_blocking (renderer.rs:468-472) and load_areas_in_queue (:289-307):
 * get_renderable_voxels_for_area (world.rs:112-147) + generate_mesh_for_voxel
 * (renderer.rs:126-195) + MeshGenerator::generate_mesh (mesh_generator.rs:109-147).
 * Edge neighbours that are not resident are generated first, exactly as the reference's
 * world.get_with_cache -> load_area does (world.rs:157-174).  The mesh stays on the device until
 * the next vx_mesh; info tells the caller how large the host buffers for vx_mesh_read must be.
 * Voxels appear in z,y,x order per area, areas in call order, faces in push order. */
int vx_mesh(vx_ctx* ctx, const uint32_t* area_xy, int n, vx_mesh_info* info);

/* Copy the last mesh to host buffers. rec_first/face_first get n+1 entries each (area i owns
 * recs[rec_first[i] .. rec_first[i+1]) and faces [face_first[i] .. face_first[i+1])).
 * Any pointer may be NULL.  Lamp positions (RenderArea.lights, renderer.rs:60-67) are the recs
 * whose voxel byte is 11. */
int vx_mesh_read(vx_ctx* ctx, uint32_t* rec_first, uint32_t* face_first, vx_voxel_rec* recs,
                 void* vertices, uint16_t* indices);

/* ---- edits: replaces World::set (world.rs:184-191 -> Area::set, area.rs:99-111, incl. column
 * height upkeep) for a batch applied in array order (later edits win), and reports the areas whose
 * meshes Renderer::update_location (renderer.rs:239-286) would have touched: the edited area plus
 * the edge neighbour when the edit is on a border column.  The caller then re-runs vx_mesh on
 * dirty_area_xy (room for 3*n pairs) -- or passes dirty_area_xy = n_dirty = NULL and follows up with
 * vx_update_locations, the exact incremental form, which needs no dirty list: the call then returns as
 * soon as its kernels are enqueued.  All edited areas must be resident. */
int vx_apply_edits(vx_ctx* ctx, const vx_edit* edits, int n, uint32_t* dirty_area_xy,
                   int* n_dirty);

/* ---- explosions: replaces the voxel loop of BombSimulator::handle_explosions / explode_at
 * (service/physics/bomb_simulator.rs:191-262) for the explosions of one tick (n <= 32; the
 * reference caps active bombs at 10): every solid voxel (voxel.rs:117-128) whose centre is within
 * 4.5 of a position becomes None (world.set + renderer.update_location).  positions: n Vec3 in
 * player coordinates, processed like the reference from the last to the first.
 *   removed_xyz  internal x, y, z of the removed voxels in the reference's visiting order
 *                (room for 3 * 1331 * n words): the `locations_to_update` the water and
 *                falling-voxel simulators are told about (bomb_simulator.rs:108-116)
 *   bomb_xyz     voxels that were Voxel::Bomb when visited (add_active_bomb_from_explosion), same
 *                order and room
 *   dirty_area_xy as vx_apply_edits (room for 9 * n pairs); the caller re-runs vx_mesh on them
 * All areas an explosion can reach (x, y within 5 voxels of it) must be resident. */
#define VX_EXPLOSION_CANDIDATES 1331
int vx_explode(vx_ctx* ctx, const float* positions, int n, uint32_t* removed_xyz, int* n_removed,
               uint32_t* bomb_xyz, int* n_bombs, uint32_t* dirty_area_xy, int* n_dirty);

/* ---- water and falling voxels: the world-mutating loops of WaterSimulator::simulate_voxels
 * (service/physics/water_simulator.rs:55-77 with check_flow_stopped :79-108, flow_down :184-206,
 * flow_sides :110-150), FallingVoxelSimulator::update_voxels (falling_voxel_simulator.rs:105-150,
 * recursion included) and the retain pass of simulate_falling (:152-166, retain_or_place_voxel
 * :189-221).  These loops mutate the world while they iterate -- the water tick over a HashSet, whose
 * order std randomises per process -- so THE ORDER IS AN INPUT: the caller passes its locations in the
 * order its own iteration yields them and the device applies them one after the other exactly as the
 * reference's loop body does.  Given the same order the result is the reference's.  Every world.set
 * lands in the resident voxels, column heights and bit planes; the changed locations come back in
 * order and go to vx_update_locations (renderer.update_location after every set).  Coordinates are
 * internal (InternalLocation), z < 128.  The areas of every location and of its four side neighbours
 * must be resident (checked up front: VX_ERR_NOT_LOADED and nothing changes).
 *   vx_water_tick     check_xyz: the tick's check_locations (mem::take), n triples.
 *                     changed_xyzv: x, y, z, new voxel per world.set (room for 4 * n quadruples);
 *                     next_xyz: the locations inserted into the NEXT tick's check_locations, in insertion
 *                     order, duplicates included -- the caller's set dedups (room for 7 * n triples).
 *   vx_falling_check  check_xyz: the locations update_voxels is called with, in call order.
 *                     spawned_xyzv: x, y, z, voxel of every falling entity created (= voxel removed;
 *                     position = the location, velocity 0), in creation order; water_next_xyz: what
 *                     water_simulator.location_updated inserted.  A chain can empty whole columns:
 *                     the caller states the capacities in entries; VX_ERR_INVALID if one is too small
 *                     (what was applied until then stays applied and is reported).
 *   vx_falling_land   after the caller has advanced its entities (velocity += delta * 0.2, capped at
 *                     3; position.z += velocity): positions (player coordinates, 3 floats each) and
 *                     voxel types in Vec order.  keep[i] = 1: the entity keeps falling; placed_xyzv:
 *                     the voxels put down (room for n quadruples); water_next_xyz as above (7 * n). */
int vx_water_tick(vx_ctx* ctx, const uint32_t* check_xyz, int n, uint32_t* changed_xyzv, int* n_changed,
                  uint32_t* next_xyz, int* n_next);
int vx_falling_check(vx_ctx* ctx, const uint32_t* check_xyz, int n, uint32_t* spawned_xyzv, int cap_spawned,
                     int* n_spawned, uint32_t* water_next_xyz, int cap_next, int* n_next);
int vx_falling_land(vx_ctx* ctx, const float* positions, const uint8_t* voxel_types, int n, uint8_t* keep,
                    uint32_t* placed_xyzv, int* n_placed, uint32_t* water_next_xyz, int* n_next);

/* ---- height map: replaces HeightMap::generate_height_map's pixel loop (height_map.rs:41-85).
 * rgba is the persistent 512*512*4 pixel buffer (stale pixels persist, as in the reference);
 * areas with |dx| or |dy| >= 16 from the camera area are skipped (height_map.rs:90-101);
 * areas of the list that are not resident read as an empty area (height 127, world.rs:64-69). */
int vx_heightmap(vx_ctx* ctx, const uint32_t* area_xy, int n, uint32_t cam_area_x,
                 uint32_t cam_area_y, uint8_t* rgba);

/* ---- area save files: replaces store_blocking / load_blocking
 * (service/persistence/world_persistence.rs:27-42,62-71) = write/read_binary_object with
 * compression (generic_persistence.rs:17-103) of AreaDTO (model/area.rs:204-227) for batches of
 * areas.  A file is  u32 LE 32771 || LZ4 block of [FB 00 80][32768 voxel bytes]  (lz4_flex
 * compress_prepend_size over bincode config::standard()).  Files written here are plain LZ4
 * blocks any decoder expands to the same bytes (they are not the byte sequence lz4_flex's
 * compressor would pick); any valid file is accepted on the way in. */
#define VX_AREA_FILE_MAX_BYTES 33300
/* Compresses n resident areas on the device; *total_bytes = size of all files together. */
int vx_encode_areas(vx_ctx* ctx, const uint32_t* area_xy, int n, uint64_t* total_bytes);
/* files: total_bytes; offsets: n + 1 (file i = files[offsets[i] .. offsets[i+1])), named
 * "<world>/area<x>_<y>.dat" by the caller (world_persistence.rs:29-31).  Either may be NULL. */
int vx_encode_read(vx_ctx* ctx, uint8_t* files, uint64_t* offsets);
/* Expands n files into resident areas (as vx_upload would) and computes their column heights
 * (AreaDTO::into_area, area.rs:209-219).  ok[i] = 0 where the reference logs an error and generates
 * the area instead (world_persistence.rs:66-70): such areas are NOT resident afterwards and their
 * max_height_out rows are undefined; the caller passes them to vx_generate. */
int vx_decode_areas(vx_ctx* ctx, const uint32_t* area_xy, int n, const uint8_t* files,
                    const uint64_t* offsets, uint8_t* ok, uint8_t* max_height_out);

/* ---- per-frame visibility: replaces, for the mesh of the last vx_mesh call, the work
 * Renderer::set_voxel_shader_and_find_visible_areas and render_voxels do every frame before the
 * draw calls (renderer.rs:365-462): prepare_visible_areas / is_area_visible (:324-346,507-519),
 * filter_visible_voxels / is_voxel_visible -- the per-frame loop over every meshed voxel
 * (:521-547,560-586), optimise_render_order (:349-361) and prepare_lights (:496-501). */
#define VX_VOXEL_VARIANTS 27
#define VX_MAX_LIGHTS 64 /* voxel_shader.rs:21 */
typedef struct vx_view {
    float cam_pos[3];       /* Camera3D.position */
    float cam_target_z;     /* Camera3D.target.z */
    float look[3];          /* (target - position).normalize_or_zero(), as render_voxels computes it */
    uint32_t render_size;   /* UserSettings::get_render_distance(), in areas */
    uint32_t head_in_water; /* PlayerInfo::is_head_in_water: water voxels are not drawn */
    uint32_t show_map;      /* RendererParams::should_show_map: every meshed voxel is drawn */
} vx_view;
typedef struct vx_visible_info {
    uint64_t n_visible; /* voxel meshes to draw (records of the last vx_mesh) */
    uint64_t n_faces;   /* their faces: render_voxels' second return value */
    uint32_t n_areas;   /* visible areas: its first return value */
    uint32_t n_lamps;   /* Lamp records in the visible areas (RenderArea.lights) */
    uint32_t type_first[VX_VOXEL_VARIANTS + 1]; /* group of voxel type t = [type_first[t], type_first[t+1]) */
    uint32_t lamps[VX_MAX_LIGHTS];              /* record indices of the first min(64, n_lamps) lamps */
} vx_visible_info;
/* Runs the pass on the device; the index list stays there until vx_visible_read. */
int vx_visible(vx_ctx* ctx, const vx_view* view, vx_visible_info* info);
/* rec_index: n_visible indices into the records of vx_mesh_read, grouped by voxel type in the
 * order draw_mesh is called (optimise_render_order); inside a group in area-then-record order (the
 * reference's order inside a group is HashMap iteration order).  area_visible: one byte per area of
 * the vx_mesh list.  Either pointer may be NULL. */
int vx_visible_read(vx_ctx* ctx, uint32_t* rec_index, uint8_t* area_visible);

/* ---- incremental re-meshing: replaces Renderer::update_location (renderer.rs:239-286) for a batch
 * of locations (the voxels world.set has just changed -- call vx_apply_edits / vx_explode first):
 * the voxel at each location and its <= 6 neighbours are re-meshed with generate_mesh_for_voxel
 * (renderer.rs:126-195) on the current voxels, i.e. WITHOUT the enclosed-voxel pre-filter of a
 * full-area load (world.rs:137-139) and with should_generate_top_face for the Up face, exactly what
 * applying the locations one by one leaves behind.  Voxels in areas that are not resident are
 * skipped (get_without_loading -> None, renderer.rs:241,281); an area one voxel further that a face
 * test reads is generated first (world.get_with_cache -> load_area, world.rs:157-174).
 * xyz: n triples of internal coordinates (z < 128).  The result is one record per touched voxel
 * (each voxel once, in the order the reference first touches it) with its vertices and indices;
 * a record with n_faces == 0 is RenderArea::remove (set_voxel_mesh with mesh = None,
 * renderer.rs:214-218).  Read it with vx_update_read; it stays valid until the next
 * vx_update_locations.  The mesh of vx_mesh is not disturbed. */
int vx_update_locations(vx_ctx* ctx, const uint32_t* xyz, int n, vx_mesh_info* info);
int vx_update_read(vx_ctx* ctx, vx_voxel_rec* recs, void* vertices, uint16_t* indices);

/* ---- compact transport of the last vx_mesh: 4 bytes per visible voxel instead of 16 + 172 per
 * face.  packed = (x + 16*y + 256*z) | voxel << 15 | face_mask << 20 with x, y area-local; the
 * 16-byte record, the vertices and the indices are a pure function of it and of the area's
 * location (host expander: voxel_game_b200/host/voxel_game.hpp `expand_records`, byte-identical
 * to vx_mesh_read by test).  rec_first: n + 1 entries as in vx_mesh_read.  Either may be NULL. */
#define VX_PACKED_LOCAL(p) ((p) & 0x7FFFu)
#define VX_PACKED_VOXEL(p) (((p) >> 15) & 0x1Fu)
#define VX_PACKED_MASK(p) (((p) >> 20) & 0x3Fu)
int vx_mesh_read_compact(vx_ctx* ctx, uint32_t* rec_first, uint32_t* packed_recs);

/* ---- deferred reads: one host thread keeps the device busy across steps.  While on, vx_encode_read
 * and vx_mesh_read_compact -- and vx_generate when it is asked for the heights only -- enqueue their
 * device -> host copies on the context's copy stream and return at once; the destination buffers
 * (pinned host memory -- a copy into pageable memory is not asynchronous) hold the result after
 * vx_wait_reads.  The next step's vx_generate / vx_encode_areas / vx_mesh can be issued right away:
 * a call that would overwrite what a pending copy still reads waits for it on the device, not on
 * the host.  Use two sets of destination buffers and call vx_wait_reads before consuming (or
 * reusing) the older one.  Replaces nothing in the reference: its loader hands areas to the main
 * thread through a queue (world_persistence.rs:97-123), this is the same decoupling for the device
 * -> host direction.  vx_set_deferred_reads(ctx, 0) waits first. */
int vx_set_deferred_reads(vx_ctx* ctx, int enabled);
int vx_wait_reads(vx_ctx* ctx);

/* ---- mesh retention: a device-resident mirror of `Renderer.meshes: HashMap<AreaLocation,
 * RenderArea>` (renderer.rs:97-101), so that the per-frame pass runs over the whole render zone
 * without re-meshing anything when the zone slides or an edit lands.  While it is on, vx_mesh makes
 * every listed area "meshed" (load_full_area) and keeps its per-voxel face masks; vx_update_locations
 * patches them (creating the RenderArea if the area had none, renderer.rs:208-212);
 * vx_mesh_unload is Renderer::unload_area (renderer.rs:111-117); vx_unload also drops the mesh.
 * Costs 32 KiB + 1 byte of device memory per pool slot.  Off by default (world pre-generation and
 * bulk loads do not need it). */
int vx_set_mesh_retention(vx_ctx* ctx, int enabled);
int vx_mesh_unload(vx_ctx* ctx, const uint32_t* area_xy, int n);
/* The resident RenderArea.mesh_map of n resident areas as dense volumes: masks_out n*32768 bytes,
 * byte x + 16*y + 256*z = face mask of that voxel's mesh, 0 = no entry; meshed_out n bytes,
 * 1 = the area has a RenderArea.  Either may be NULL. */
int vx_read_face_masks(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* masks_out, uint8_t* meshed_out);
/* vx_visible over the resident meshes of the listed areas (the render zone, n <= 131072) instead
 * of the last vx_mesh batch; areas that are not resident or not meshed draw nothing and do not
 * count as visible (prepare_visible_areas iterates over `meshes`, renderer.rs:507-519).
 * vx_visible_read then returns, per drawn voxel, zone index << 15 | (x + 16*y + 256*z) in
 * rec_index (same grouping and order), info.lamps holds the same codes, and area_visible has n
 * entries. */
int vx_visible_zone(vx_ctx* ctx, const uint32_t* area_xy, int n, const vx_view* view, vx_visible_info* info);

/* ---- L-ext (north star only; the reference has no voxel light): light levels
 * 0..15 per voxel over the resident areas listed, see DESIGN.md "L-ext". light_out n*32768. */
int vx_light(vx_ctx* ctx, const uint32_t* area_xy, int n, uint8_t* light_out, int* iterations);

/* ---- cost hint for sharding, computed from the seed alone -- nothing is generated or made resident:
 * per listed area the number of cave-zone columns (CaveGenerator::is_cave_zone,
 * cave_generator.rs:36-41), 0..256, and (cave_voxels_out, may be NULL) the number of voxels of those
 * columns that run the cave noise (should_be_cave's range 25 <= z_inverted < 115 up to the terrain
 * height, cave_generator.rs:43-64).  A cave voxel costs up to 14 3-D noise octaves, a whole ordinary
 * column 16-18 2-D octaves, and caves are not spread evenly: a region is cut into equal-COST strips
 * for several GPUs by these counts (voxel_game_b200/world.py balanced_rows; bench.py --gpus N). */
int vx_cave_columns(vx_ctx* ctx, const uint32_t* area_xy, int n, uint32_t* cave_columns_out,
                    uint32_t* cave_voxels_out);

/* ---- diagnostics: measured FP64 DADD+DMUL issue rate of this GPU in Gop/s (no FMA), the
 * roofline denominator of the noise kernels. */
int vx_diag_fp64_rate(vx_ctx* ctx, double* gops_per_s, float* ms);

#ifdef __cplusplus
}
#endif
#endif
