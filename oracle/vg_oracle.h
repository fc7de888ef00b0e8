/*
 * vg_oracle.h -- CPU ORACLE for the voxel_game chunk pipeline.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped library
 * (voxel_game_b200/csrc) never includes, links or calls anything in this directory.
 *
 * It is a plain-C restatement of the reference's (IvanDimovSIT/voxel_game v1.7.0) CPU
 * algorithm for: terrain generation -> column heights ("light") -> face-culled meshing,
 * plus the edit path.  Every function cites the reference file:line it follows.
 *
 * PARITY STATUS
 *   - Game logic (gen rules, trees, heights, face predicate, vertex emission, edits) follows
 *     reference source that IS on disk and is pinned by the reference's own known-answer
 *     tests (tests/test_oracle_known_answers.py).
 *   - "PARITY UNPINNED" at the libnoise boundary: the f64 simplex/fBm arithmetic lives in the
 *     crate libnoise 1.1.2 (+ rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4), which is
 *     not vendored under /root/reference and the reference has no test that pins a noise
 *     value.  noise_libnoise.c restates the published algorithm; see its header for what
 *     is and is not anchored.
 *   - The reference (Rust) cannot be compiled in this image (no cargo/rustc), so there is
 *     no oracle/_ref binary.
 */
#ifndef VG_ORACLE_H
#define VG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/model/area.rs:7-9 */
#define VGO_AREA_SIZE 16
#define VGO_AREA_HEIGHT 128
#define VGO_VOXELS_IN_AREA (VGO_AREA_SIZE * VGO_AREA_SIZE * VGO_AREA_HEIGHT)
#define VGO_COLUMNS_IN_AREA (VGO_AREA_SIZE * VGO_AREA_SIZE)
/* src/model/location.rs:6 */
#define VGO_LOCATION_OFFSET 1000000

/* src/model/voxel.rs:8-36 (implicit discriminants) */
enum {
    VGO_NONE = 0, VGO_COBBLESTONE, VGO_SAND, VGO_GRASS, VGO_WOOD, VGO_LEAVES, VGO_BRICK, VGO_DIRT,
    VGO_BOARDS, VGO_STONE, VGO_CLAY, VGO_LAMP, VGO_TRAMPOLINE, VGO_CACTUS, VGO_WATER_SOURCE,
    VGO_WATER_DOWN, VGO_WATER1, VGO_WATER2, VGO_WATER3, VGO_WATER4, VGO_STONE_BRICK,
    VGO_STONE_PILLAR, VGO_SNOW, VGO_ICE, VGO_BOMB, VGO_ACTIVE_BOMB, VGO_GLASS, VGO_VOXEL_VARIANTS
};

/* Face bits, in the push order of src/graphics/renderer.rs:139-180 */
enum {
    VGO_FACE_LEFT = 1,   /* x+1 */
    VGO_FACE_RIGHT = 2,  /* x-1 */
    VGO_FACE_FRONT = 4,  /* y+1 */
    VGO_FACE_BACK = 8,   /* y-1 */
    VGO_FACE_DOWN = 16,  /* z+1 */
    VGO_FACE_UP = 32     /* z-1 */
};

/* ---------------------------------------------------------------- noise (libnoise 1.1.2) */
typedef struct {
    uint8_t perm[512]; /* PermutationTable::new(seed, 256, doubleup = true) */
} vgo_simplex;

typedef struct {
    const vgo_simplex* src;
    int octaves;
    double frequency, lacunarity, persistence, normalization_factor;
} vgo_fbm;

void vgo_simplex_new(vgo_simplex* s, uint64_t seed);
double vgo_simplex2(const vgo_simplex* s, double x, double y);
double vgo_simplex3(const vgo_simplex* s, double x, double y, double z);
void vgo_fbm_new(vgo_fbm* f, const vgo_simplex* src, int octaves, double frequency,
                 double lacunarity, double persistence);
double vgo_fbm2(const vgo_fbm* f, double x, double y);
double vgo_fbm3(const vgo_fbm* f, double x, double y, double z);

/* Switches for the choices of libnoise 1.1.2 that nothing on this machine pins (noise_libnoise.c
 * header).  0 = the recalled crate behaviour every committed fixture was made with. */
#define VGO_VAR_FBM_RUNNING_FREQ 1u /* Fbm samples point * freq, freq *= lacunarity (default: point *= lacunarity) */
#define VGO_VAR_TIE2_OTHER 2u       /* 2-D middle corner: x0 == y0 goes to (0,1) (default: (1,0)) */
#define VGO_VAR_TIE3_OTHER 4u       /* 3-D traversal: strict > (default: >=) */
#define VGO_VAR_SKEW_EXACT 8u       /* 2-D skew constants: f64 nearest to the exact values (default: 15-digit literals) */
void vgo_set_noise_variant(unsigned flags, const uint8_t* order2, const uint8_t* order3);
unsigned vgo_get_noise_variant(void);
/* f64 additions / multiplications executed by the noise arithmetic on the calling thread since
 * the last reset; zero unless the library was built with -DVGO_COUNT_OPS. */
void vgo_op_counts(uint64_t* adds, uint64_t* muls, int reset);
/* raw pieces exposed for tests */
void vgo_chacha12_words(uint64_t seed_u64, uint32_t* out, int n_words);
uint64_t vgo_siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t* data,
                     size_t len);

/* ---------------------------------------------------------------- generation */
/* src/service/area_generation/generator.rs:24-28 */
uint64_t vgo_hash_world_name(const char* world_name);

typedef struct vgo_generator vgo_generator;
/* generator.rs:135-146 (the five ::new(seed) constructors) */
vgo_generator* vgo_generator_new(uint64_t seed);
void vgo_generator_free(vgo_generator* g);

typedef struct {
    uint64_t simplex2_evals;
    uint64_t simplex3_evals;
    uint64_t cave_voxels;   /* voxels on which should_be_cave evaluated the 14-octave noise */
    uint64_t cave_columns;  /* columns with is_cave_zone */
    uint64_t trees_placed;
} vgo_gen_stats;

/* generator.rs:49-64; voxels[32768] index x + 16y + 256z, max_height[256] index x + 16y */
void vgo_generate_area(const vgo_generator* g, uint32_t area_x, uint32_t area_y, uint8_t* voxels,
                       uint8_t* max_height, vgo_gen_stats* stats);

/* Column samples (generator.rs:108-132), for tests: out = {terrain_height, is_cave_zone,
 * lake_depth, biome(0 dry,1 wet,2 cold)} */
void vgo_sample_column(const vgo_generator* g, uint32_t area_x, uint32_t area_y, uint32_t x,
                       uint32_t y, uint32_t out[4]);

/* Batch = AreaLoader::load_all_blocking (world_persistence.rs:85-94) with every area a miss.
 * rebuild_tables_per_area != 0 mirrors generator.rs:51 (AreaGenerator::new per area).
 * threads <= 0 -> all OpenMP threads.  Returns threads used. */
int vgo_generate_areas(uint64_t seed, const uint32_t* area_xy, int n, uint8_t* voxels,
                       uint8_t* max_height, int rebuild_tables_per_area, int threads,
                       vgo_gen_stats* stats);

/* trees.rs:142-196; returns TreeType discriminant 0 None,1 Short,2 Tall,3 TallCactus,4 DeadTree,
 * 5 HugeTree,6 ShortCactus,7 Bush */
int vgo_should_generate_tree(uint8_t voxel, uint64_t seed, uint32_t area_x, uint32_t area_y,
                             uint32_t lx, uint32_t ly, uint32_t lz);
/* trees.rs:207-241 */
int vgo_can_generate_tree(const uint8_t* voxels, uint32_t lx, uint32_t ly, uint32_t lz, int tree);
void vgo_generate_tree(uint8_t* voxels, uint32_t lx, uint32_t ly, uint32_t lz, int tree);
/* algorithms.rs:28-55 */
uint64_t vgo_combine_seed(uint64_t seed, uint32_t ax, uint32_t ay, uint32_t lx, uint32_t ly,
                          uint32_t lz);
uint64_t vgo_split_mix64(uint64_t seed);

/* ---------------------------------------------------------------- light (column heights) */
/* area.rs:34-56 */
void vgo_update_all_column_heights(const uint8_t* voxels, uint8_t* max_height);
/* area.rs:19-27 */
void vgo_area_new(uint8_t* voxels, uint8_t* max_height);
/* area.rs:99-111 */
void vgo_area_set(uint8_t* voxels, uint8_t* max_height, uint32_t x, uint32_t y, uint32_t z,
                  uint8_t voxel);
/* height_map.rs:41-85: pixels = 512*512*4 RGBA8, NOT cleared (stale pixels persist) */
void vgo_height_map(const uint32_t* area_xy, int n, const uint8_t* max_height /* n*256 */,
                    uint32_t cam_area_x, uint32_t cam_area_y, uint8_t* pixels);

/* ---------------------------------------------------------------- mesh */
/* mesh_generator.rs:501-518 */
int vgo_should_generate_face(uint8_t cur, uint8_t nbr);
int vgo_should_generate_top_face(uint8_t cur, uint8_t nbr);

/* A loaded world: a dense nx*ny grid of areas starting at (ax0, ay0); area (ax, ay) is at
 * grid index (ax-ax0) + nx*(ay-ay0).  Arrays are caller-owned. */
typedef struct {
    uint32_t ax0, ay0, nx, ny;
    uint8_t* voxels;     /* nx*ny*32768 */
    uint8_t* max_height; /* nx*ny*256 */
    uint8_t* face_mask;  /* nx*ny*32768: Renderer.meshes as face bits per voxel; 0 = no entry */
} vgo_world;

/* Renderer::load_full_area (renderer.rs:310-322) = get_renderable_voxels_for_area
 * (world.rs:112-147) + generate_mesh_for_voxel (renderer.rs:126-195).  Writes face_mask for the
 * area.  All four edge neighbours must be inside the world grid.  Returns face count, or -1. */
long vgo_mesh_area(vgo_world* w, uint32_t ax, uint32_t ay);

/* One visible voxel ("MeshInfo" key + value without the vertex payload). 16 bytes. */
typedef struct {
    uint32_t gx, gy;       /* global InternalLocation x, y */
    uint32_t packed;       /* gz | voxel << 8 | face_mask << 16 | n_faces << 24 */
    uint32_t first_face;   /* ordinal of this voxel's first face in the emitted stream */
} vgo_voxel_rec;

/* MeshGenerator::generate_mesh (mesh_generator.rs:109-147,200-498) for every voxel of the area
 * with a non-zero mask, voxels visited in z,y,x order (world.rs:129-131).  vertices: 4 per face,
 * 40 B each {pos f32x3, uv f32x2, color u8x4, normal f32x4}; indices 6 u16 per face.
 * first_face_base is added to rec.first_face.  Any output pointer may be NULL.
 * Returns the number of faces; *n_recs gets the record count. */
long vgo_emit_area(const vgo_world* w, uint32_t ax, uint32_t ay, uint32_t first_face_base,
                   vgo_voxel_rec* recs, long* n_recs, uint8_t* vertices, uint16_t* indices);

/* Batch of vgo_mesh_area + vgo_emit_area (scratch outputs), for CPU-baseline timing.
 * threads = 1: serial like Renderer::load_all_blocking (renderer.rs:468-472); <= 0: all threads. */
long vgo_mesh_areas(vgo_world* w, const uint32_t* area_xy, int n, int threads, long* total_recs);

/* world.set (world.rs:184-191) + renderer.update_location (renderer.rs:239-286).
 * Locations whose area is outside the grid are treated as unloaded (skipped). Returns 0 or -1. */
int vgo_apply_edit(vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz, uint8_t voxel);
/* renderer.update_location alone (the voxel and its <= 6 neighbours re-meshed, no pre-filter) */
void vgo_update_location(vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz);

/* ---------------------------------------------------------------- L-ext (no reference counterpart)
 * Self-specified voxel light field over the whole loaded grid, see DESIGN.md "L-ext".
 * Sequential BFS flood fill; light = nx*ny*32768 bytes, same layout as voxels. */
void vgo_light_flood_world(const vgo_world* w, uint8_t* light);

/* ---------------------------------------------------------------- per-frame visibility
 * Renderer::set_voxel_shader_and_find_visible_areas + render_voxels (renderer.rs:365-462).
 * The f32 vector arithmetic lives in glam 0.27.0 (macroquad's Vec3; crate not under
 * /root/reference, no reference test pins a value: PARITY UNPINNED at that boundary); frame.c
 * restates it: dot = (x*x + y*y) + z*z, length = sqrt(dot), normalize_or_zero = v * (1 / length)
 * if that reciprocal is finite and > 0, else zero. */
typedef struct {
    float cam_pos[3];       /* Camera3D.position */
    float cam_target_z;     /* Camera3D.target.z */
    float look[3];          /* (target - position).normalize_or_zero() */
    uint32_t render_size;   /* UserSettings::get_render_distance(), in areas */
    uint32_t head_in_water; /* PlayerInfo::is_head_in_water */
    uint32_t show_map;      /* RendererParams::should_show_map */
} vgo_view;

/* calculate_area_render_distance (renderer.rs:549-558) */
float vgo_area_render_distance(const float look[3], uint32_t render_size);
/* is_area_visible (renderer.rs:324-346) */
int vgo_is_area_visible(uint32_t ax, uint32_t ay, const vgo_view* v, float render_distance);
/* is_voxel_visible (renderer.rs:560-586) */
int vgo_is_voxel_visible(uint32_t gx, uint32_t gy, uint32_t gz, const float look[3],
                         const float cam_pos[3], float max_distance);
/* prepare_visible_areas + filter_visible_voxels + optimise_render_order + prepare_lights over the
 * meshed areas i = 0..n-1 (records recs[rec_first[i] .. rec_first[i+1])).
 *   area_visible[n]        1 = the area is in visible_areas
 *   visible[]              indices of the drawn records, grouped by voxel type (optimise_render_order);
 *                          inside a group in record order (the reference's order inside a group is
 *                          HashMap iteration order, i.e. unspecified)
 *   type_first[28]         group boundaries in visible[]
 *   lamps[]                indices of the Lamp records of visible areas (prepare_lights), record order
 * Returns the number of visible records; *n_faces gets their face count, *n_lamps the lamp count. */
long vgo_filter_visible(const vgo_voxel_rec* recs, const uint32_t* rec_first, const uint32_t* area_xy,
                        int n, const vgo_view* v, uint8_t* area_visible, uint32_t* visible,
                        uint32_t* type_first, long* n_faces, uint32_t* lamps, long* n_lamps);

/* ---------------------------------------------------------------- area save files (codec.c)
 * file = u32 LE size || LZ4 block of bincode(AreaDTO) (generic_persistence.rs:17-103,
 * world_persistence.rs:27-71, area.rs:204-227); see codec.c for what is and is not pinned. */
#define VGO_AREA_PAYLOAD_BYTES (3 + VGO_VOXELS_IN_AREA)
#define VGO_AREA_FILE_MAX_BYTES 33300
void vgo_bincode_area(const uint8_t* voxels, uint8_t* out);
long vgo_lz4_decompress_block(const uint8_t* src, long n, uint8_t* dst, long cap);
long vgo_lz4_compress_block(const uint8_t* src, long n, uint8_t* dst);
long vgo_area_file_encode(const uint8_t* voxels, uint8_t* out);
int vgo_area_file_decode(const uint8_t* file, long n, uint8_t* voxels);
long vgo_area_files_encode(const uint8_t* voxels, int n, int threads, uint8_t* out, long* sizes);

/* ---------------------------------------------------------------- bomb explosions (physics.c)
 * BombSimulator::handle_explosions / explode_at (bomb_simulator.rs:191-262).  removed / bombs need
 * room for 3 * 1331 * n words. */
#define VGO_EXPLOSION_CANDIDATES 1331 /* (2 * ceil(4.5) + 1)^3 */
/* water / falling-voxel simulators as functions of ORDERED lists (physics.c) */
int vgo_water_tick(vgo_world* w, const uint32_t* check, int n, uint32_t* changed, long* n_changed, uint32_t* next,
                   long* n_next);
int vgo_falling_check(vgo_world* w, const uint32_t* check, int n, uint32_t* spawned, long cap_spawned, long* n_spawned,
                      uint32_t* water_next, long cap_next, long* n_next);
int vgo_falling_land(vgo_world* w, const float* positions, const uint8_t* types, int n, uint8_t* keep, uint32_t* placed,
                     long* n_placed, uint32_t* water_next, long* n_next);
int vgo_explode(vgo_world* w, const float* positions, int n, uint32_t* removed, long* n_removed,
                uint32_t* bombs, long* n_bombs);

int vgo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
