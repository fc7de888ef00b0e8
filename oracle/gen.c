/*
 * gen.c -- ORACLE (test infrastructure only; see vg_oracle.h).
 * Terrain generation + column heights, following src/service/area_generation/ (all .rs files) and
 * src/model/area.rs of the reference.  Plain, sequential, one area at a time; OpenMP over areas
 * only in vgo_generate_areas (mirrors rayon par_iter at world_persistence.rs:90-93).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "vg_oracle.h"

/* Rust `f64 as u32` / `f64 as i32`: truncate toward zero, saturate, NaN -> 0 */
static uint32_t f64_as_u32(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
static int32_t f64_as_i32(double v) {
    if (!(v == v)) return 0;
    if (v <= -2147483648.0) return INT32_MIN;
    if (v >= 2147483647.0) return INT32_MAX;
    return (int32_t)v;
}

/* algorithms.rs:59-61 */
static double normalise_sample(double v) { return (v + 1.0) * 50.0; }

enum { BIOME_DRY = 0, BIOME_WET = 1, BIOME_COLD = 2 };

/* cave_generator.rs:10-19 ; (128 as f32 * 0.2) as u32 = 25, (128 as f32 * 0.9) as u32 = 115 */
#define MIN_CAVES_THRESHOLD 70
#define MAX_CAVES_THRESHOLD 90
#define CAVE_MIN_HEIGHT 25u
#define CAVE_MAX_HEIGHT 115u
#define SHOULD_CHECK_FOR_CAVES_THRESHOLD 80
/* lake_generator.rs:12-15 ; (128*0.32) = 40, (128*0.27) = 34 */
#define LAKE_MAX_TERRAIN_HEIGHT 40u
#define LAKE_MAX_Z_INVERTED 34u
#define MIN_LAKE_HEIGHT 2
#define MIN_LAKE_NOISE 70
/* voxel_type_generator.rs:12-18 ; (128*0.55) = 70 */
#define CLAY_THRESHOLD 70.0
#define ICE_THRESHOLD 60.0
#define BASE_STONE_THRESHOLD 120.0
#define BASE_SNOW_THRESHOLD 80.0
#define SAND_HEIGHT 3u
#define SNOW_HEIGHT 2u
#define SNOW_MOUNTAN_HEIGHT 70u

struct vgo_generator {
    uint64_t seed;
    /* the nine Simplex::new calls of generator.rs:135-146, in constructor order */
    vgo_simplex height1, height2, height_mod; /* terrain_type.rs:16-18 */
    vgo_simplex biome;                        /* biome_type.rs:21 */
    vgo_simplex cave1, cave2, cave_check;     /* cave_generator.rs:27-32 */
    vgo_simplex alt;                          /* voxel_type_generator.rs:25-27 */
    vgo_simplex lake;                         /* lake_generator.rs:22 */
    vgo_fbm f_height1, f_height2, f_height_mod, f_biome, f_cave1, f_cave2, f_cave_check, f_alt,
        f_lake;
};

static void generator_init(vgo_generator* g, uint64_t seed) {
    g->seed = seed;
    vgo_simplex_new(&g->height1, seed);
    vgo_simplex_new(&g->height2, seed ^ UINT64_MAX);
    vgo_simplex_new(&g->height_mod, seed + 1);
    vgo_simplex_new(&g->biome, seed);
    vgo_simplex_new(&g->cave1, seed);
    vgo_simplex_new(&g->cave2, seed ^ UINT64_MAX);
    vgo_simplex_new(&g->cave_check, seed);
    vgo_simplex_new(&g->alt, seed + 1);
    vgo_simplex_new(&g->lake, seed + 10);
    vgo_fbm_new(&g->f_height1, &g->height1, 6, 0.008, 1.8, 0.45);
    vgo_fbm_new(&g->f_height2, &g->height2, 4, 0.004, 1.8, 0.45);
    vgo_fbm_new(&g->f_height_mod, &g->height_mod, 2, 0.004, 0.004, 0.2);
    vgo_fbm_new(&g->f_biome, &g->biome, 2, 0.0015, 1.5, 0.3);
    vgo_fbm_new(&g->f_cave1, &g->cave1, 6, 0.08, 1.3, 0.45);
    vgo_fbm_new(&g->f_cave2, &g->cave2, 8, 0.04, 1.2, 0.35);
    vgo_fbm_new(&g->f_cave_check, &g->cave_check, 2, 0.003, 1.3, 0.3);
    vgo_fbm_new(&g->f_alt, &g->alt, 2, 0.07, 2.0, 0.3);
    vgo_fbm_new(&g->f_lake, &g->lake, 2, 0.004, 1.3, 0.35);
}

vgo_generator* vgo_generator_new(uint64_t seed) {
    vgo_generator* g = (vgo_generator*)malloc(sizeof *g);
    if (g) generator_init(g, seed);
    return g;
}
void vgo_generator_free(vgo_generator* g) { free(g); }

typedef struct {
    uint32_t terrain_height, max_generated_height, lake_depth;
    int is_cave_zone, biome;
} column_samples;

/* generator.rs:108-132 */
static column_samples sample_column_characteristics(const vgo_generator* g, uint32_t ax,
                                                    uint32_t ay, uint32_t x, uint32_t y,
                                                    vgo_gen_stats* st) {
    column_samples c;
    /* algorithms.rs:12-17: u32 arithmetic, then cast */
    double px = (double)(x + ax * VGO_AREA_SIZE), py = (double)(y + ay * VGO_AREA_SIZE);

    /* terrain_type.rs:26-32 */
    double n1 = vgo_fbm2(&g->f_height1, px, py), n2 = vgo_fbm2(&g->f_height2, px, py);
    double value = fmax(n1, n2);
    if (value < 0.1) value = 0.1;
    else if (value > 1.0) value = 1.0;
    double height_mod = 1.0 - (vgo_fbm2(&g->f_height_mod, px, py) + 1.0) / 2.0;
    c.terrain_height =
        f64_as_u32(height_mod * value * 0.55 * (double)VGO_AREA_HEIGHT) + VGO_AREA_HEIGHT / 4;
    st->simplex2_evals += 12;

    /* cave_generator.rs:36-41 */
    c.is_cave_zone = f64_as_i32(normalise_sample(vgo_fbm2(&g->f_cave_check, px, py))) >=
                     SHOULD_CHECK_FOR_CAVES_THRESHOLD;
    st->simplex2_evals += 2;

    /* lake_generator.rs:27-52 */
    if (c.is_cave_zone || c.terrain_height > LAKE_MAX_TERRAIN_HEIGHT ||
        c.terrain_height <= LAKE_MAX_Z_INVERTED) {
        c.lake_depth = 0;
    } else {
        int32_t height_value =
            f64_as_i32(normalise_sample(vgo_fbm2(&g->f_lake, px, py))) - MIN_LAKE_NOISE;
        int32_t lake_depth = height_value / 3;
        c.lake_depth = lake_depth < MIN_LAKE_HEIGHT ? 0u : (uint32_t)lake_depth;
        st->simplex2_evals += 2;
    }

    /* biome_type.rs:25-34 */
    int32_t b = f64_as_i32(normalise_sample(vgo_fbm2(&g->f_biome, px, py)));
    c.biome = (b >= 0 && b < 20) ? BIOME_DRY : (b >= 20 && b < 80) ? BIOME_WET : BIOME_COLD;
    st->simplex2_evals += 2;

    c.max_generated_height = 0;
    return c;
}

/* voxel_type_generator.rs:176-193 */
static int should_generate_alternative_voxel(const vgo_generator* g, double px, double py,
                                             uint32_t z_inverted, double threshold,
                                             vgo_gen_stats* st) {
    st->simplex3_evals += 2;
    return normalise_sample(vgo_fbm3(&g->f_alt, px, py, (double)z_inverted)) >= threshold;
}
static double calculate_height_mix(uint32_t z_inverted, double threshold) {
    return (1.0 - (double)z_inverted / (double)VGO_AREA_HEIGHT) * threshold;
}

/* voxel_type_generator.rs:31-174 */
static uint8_t calculate_voxel_type(const vgo_generator* g, double px, double py, uint32_t zi,
                                    const column_samples* c, vgo_gen_stats* st) {
    uint32_t height = c->terrain_height;
    switch (c->biome) {
    case BIOME_DRY:
        if (zi + SAND_HEIGHT >= height) {
            double thr = calculate_height_mix(zi, BASE_STONE_THRESHOLD);
            return should_generate_alternative_voxel(g, px, py, zi, thr, st) ? VGO_STONE : VGO_SAND;
        }
        return VGO_STONE;
    case BIOME_WET:
        if (zi >= SNOW_MOUNTAN_HEIGHT) {
            if (zi + SNOW_HEIGHT < height) return VGO_STONE;
            double thr = calculate_height_mix(zi, BASE_SNOW_THRESHOLD);
            if (should_generate_alternative_voxel(g, px, py, zi, thr, st)) return VGO_SNOW;
            else if (zi >= height) return VGO_DIRT;
            else return VGO_STONE;
        }
        if (zi >= height)
            return should_generate_alternative_voxel(g, px, py, zi, CLAY_THRESHOLD, st) ? VGO_CLAY
                                                                                        : VGO_GRASS;
        if (zi + 1 >= height)
            return should_generate_alternative_voxel(g, px, py, zi, CLAY_THRESHOLD, st) ? VGO_CLAY
                                                                                        : VGO_DIRT;
        return VGO_STONE;
    default: /* cold */
        if (zi >= height)
            return should_generate_alternative_voxel(g, px, py, zi, ICE_THRESHOLD, st) ? VGO_ICE
                                                                                       : VGO_SNOW;
        if (zi + 1 >= height)
            return should_generate_alternative_voxel(g, px, py, zi, ICE_THRESHOLD, st) ? VGO_ICE
                                                                                       : VGO_DIRT;
        return VGO_STONE;
    }
}

/* lake_generator.rs:54-74; returns -1 for Option::None */
static int lake_generate_voxel(const column_samples* c, uint32_t zi) {
    if (c->lake_depth == 0) return -1;
    if (zi > LAKE_MAX_Z_INVERTED) return VGO_NONE;
    uint32_t min_water = LAKE_MAX_Z_INVERTED - c->lake_depth;
    if (zi >= min_water) {
        if (c->biome == BIOME_COLD && zi == LAKE_MAX_Z_INVERTED) return VGO_ICE;
        return VGO_WATER_SOURCE;
    }
    return -1;
}

/* cave_generator.rs:43-64 */
static int should_be_cave(const vgo_generator* g, const column_samples* c, uint8_t voxel, double px,
                          double py, uint32_t zi, vgo_gen_stats* st) {
    if (!c->is_cave_zone || voxel == VGO_NONE || voxel == VGO_WATER_SOURCE) return 0;
    if (!(zi >= CAVE_MIN_HEIGHT && zi < CAVE_MAX_HEIGHT)) return 0;
    double pz = (double)zi;
    double a = vgo_fbm3(&g->f_cave1, px, py, pz), b = vgo_fbm3(&g->f_cave2, px, py, pz);
    st->simplex3_evals += 14;
    st->cave_voxels += 1;
    int32_t value = f64_as_i32(normalise_sample(fmax(a, b)));
    return value >= MIN_CAVES_THRESHOLD && value < MAX_CAVES_THRESHOLD;
}

/* ------------------------------------------------------------------ trees (trees.rs) */
uint64_t vgo_combine_seed(uint64_t seed, uint32_t ax, uint32_t ay, uint32_t lx, uint32_t ly,
                          uint32_t lz) {
    uint64_t s = seed;
#define ROTL(v, r) (((v) << (r)) | ((v) >> (64 - (r))))
    s *= 0x517cc1b727220a95ULL;
    s += ax;
    s = ROTL(s, 11);
    s += ay;
    s = ROTL(s, 13);
    s += lx;
    s = ROTL(s, 17);
    s += ly;
    s = ROTL(s, 19);
    s += lz;
    s = ROTL(s, 23);
#undef ROTL
    s *= 0xff51afd7ed558ccdULL;
    return s;
}

uint64_t vgo_split_mix64(uint64_t seed) {
    uint64_t z = seed + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

enum { T_NONE = 0, T_SHORT, T_TALL, T_TALL_CACTUS, T_DEAD, T_HUGE, T_SHORT_CACTUS, T_BUSH };

typedef struct {
    int8_t x, y, z;
    uint8_t voxel;
} tree_cell;

/* trees.rs:15-75 */
static const tree_cell SHORT_TREE[] = {{0, 0, -1, VGO_WOOD},   {0, 0, -2, VGO_WOOD},
                                       {0, 0, -3, VGO_WOOD},   {0, 0, -4, VGO_LEAVES},
                                       {0, -1, -3, VGO_LEAVES}, {0, 1, -3, VGO_LEAVES},
                                       {1, 0, -3, VGO_LEAVES},  {-1, 0, -3, VGO_LEAVES}};
static const tree_cell TALL_TREE[] = {{0, 0, -1, VGO_WOOD},    {0, 0, -2, VGO_WOOD},
                                      {0, 0, -3, VGO_WOOD},    {0, 0, -4, VGO_WOOD},
                                      {0, 0, -5, VGO_LEAVES},  {0, -1, -4, VGO_LEAVES},
                                      {0, 1, -4, VGO_LEAVES},  {1, 0, -4, VGO_LEAVES},
                                      {-1, 0, -4, VGO_LEAVES}};
static const tree_cell HUGE_TREE[] = {
    {0, 0, -1, VGO_WOOD},     {0, 0, -2, VGO_WOOD},    {0, 0, -3, VGO_WOOD},
    {0, 0, -4, VGO_WOOD},     {0, 0, -5, VGO_WOOD},    {1, 0, -4, VGO_WOOD},
    {0, 1, -4, VGO_WOOD},     {0, -1, -4, VGO_WOOD},   {-1, 0, -4, VGO_WOOD},
    {0, -1, -5, VGO_LEAVES},  {0, -2, -5, VGO_LEAVES}, {0, 1, -5, VGO_LEAVES},
    {0, 2, -5, VGO_LEAVES},   {1, 0, -5, VGO_LEAVES},  {2, 0, -5, VGO_LEAVES},
    {-1, 0, -5, VGO_LEAVES},  {-2, 0, -5, VGO_LEAVES}, {-1, -1, -5, VGO_LEAVES},
    {-1, 1, -5, VGO_LEAVES},  {1, 1, -5, VGO_LEAVES},  {1, -1, -5, VGO_LEAVES},
    {0, 0, -6, VGO_LEAVES},   {0, 1, -6, VGO_LEAVES},  {1, 0, -6, VGO_LEAVES},
    {0, -1, -6, VGO_LEAVES},  {-1, 0, -6, VGO_LEAVES}, {0, 0, -7, VGO_LEAVES}};
static const tree_cell TALL_CACTUS[] = {
    {0, 0, -1, VGO_CACTUS}, {0, 0, -2, VGO_CACTUS}, {0, 0, -3, VGO_CACTUS}};
static const tree_cell DEAD_TREE[] = {{0, 0, -1, VGO_WOOD}, {0, 0, -2, VGO_WOOD}};
static const tree_cell SHORT_CACTUS[] = {{0, 0, -1, VGO_CACTUS}};
static const tree_cell BUSH[] = {{0, 0, -1, VGO_LEAVES}};

static const tree_cell* tree_cells(int tree, int* n) {
    switch (tree) {
    case T_SHORT: *n = 8; return SHORT_TREE;
    case T_TALL: *n = 9; return TALL_TREE;
    case T_TALL_CACTUS: *n = 3; return TALL_CACTUS;
    case T_SHORT_CACTUS: *n = 1; return SHORT_CACTUS;
    case T_DEAD: *n = 2; return DEAD_TREE;
    case T_HUGE: *n = 27; return HUGE_TREE;
    case T_BUSH: *n = 1; return BUSH;
    default: *n = 0; return NULL;
    }
}

/* trees.rs:104-115 */
static int tree_allows_base(int tree, uint8_t v) {
    switch (tree) {
    case T_SHORT:
    case T_TALL: return v == VGO_GRASS || v == VGO_CLAY;
    case T_TALL_CACTUS:
    case T_SHORT_CACTUS: return v == VGO_SAND;
    case T_DEAD: return v == VGO_GRASS || v == VGO_CLAY || v == VGO_DIRT || v == VGO_SAND;
    case T_HUGE: return v == VGO_GRASS;
    case T_BUSH: return v == VGO_GRASS || v == VGO_SAND;
    default: return 0;
    }
}
/* trees.rs:117-131 */
static uint64_t tree_weight(int tree) {
    switch (tree) {
    case T_SHORT: return 1000;
    case T_TALL: return 1000;
    case T_TALL_CACTUS: return 500;
    case T_SHORT_CACTUS: return 200;
    case T_BUSH: return 100;
    case T_DEAD: return 10;
    case T_HUGE: return 300;
    default: return 0;
    }
}

/* trees.rs:142-196 */
int vgo_should_generate_tree(uint8_t voxel, uint64_t seed, uint32_t ax, uint32_t ay, uint32_t lx,
                             uint32_t ly, uint32_t lz) {
    if (!(voxel == VGO_GRASS || voxel == VGO_DIRT || voxel == VGO_CLAY || voxel == VGO_SAND))
        return T_NONE;
    uint64_t combined = vgo_combine_seed(seed, ax, ay, lx, ly, lz);
    uint64_t random_value = vgo_split_mix64(combined);
    if (((random_value >> 4) % 60) != 0) return T_NONE;

    /* TreeType::ALL_TYPES order, trees.rs:88-96 */
    static const int ALL_TYPES[7] = {T_SHORT, T_TALL, T_TALL_CACTUS, T_SHORT_CACTUS,
                                     T_DEAD,  T_HUGE, T_BUSH};
    int possible[7], n = 0;
    for (int i = 0; i < 7; i++)
        if (tree_allows_base(ALL_TYPES[i], voxel)) possible[n++] = ALL_TYPES[i];
    if (n == 0) return T_NONE;

    uint64_t select = vgo_split_mix64(random_value) >> 4;
    uint64_t total = 0;
    for (int i = 0; i < n; i++) total += tree_weight(possible[i]);
    uint64_t selected = select % total;
    for (int i = 0; i < n; i++) {
        uint64_t w = tree_weight(possible[i]);
        if (w > selected) return possible[i];
        selected -= w;
    }
    return T_NONE;
}

static inline size_t vidx(uint32_t x, uint32_t y, uint32_t z) {
    return (size_t)x + (size_t)y * VGO_AREA_SIZE + (size_t)z * VGO_AREA_SIZE * VGO_AREA_SIZE;
}

/* trees.rs:207-228 */
int vgo_can_generate_tree(const uint8_t* voxels, uint32_t lx, uint32_t ly, uint32_t lz, int tree) {
    int n;
    const tree_cell* cells = tree_cells(tree, &n);
    for (int i = 0; i < n; i++) {
        int ox = cells[i].x + (int)lx, oy = cells[i].y + (int)ly, oz = cells[i].z + (int)lz;
        if (ox < 0 || ox >= VGO_AREA_SIZE) return 0;
        if (oy < 0 || oy >= VGO_AREA_SIZE) return 0;
        if (oz < 0 || oz >= VGO_AREA_HEIGHT) return 0;
        if (voxels[vidx((uint32_t)ox, (uint32_t)oy, (uint32_t)oz)] != VGO_NONE) return 0;
    }
    return 1;
}

/* trees.rs:230-241 */
void vgo_generate_tree(uint8_t* voxels, uint32_t lx, uint32_t ly, uint32_t lz, int tree) {
    int n;
    const tree_cell* cells = tree_cells(tree, &n);
    for (int i = 0; i < n; i++) {
        int ox = cells[i].x + (int)lx, oy = cells[i].y + (int)ly, oz = cells[i].z + (int)lz;
        voxels[vidx((uint32_t)ox, (uint32_t)oy, (uint32_t)oz)] = cells[i].voxel;
    }
}

/* ------------------------------------------------------------------ heights (area.rs) */
static const uint32_t TRANSPARENT_MASK = /* voxel.rs:39-49 */
    (1u << VGO_NONE) | (1u << VGO_GLASS) | (1u << VGO_WATER_SOURCE) | (1u << VGO_WATER_DOWN) |
    (1u << VGO_WATER1) | (1u << VGO_WATER2) | (1u << VGO_WATER3) | (1u << VGO_WATER4) |
    (1u << VGO_ICE);
static inline int is_transparent(uint8_t v) { return (TRANSPARENT_MASK >> v) & 1u; }

/* area.rs:46-56 */
static uint8_t calculate_column_height(const uint8_t* voxels, uint32_t x, uint32_t y) {
    for (uint32_t z = 0; z < VGO_AREA_HEIGHT; z++)
        if (!is_transparent(voxels[vidx(x, y, z)])) return (uint8_t)z;
    return VGO_AREA_HEIGHT - 1;
}

void vgo_update_all_column_heights(const uint8_t* voxels, uint8_t* max_height) {
    for (uint32_t y = 0; y < VGO_AREA_SIZE; y++)
        for (uint32_t x = 0; x < VGO_AREA_SIZE; x++)
            max_height[x + VGO_AREA_SIZE * y] = calculate_column_height(voxels, x, y);
}

void vgo_area_new(uint8_t* voxels, uint8_t* max_height) {
    memset(voxels, VGO_NONE, VGO_VOXELS_IN_AREA);
    memset(max_height, VGO_AREA_HEIGHT - 1, VGO_COLUMNS_IN_AREA);
}

void vgo_area_set(uint8_t* voxels, uint8_t* max_height, uint32_t x, uint32_t y, uint32_t z,
                  uint8_t voxel) {
    voxels[vidx(x, y, z)] = voxel;
    if (voxel == VGO_NONE || is_transparent(voxel)) {
        max_height[x + VGO_AREA_SIZE * y] = calculate_column_height(voxels, x, y);
    } else {
        uint8_t prev = max_height[x + VGO_AREA_SIZE * y];
        if (prev > (uint8_t)z) max_height[x + VGO_AREA_SIZE * y] = (uint8_t)z;
    }
}

/* ------------------------------------------------------------------ generate_area */
void vgo_sample_column(const vgo_generator* g, uint32_t ax, uint32_t ay, uint32_t x, uint32_t y,
                       uint32_t out[4]) {
    vgo_gen_stats st = {0};
    column_samples c = sample_column_characteristics(g, ax, ay, x, y, &st);
    out[0] = c.terrain_height;
    out[1] = (uint32_t)c.is_cave_zone;
    out[2] = c.lake_depth;
    out[3] = (uint32_t)c.biome;
}

void vgo_generate_area(const vgo_generator* g, uint32_t ax, uint32_t ay, uint8_t* voxels,
                       uint8_t* max_height, vgo_gen_stats* stats) {
    vgo_gen_stats local = {0};
    vgo_gen_stats* st = stats ? stats : &local;
    struct {
        uint8_t x, y, z, tree;
    } tree_locations[VGO_COLUMNS_IN_AREA];
    int n_trees = 0;

    vgo_area_new(voxels, max_height); /* generator.rs:53 */
    for (uint32_t x = 0; x < VGO_AREA_SIZE; x++) {
        for (uint32_t y = 0; y < VGO_AREA_SIZE; y++) {
            /* generator.rs:67-105 generate_column */
            column_samples c = sample_column_characteristics(g, ax, ay, x, y, st);
            if (c.is_cave_zone) st->cave_columns++;
            double px = (double)(x + ax * VGO_AREA_SIZE), py = (double)(y + ay * VGO_AREA_SIZE);
            for (uint32_t zi = 1; zi <= c.terrain_height; zi++) {
                uint8_t current = calculate_voxel_type(g, px, py, zi, &c, st);
                int lake = lake_generate_voxel(&c, zi);
                uint8_t lake_voxel = lake >= 0 ? (uint8_t)lake : current;
                if (should_be_cave(g, &c, lake_voxel, px, py, zi, st)) continue;
                if (lake_voxel != VGO_NONE) c.max_generated_height = zi;
                voxels[vidx(x, y, VGO_AREA_HEIGHT - zi)] = lake_voxel;
            }
            uint32_t lz = VGO_AREA_HEIGHT - c.max_generated_height;
            int tree = vgo_should_generate_tree(voxels[vidx(x, y, lz)], g->seed, ax, ay, x, y, lz);
            if (tree != T_NONE) {
                tree_locations[n_trees].x = (uint8_t)x;
                tree_locations[n_trees].y = (uint8_t)y;
                tree_locations[n_trees].z = (uint8_t)lz;
                tree_locations[n_trees].tree = (uint8_t)tree;
                n_trees++;
            }
        }
    }
    /* trees.rs:198-205 */
    for (int i = 0; i < n_trees; i++) {
        if (vgo_can_generate_tree(voxels, tree_locations[i].x, tree_locations[i].y,
                                  tree_locations[i].z, tree_locations[i].tree)) {
            vgo_generate_tree(voxels, tree_locations[i].x, tree_locations[i].y,
                              tree_locations[i].z, tree_locations[i].tree);
            st->trees_placed++;
        }
    }
    vgo_update_all_column_heights(voxels, max_height); /* generator.rs:61 */
}

int vgo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int vgo_generate_areas(uint64_t seed, const uint32_t* area_xy, int n, uint8_t* voxels,
                       uint8_t* max_height, int rebuild_tables_per_area, int threads,
                       vgo_gen_stats* stats) {
    vgo_generator shared;
    if (!rebuild_tables_per_area) generator_init(&shared, seed);
    int used = 1;
    uint64_t s2 = 0, s3 = 0, cv = 0, cc = 0, tp = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) \
    reduction(+ : s2, s3, cv, cc, tp)
#endif
    for (int i = 0; i < n; i++) {
        vgo_gen_stats st = {0};
        if (rebuild_tables_per_area) {
            vgo_generator own; /* generator.rs:51: AreaGenerator::new per area */
            generator_init(&own, seed);
            vgo_generate_area(&own, area_xy[2 * i], area_xy[2 * i + 1],
                              voxels + (size_t)i * VGO_VOXELS_IN_AREA,
                              max_height + (size_t)i * VGO_COLUMNS_IN_AREA, &st);
        } else {
            vgo_generate_area(&shared, area_xy[2 * i], area_xy[2 * i + 1],
                              voxels + (size_t)i * VGO_VOXELS_IN_AREA,
                              max_height + (size_t)i * VGO_COLUMNS_IN_AREA, &st);
        }
        s2 += st.simplex2_evals;
        s3 += st.simplex3_evals;
        cv += st.cave_voxels;
        cc += st.cave_columns;
        tp += st.trees_placed;
    }
    if (stats) {
        stats->simplex2_evals = s2;
        stats->simplex3_evals = s3;
        stats->cave_voxels = cv;
        stats->cave_columns = cc;
        stats->trees_placed = tp;
    }
    return used;
}
