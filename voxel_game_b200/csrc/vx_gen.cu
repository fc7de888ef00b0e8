// vx_gen.cu -- terrain generation kernels (sm_100a).
//
//   K1  gen_columns : one CTA per area, one thread per voxel column.  Column samples
//                     (generator.rs:108-132), voxel types (voxel_type_generator.rs:31-174; the second
//                     octave of the surface noise, needed by a quarter of the evaluations, is
//                     queued and evaluated densely by the CTA), lake override
//                     (lake_generator.rs:54-74).  The 32 KiB area is assembled in shared
//                     memory and written to HBM with 16-byte coalesced stores.  An area without
//                     cave-zone columns (most) is FINISHED here, still in shared memory: tree
//                     candidates and ordered placement (trees.rs:142-241), column heights
//                     (area.rs:34-56) and the S/T bit planes the mesher reads.  Cave-zone columns
//                     are appended to a work list instead of being carved in place.
//   K1b gen_caves   : persistent (7 CTAs of 128 threads per SM), warp-autonomous lanes over the list of cave-zone columns
//                     (cave_generator.rs:43-64): exact early exit on octave bounds, lanes refill
//                     themselves, warps claim the next chunk while still busy.  FP64-bound part.
//   K2  gen_finish  : persistent CTAs over the areas with cave columns (listed by K1): trees, heights,
//                     planes from HBM after carving.
//   L1  column_heights : Area::update_all_column_heights (+ planes) for uploaded areas.
//
// FP64 only (the reference's libnoise is f64 throughout), no FMA contraction (-fmad=false).
#include "vx_kernels.hpp"
#include "vx_noise.cuh"

namespace vx {

namespace {

enum Biome { DRY = 0, WET = 1, COLD = 2 };

// (128 as f32 * k) as u32 constants of the reference
constexpr unsigned CAVE_MIN_ZI = 25;      // cave_generator.rs:14
constexpr unsigned CAVE_MAX_ZI = 115;     // cave_generator.rs:17
constexpr unsigned LAKE_MAX_TERRAIN = 40; // lake_generator.rs:12
constexpr unsigned LAKE_TOP_ZI = 34;      // lake_generator.rs:13
constexpr unsigned SNOW_MOUNTAIN_ZI = 70; // voxel_type_generator.rs:18

__device__ __forceinline__ double normalise(double v) { return (v + 1.0) * 50.0; }  // algorithms.rs:59

// calculate_height_mix (voxel_type_generator.rs:190-192)
__device__ __forceinline__ double height_mix(unsigned zi, double threshold) {
    return (1.0 - (double)zi / 128.0) * threshold;
}

// Voxel type of one of the four topmost voxels of a column (zi + 3 >= H); everything deeper is
// Stone in all three biomes (voxel_type_generator.rs:64-174).  Split in two so that the expensive
// part -- the 2-octave 3-D noise of should_generate_alternative_voxel -- sits at ONE call site
// that the lanes of a warp reach together:
//   surface_query : does this (biome, zi, H) consult the noise, and against which threshold?
//   surface_pick  : the voxel, given the noise verdict.
__device__ __forceinline__ bool surface_query(int biome, unsigned zi, unsigned H, double& threshold) {
    if (biome == DRY) {  // zi + SAND_HEIGHT >= H holds for every caller
        threshold = height_mix(zi, 120.0);
        return true;
    }
    if (biome == WET) {
        if (zi >= SNOW_MOUNTAIN_ZI) {
            threshold = height_mix(zi, 80.0);
            return zi + 2 >= H;
        }
        threshold = 70.0;
        return zi + 1 >= H;
    }
    threshold = 60.0;
    return zi + 1 >= H;
}
__device__ __forceinline__ uint8_t surface_pick(int biome, unsigned zi, unsigned H, bool alt) {
    if (biome == DRY) return alt ? V_STONE : V_SAND;
    if (biome == WET) {
        if (zi >= SNOW_MOUNTAIN_ZI) {
            if (zi + 2 < H) return V_STONE;
            if (alt) return V_SNOW;
            return zi >= H ? V_DIRT : V_STONE;
        }
        if (zi >= H) return alt ? V_CLAY : V_GRASS;
        if (zi + 1 >= H) return alt ? V_CLAY : V_DIRT;
        return V_STONE;
    }
    if (zi >= H) return alt ? V_ICE : V_SNOW;
    if (zi + 1 >= H) return alt ? V_ICE : V_DIRT;
    return V_STONE;
}

struct TreeCell { int8_t x, y, z; uint8_t voxel; };
enum TreeType { T_NONE = 0, T_SHORT, T_TALL, T_TALL_CACTUS, T_DEAD, T_HUGE, T_SHORT_CACTUS, T_BUSH };

// trees.rs:15-75, indexed by TreeType (trees.rs:77-87)
__constant__ uint8_t c_tree_cells_n[8] = {0, 8, 9, 3, 2, 27, 1, 1};
__device__ const TreeCell c_tree_cells[8][27] = {  // global, not __constant__: the lanes of a warp read 27 different cells (one coalesced load; the constant bank would serialise them)
    {},
    // Short
    {{0, 0, -1, V_WOOD}, {0, 0, -2, V_WOOD}, {0, 0, -3, V_WOOD}, {0, 0, -4, V_LEAVES},
     {0, -1, -3, V_LEAVES}, {0, 1, -3, V_LEAVES}, {1, 0, -3, V_LEAVES}, {-1, 0, -3, V_LEAVES}},
    // Tall
    {{0, 0, -1, V_WOOD}, {0, 0, -2, V_WOOD}, {0, 0, -3, V_WOOD}, {0, 0, -4, V_WOOD},
     {0, 0, -5, V_LEAVES}, {0, -1, -4, V_LEAVES}, {0, 1, -4, V_LEAVES}, {1, 0, -4, V_LEAVES},
     {-1, 0, -4, V_LEAVES}},
    // TallCactus
    {{0, 0, -1, V_CACTUS}, {0, 0, -2, V_CACTUS}, {0, 0, -3, V_CACTUS}},
    // DeadTree
    {{0, 0, -1, V_WOOD}, {0, 0, -2, V_WOOD}},
    // HugeTree
    {{0, 0, -1, V_WOOD}, {0, 0, -2, V_WOOD}, {0, 0, -3, V_WOOD}, {0, 0, -4, V_WOOD},
     {0, 0, -5, V_WOOD}, {1, 0, -4, V_WOOD}, {0, 1, -4, V_WOOD}, {0, -1, -4, V_WOOD},
     {-1, 0, -4, V_WOOD}, {0, -1, -5, V_LEAVES}, {0, -2, -5, V_LEAVES}, {0, 1, -5, V_LEAVES},
     {0, 2, -5, V_LEAVES}, {1, 0, -5, V_LEAVES}, {2, 0, -5, V_LEAVES}, {-1, 0, -5, V_LEAVES},
     {-2, 0, -5, V_LEAVES}, {-1, -1, -5, V_LEAVES}, {-1, 1, -5, V_LEAVES}, {1, 1, -5, V_LEAVES},
     {1, -1, -5, V_LEAVES}, {0, 0, -6, V_LEAVES}, {0, 1, -6, V_LEAVES}, {1, 0, -6, V_LEAVES},
     {0, -1, -6, V_LEAVES}, {-1, 0, -6, V_LEAVES}, {0, 0, -7, V_LEAVES}},
    // ShortCactus
    {{0, 0, -1, V_CACTUS}},
    // Bush
    {{0, 0, -1, V_LEAVES}},
};

__device__ __forceinline__ uint64_t rotl64(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }

// algorithms.rs:41-55
__device__ __forceinline__ uint64_t split_mix64(uint64_t s) {
    uint64_t z = s + 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// should_generate_tree (trees.rs:142-196) with the weighted pick unrolled per base voxel.
// TreeType::ALL_TYPES order = Short, Tall, TallCactus, ShortCactus, DeadTree, HugeTree, Bush;
// weights 1000, 1000, 500, 200, 10, 300, 100 (trees.rs:88-131).
__device__ int pick_tree(uint8_t base, uint64_t seed, uint32_t ax, uint32_t ay, uint32_t lx,
                         uint32_t ly, uint32_t lz) {
    if (!(base == V_GRASS || base == V_DIRT || base == V_CLAY || base == V_SAND)) return T_NONE;
    // combine_seed (algorithms.rs:28-48)
    uint64_t c = seed * 0x517cc1b727220a95ull;
    c = rotl64(c + ax, 11);
    c = rotl64(c + ay, 13);
    c = rotl64(c + lx, 17);
    c = rotl64(c + ly, 19);
    c = rotl64(c + lz, 23);
    c *= 0xff51afd7ed558ccdull;
    const uint64_t r = split_mix64(c);
    if (((r >> 4) % 60ull) != 0) return T_NONE;
    const uint64_t pick = split_mix64(r) >> 4;
    // candidate lists in ALL_TYPES order, filtered by get_allowed_base (trees.rs:104-115)
    switch (base) {
    case V_GRASS: {  // Short 1000, Tall 1000, Dead 10, Huge 300, Bush 100  (sum 2410)
        uint64_t w = pick % 2410ull;
        if (w < 1000) return T_SHORT;
        w -= 1000;
        if (w < 1000) return T_TALL;
        w -= 1000;
        if (w < 10) return T_DEAD;
        w -= 10;
        if (w < 300) return T_HUGE;
        return T_BUSH;
    }
    case V_CLAY: {  // Short 1000, Tall 1000, Dead 10 (sum 2010)
        uint64_t w = pick % 2010ull;
        if (w < 1000) return T_SHORT;
        w -= 1000;
        if (w < 1000) return T_TALL;
        return T_DEAD;
    }
    case V_DIRT:  // Dead only
        return T_DEAD;
    default: {  // Sand: TallCactus 500, ShortCactus 200, Dead 10, Bush 100 (sum 810)
        uint64_t w = pick % 810ull;
        if (w < 500) return T_TALL_CACTUS;
        w -= 500;
        if (w < 200) return T_SHORT_CACTUS;
        w -= 200;
        if (w < 10) return T_DEAD;
        return T_BUSH;
    }
    }
}

// generate_trees (trees.rs:198-241): candidates in push order (s_tree indexed x * 16 + y, the
// column order of generator.rs:54-58; entry = tree type | base z << 8, 0 = none), each checked
// against the voxels as modified by the trees before it.  Run by ONE warp; lanes = template cells.
// `vox` may point to shared memory (K1) or to the area in HBM (K2).  Returns (to every lane) the
// smallest z a placed tree reaches, AREA_HEIGHT if none was placed: above it, and above the
// terrain, everything is still None.
__device__ __forceinline__ int place_trees(uint8_t* vox, const uint16_t* s_tree, unsigned lane) {
    int top = AREA_HEIGHT;
    for (int chunk = 0; chunk < COLUMNS_IN_AREA / 32; ++chunk) {
        const unsigned mine = s_tree[chunk * 32 + lane];
        unsigned pending = __ballot_sync(0xFFFFFFFFu, mine != 0);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const unsigned t = __shfl_sync(0xFFFFFFFFu, mine, src);
            const int o = chunk * 32 + src;
            const int type = t & 0xFF, bz = t >> 8, bx = o >> 4, by = o & 15;
            const int ncell = c_tree_cells_n[type];
            bool ok = true;
            int idx = 0;
            uint8_t cell_voxel = 0;
            if ((int)lane < ncell) {
                const TreeCell c = c_tree_cells[type][lane];
                const int cx = bx + c.x, cy = by + c.y, cz = bz + c.z;
                cell_voxel = c.voxel;
                idx = cx + 16 * cy + 256 * cz;
                ok = cx >= 0 && cx < AREA_SIZE && cy >= 0 && cy < AREA_SIZE && cz >= 0 && cz < AREA_HEIGHT;
                if (ok) ok = vox[idx] == V_NONE;
            }
            if (__all_sync(0xFFFFFFFFu, ok) && (int)lane < ncell) {
                vox[idx] = cell_voxel;
                top = min(top, idx >> 8);
            }
            __syncwarp();  // placed cells are visible to the next candidate's check
        }
    }
    return __reduce_min_sync(0xFFFFFFFFu, top);
}

// S ("voxel != None") and T ("voxel in Voxel::TRANSPARENT") bit planes of 16-voxel rows from the
// voxel bytes (layout: vx_mesh.cu).  One row = one 16-byte load.
__device__ __forceinline__ void planes_of_row(const uint4 q, uint32_t& s_bits, uint32_t& t_bits) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t s = 0, t = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // S: four bytes at once.  ((byte & 0x7F) + 0x7F) | byte has bit 7 set exactly for the non-zero
        // bytes, and nothing carries into the next byte; the four flags are gathered into a nibble by
        // a multiply whose partial products land on distinct bits (bits 24..27 = bytes 0..3).
        const uint32_t nz = ((((w[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w[k]) >> 7) & 0x01010101u;
        s |= ((nz * 0x01020408u) >> 24 & 15u) << (4 * k);
        // T: a bit test in the TRANSPARENT set per byte
#pragma unroll
        for (int b = 0; b < 4; ++b) t |= ((TRANSPARENT_BITS >> ((w[k] >> (8 * b)) & 0xFFu)) & 1u) << (4 * k + b);
    }
    s_bits = s;
    t_bits = t;
}

// Surface voxels whose second noise octave has to be evaluated; lives where the area is assembled
// afterwards (GenSmem::vox).
struct SurfaceQueue {
    double2 acc[4 * COLUMNS_IN_AREA];  // first octave's weighted value, threshold
    uint32_t who[4 * COLUMNS_IN_AREA]; // column | k << 8 | zi << 16
    uint32_t alt[COLUMNS_IN_AREA];     // verdicts coming back: bit k per column
};
static_assert(sizeof(SurfaceQueue) <= VOXELS_IN_AREA, "the queue aliases the voxel buffer");

struct __align__(16) GenSmem {
    uint4 vox[VOXELS_IN_AREA / 16];
    NoiseShared noise;
    unsigned cave_n;
    unsigned cave_base;
    unsigned cave_cursor; // next free entry of the CTA's block of cave items
    unsigned evals[2];    // 2-D / 3-D octave evaluations of the CTA
    unsigned queue_n;     // entries of the SurfaceQueue
    unsigned stone_rows;  // zi <= stone_rows is Stone in every column of the area
    int min_z;            // smallest z holding generated terrain (= 128 - max H)
    uint16_t tree[COLUMNS_IN_AREA];  // tree candidates, indexed x * 16 + y
};

// A two-octave Fbm that the reference only ever compares with integer cuts of
// `normalise(v) as i32` (cave zone: >= 80; biome: 0, 20, 80).  The second octave is bounded by
// |simplex| <= SIMPLEX_BOUND, and normalise / the cast are monotone, so it is evaluated only if it
// can still move the value across a cut.  Returns the number of cuts <= the value.  These noises
// have wavelengths of hundreds of voxels: the 32 columns of a warp nearly always agree.
template <int ID, int NCUTS>
__device__ __forceinline__ int region_of(double v, const int (&cuts)[NCUTS]) {
    const int b = f64_as_i32(normalise(v));
    int r = 0;
#pragma unroll
    for (int i = 0; i < NCUTS; ++i) r += b >= cuts[i] ? 1 : 0;
    return r;
}

// The part of fbm2_region behind the first octave: `acc` holds its weighted value, (x, y) its
// coordinates, (px, py) the point itself.
template <int ID, int NCUTS>
__device__ __forceinline__ int region_finish(const NoiseTables& nt, const NoiseShared* s, double acc, double x, double y,
                                             double px, double py, const int (&cuts)[NCUTS], unsigned& evals) {
    const FbmParams& q = nt.fbm[ID];
    const double mid = acc * q.norm, rest = SIMPLEX_BOUND * q.amp[1] * q.norm + 1e-9;
    const int r_lo = region_of<ID>(mid - rest, cuts), r_hi = region_of<ID>(mid + rest, cuts);
    if (r_lo == r_hi) return r_lo;
#if VX_NOISE_VARIANT & 1
    x = px * q.freq[1];
    y = py * q.freq[1];
#else
    (void)px, (void)py;
    x *= q.lacunarity;
    y *= q.lacunarity;
#endif
    acc += q.amp[1] * simplex2<FBM_TABLE<ID>>(s, x, y);
    ++evals;
    return region_of<ID>(acc * q.norm, cuts);
}

template <int ID, int NCUTS>
__device__ __forceinline__ int fbm2_region(const NoiseTables& nt, const NoiseShared* s, double x, double y,
                                           const int (&cuts)[NCUTS], unsigned& evals) {
    const FbmParams& q = nt.fbm[ID];
    const double px = x, py = y;
    x *= q.frequency;
    y *= q.frequency;
    double acc = 0.0;
    acc += q.amp[0] * simplex2<FBM_TABLE<ID>>(s, x, y);
    ++evals;
    return region_finish<ID>(nt, s, acc, x, y, px, py, cuts, evals);
}

// Two such Fbms over the same point: their first octaves side by side (two independent evaluations in
// flight), then each one's verdict -- and second octave, where it is still needed -- on its own.
template <int IDA, int NA, int IDB, int NB>
__device__ __forceinline__ void fbm2_region_pair(const NoiseTables& nt, const NoiseShared* s, double px, double py,
                                                 const int (&cuts_a)[NA], const int (&cuts_b)[NB], unsigned& evals,
                                                 int& ra, int& rb) {
    const FbmParams& qa = nt.fbm[IDA];
    const FbmParams& qb = nt.fbm[IDB];
    const double xa = px * qa.frequency, ya = py * qa.frequency, xb = px * qb.frequency, yb = py * qb.frequency;
    double acc_a = 0.0, acc_b = 0.0;
    acc_a += qa.amp[0] * simplex2<FBM_TABLE<IDA>>(s, xa, ya);
    acc_b += qb.amp[0] * simplex2<FBM_TABLE<IDB>>(s, xb, yb);
    evals += 2;
    ra = region_finish<IDA>(nt, s, acc_a, xa, ya, px, py, cuts_a, evals);
    rb = region_finish<IDB>(nt, s, acc_b, xb, yb, px, py, cuts_b, evals);
}

#ifndef VX_K1_CTAS_PER_SM
#define VX_K1_CTAS_PER_SM 4
#endif
__global__ void __launch_bounds__(256, VX_K1_CTAS_PER_SM) gen_columns_kernel(const __grid_constant__ NoiseTables nt,
                                                          const uint4* __restrict__ jobs,
                                                          uint8_t* __restrict__ pool_voxels,
                                                          uint8_t* __restrict__ pool_heights,
                                                          uint2* __restrict__ slot_xy,
                                                          CaveItem* __restrict__ cave_items,
                                                          uint32_t* __restrict__ cave_count,
                                                          unsigned long long* __restrict__ stats,
                                                          uint16_t* __restrict__ pool_planes,
                                                          uint32_t* __restrict__ cave_areas) {
    __shared__ __align__(256) GenSmem sm;  // vox is 32 KiB, so the table image behind it is 256-byte aligned
    const int tid = threadIdx.x;
    const uint4 job = jobs[blockIdx.x];
    const uint32_t ax = job.x, ay = job.y, slot = job.z;

    load_noise_shared<FBM_TABLE<FBM_ALT>>(&sm.noise, nt, tid, 256);
    if (tid == 0) {
        sm.cave_n = 0;
        sm.cave_cursor = 0;
        sm.evals[0] = sm.evals[1] = 0;
        sm.queue_n = 0;
        sm.stone_rows = AREA_HEIGHT;
        sm.min_z = AREA_HEIGHT;
        slot_xy[slot] = make_uint2(ax, ay);
    }
    __syncthreads();

    const NoiseShared* s = &sm.noise;
    const unsigned sb = smem_address(sm.noise.bytes), sg = smem_address(sm.noise.grad3);
    uint8_t* vox = reinterpret_cast<uint8_t*>(sm.vox);
    const unsigned x = tid & 15, y = tid >> 4;
    // algorithms.rs:12-17: u32 sum, then cast
    const double px = (double)(x + ax * AREA_SIZE), py = (double)(y + ay * AREA_SIZE);

    // TerrainTypeGenerator::sample (terrain_type.rs:26-32)
#ifdef VX_K1_SERIAL_FBM
    const double n1 = fbm2<FBM_HEIGHT1>(nt, s, px, py), n2 = fbm2<FBM_HEIGHT2>(nt, s, px, py);
    const double nm = fbm2<FBM_HEIGHT_MOD>(nt, s, px, py);
#else
    double n1, n2, nm;  // 6 + 4 + 2 octaves, two evaluations side by side per iteration
    fbm2_three<FBM_HEIGHT1, FBM_HEIGHT2, FBM_HEIGHT_MOD, 6, 4>(nt, s, px, py, n1, n2, nm);
#endif
    double value = fmax(n1, n2);
    value = value < 0.1 ? 0.1 : (value > 1.0 ? 1.0 : value);
    const double height_mod = 1.0 - (nm + 1.0) / 2.0;
    const unsigned H = f64_as_u32(height_mod * value * 0.55 * 128.0) + 32u;
    unsigned evals2 = 12, evals3 = 0;  // simplex evaluations of this column (work accounting)
    // CaveGenerator::is_cave_zone (cave_generator.rs:36-41): normalise(v) as i32 >= 80
    // and BiomeTypeGenerator::sample (biome_type.rs:25-34): 0..20 Dry, 20..80 Wet, everything else (below
    // 0 too) Cold -- the two noises' first octaves run side by side
    const int cave_cuts[1] = {80};
    const int biome_cuts[3] = {0, 20, 80};
    int cave_region, biome_region;
    fbm2_region_pair<FBM_CAVE_CHECK, 1, FBM_BIOME, 3>(nt, s, px, py, cave_cuts, biome_cuts, evals2, cave_region, biome_region);
    const bool cave_zone = cave_region == 1;
    // LakeGenerator::sample_lake_depth (lake_generator.rs:27-52)
    unsigned lake_depth = 0;
    if (!cave_zone && H <= LAKE_MAX_TERRAIN && H > LAKE_TOP_ZI) {
        evals2 += 2;
        const int d = (f64_as_i32(normalise(fbm2<FBM_LAKE>(nt, s, px, py))) - 70) / 3;
        lake_depth = d < 2 ? 0u : (unsigned)d;
    }
    const int biome = biome_region == 1 ? DRY : (biome_region == 2 ? WET : COLD);

    // The four topmost voxels first (k-th voxel from H - 3): should_generate_alternative_voxel
    // (voxel_type_generator.rs:176-188) is a 2-octave Fbm (weights 1/1.3 and 0.3/1.3) against a
    // threshold.  The second octave is bounded by |simplex| <= SIMPLEX_BOUND, so it is needed only
    // where it can still change the verdict (normalise() is monotone: testing it at the bounds is
    // exact) -- about a quarter of the evaluations, scattered over the lanes.  Those are queued and
    // evaluated densely by the whole CTA instead of by half-empty warps.  (Queueing the FIRST octaves
    // too -- only 2.6 of a column's 4 surface voxels consult the noise -- was measured and lost: 0.478 ->
    // 0.497 ms, the listing, the extra barrier and the shared atomics cost more than the idle lanes;
    // so was evaluating two surface voxels side by side per iteration: it spills at the 64-register cap.)
    // A lake column never consults the noise where the lake overrides the voxel
    // (lake_generator.rs:54-74 replaces the value; the reference computes it and throws it away).
    SurfaceQueue& sq = *reinterpret_cast<SurfaceQueue*>(sm.vox);  // the area is assembled there afterwards
    const FbmParams& qa = nt.fbm[FBM_ALT];
    const double alt_rest = SIMPLEX_BOUND * qa.amp[1] * qa.norm + 1e-9;
    unsigned alt_bits = 0;  // bit k: the voxel at zi = H - 3 + k takes its alternative type
    sq.alt[tid] = 0;
#pragma unroll 1
    for (unsigned k = 0; k < 4; ++k) {
        const unsigned zi = H - 3 + k;
        double threshold;
        bool need = surface_query(biome, zi, H, threshold);
        if (lake_depth != 0 && zi >= LAKE_TOP_ZI - lake_depth) need = false;
        bool undecided = false;
        double acc = 0.0;
        if (need) {
            const double x3 = px * qa.frequency, y3 = py * qa.frequency, z3 = (double)zi * qa.frequency;
            acc += qa.amp[0] * simplex3<FBM_TABLE<FBM_ALT>>(sb, sg, x3, y3, z3);
            evals3 += 1;
            const double mid = acc * qa.norm;
            if (normalise(mid - alt_rest) - 1e-9 >= threshold) alt_bits |= 1u << k;
            else if (!(normalise(mid + alt_rest) + 1e-9 < threshold)) undecided = true;
        }
        const unsigned und = __ballot_sync(0xFFFFFFFFu, undecided);
        if (und) {  // warp-aggregated push
            unsigned base = 0;
            if ((tid & 31) == 0) base = atomicAdd(&sm.queue_n, (unsigned)__popc(und));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (undecided) {
                const unsigned slot = base + __popc(und & ((1u << (tid & 31)) - 1u));
                sq.acc[slot] = make_double2(acc, threshold);
                sq.who[slot] = (unsigned)tid | (k << 8) | (zi << 16);
            }
        }
    }
    __syncthreads();
    for (unsigned i = tid; i < sm.queue_n; i += 256) {
        const unsigned who = sq.who[i], c = who & 255u, k = (who >> 8) & 3u, zi = who >> 16;
        const double2 d = sq.acc[i];  // first octave's weighted value, threshold
        // the column's coordinates as its own thread computed them; octave 1 = octave 0 * lacunarity
        const double cx = (double)((c & 15u) + ax * AREA_SIZE), cy = (double)((c >> 4) + ay * AREA_SIZE);
#if VX_NOISE_VARIANT & 1
        const double x3 = cx * qa.freq[1], y3 = cy * qa.freq[1], z3 = (double)zi * qa.freq[1];
#else
        double x3 = cx * qa.frequency, y3 = cy * qa.frequency, z3 = (double)zi * qa.frequency;
        x3 *= qa.lacunarity; y3 *= qa.lacunarity; z3 *= qa.lacunarity;
#endif
        double acc = d.x;
        acc += qa.amp[1] * simplex3<FBM_TABLE<FBM_ALT>>(sb, sg, x3, y3, z3);
        evals3 += 1;
        if (normalise(acc * qa.norm) >= d.y) atomicOr(&sq.alt[c], 1u << k);
    }
    __syncthreads();
    alt_bits |= sq.alt[tid];
    unsigned tops = 0;  // voxel of zi = H - 3 + k in byte k
#pragma unroll
    for (unsigned k = 0; k < 4; ++k) tops |= (unsigned)surface_pick(biome, H - 3 + k, H, (alt_bits >> k) & 1u) << (8 * k);

    // Deep down every column of the area is Stone: the lowest `stone_rows` slabs are filled by the
    // whole CTA with 16-byte stores (and everything above with None) before the per-column loop
    // writes the few voxels that differ.
    atomicMin(&sm.stone_rows, lake_depth != 0 ? LAKE_TOP_ZI - lake_depth - 1 : H - 4);
    atomicMin(&sm.min_z, (int)(AREA_HEIGHT - H));
    if (cave_zone) atomicAdd(&sm.cave_n, 1u);
    __syncthreads();
    const unsigned stone_rows = sm.stone_rows;
    {
        const unsigned s4 = V_STONE * 0x01010101u, first_stone_z = AREA_HEIGHT - stone_rows;
        const uint4 stone = make_uint4(s4, s4, s4, s4), none = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (unsigned k = 0; k < VOXELS_IN_AREA / 16 / 256; ++k) {  // i / 16 = z of the 16-byte row
            const unsigned i = tid + 256 * k;
            if ((i >> 4) >= first_stone_z) sm.vox[i] = stone;
            else sm.vox[i] = none;
        }
    }
    __syncthreads();

    // generate_column (generator.rs:70-98) without the cave test; bottom -> top, above the slabs
    // filled already
    if (lake_depth == 0) {
        // Stone up to four below the surface (voxel_type_generator.rs:64-174), then the surface voxels
        uint8_t* col = vox + tid + 256 * AREA_HEIGHT;  // col[-256 * zi] is the voxel at zi
        for (unsigned zi = stone_rows + 1; zi + 3 < H; ++zi) col[-256 * (int)zi] = V_STONE;
#pragma unroll
        for (unsigned k = 0; k < 4; ++k) col[-256 * (int)(H - 3 + k)] = (uint8_t)(tops >> (8 * k));
    } else {
        for (unsigned zi = stone_rows + 1; zi <= H; ++zi) {
            uint8_t v;
            if (zi > LAKE_TOP_ZI) {
                v = V_NONE;  // lake_generator.rs:59-61
            } else if (zi >= LAKE_TOP_ZI - lake_depth) {
                v = (biome == COLD && zi == LAKE_TOP_ZI) ? V_ICE : V_WATER_SOURCE;
            } else if (zi + 3 < H) {
                v = V_STONE;
            } else {
                v = (uint8_t)(tops >> (8 * (zi + 3 - H)));
            }
            vox[tid + 256 * (AREA_HEIGHT - zi)] = v;
        }
    }

    // work accounting: one pair of global atomics per CTA
    evals2 = __reduce_add_sync(0xFFFFFFFFu, evals2);
    evals3 = __reduce_add_sync(0xFFFFFFFFu, evals3);
    if ((tid & 31) == 0) {
        atomicAdd(&sm.evals[0], evals2);
        atomicAdd(&sm.evals[1], evals3);
    }
    uint8_t* heights = pool_heights + (size_t)slot * COLUMNS_IN_AREA;
    __syncthreads();  // the area is complete in shared memory
    const unsigned n_cave = sm.cave_n;  // final since the barrier before the fill
    if (tid == 0) {
        atomicAdd(&stats[0], (unsigned long long)sm.evals[0]);
        atomicAdd(&stats[1], (unsigned long long)sm.evals[1]);
    }

    if (n_cave != 0) {
        // Cave zone in this area: K1b carves it in HBM and K2 finishes it (trees, heights, planes).
        // Scratch for both: terrain height (7 bits) | cave flag per column; work items per cave column.
        heights[tid] = (uint8_t)(H | (cave_zone ? 0x80u : 0u));
        if (tid == 0) {
            sm.cave_base = atomicAdd(cave_count, n_cave);
            // K2's work list: the jobs that still need finishing (cave_count + 1 counts them)
            cave_areas[atomicAdd(cave_count + 1, 1u)] = blockIdx.x;
        }
        __syncthreads();
        if (cave_zone) cave_items[sm.cave_base + atomicAdd(&sm.cave_cursor, 1u)] = slot * COLUMNS_IN_AREA + tid;
    } else {
        // No caves (most areas): trees, column heights and bit planes straight from shared memory.
        // The topmost generated non-None voxel (generator.rs:91-93,100) is the surface voxel, or the
        // lake surface (water / ice cap at zi = 34) in a lake column.
        const unsigned z_top = AREA_HEIGHT - (lake_depth != 0 ? LAKE_TOP_ZI : H);
        const int tree = pick_tree(vox[tid + 256u * z_top], nt.seed, ax, ay, x, y, z_top);
        sm.tree[x * 16 + y] = tree ? (uint16_t)(tree | (z_top << 8)) : (uint16_t)0;
        __syncthreads();
        if (tid < 32) {
            const int tree_top = place_trees(vox, sm.tree, tid);
            if (tid == 0) sm.min_z = min(sm.min_z, tree_top);  // topmost slab that holds anything
        }
        __syncthreads();
        // Column heights (update_all_column_heights, area.rs:34-56) and the bit planes in one walk.
        // Slabs above z_lo (the highest terrain or tree voxel) are all None, slabs from z_hi
        // on are all Stone -- so a column's height is z_hi unless something opaque lies above it.
        // In between one ballot per slab and warp gives the two rows (y = 2 * warp, 2 * warp + 1)
        // of both planes.
        const int z_lo = sm.min_z, z_hi = AREA_HEIGHT - (int)stone_rows;
        uint16_t* planes = pool_planes + (size_t)slot * (VOXELS_IN_AREA / 8);
        {
            const int z = tid >> 1;  // thread i owns rows 8i .. 8i+7 = half of slab z = i / 2
            if (z < z_lo || z >= z_hi) {
                const uint32_t sv = z < z_lo ? 0u : 0xFFFFFFFFu;
                reinterpret_cast<uint4*>(planes)[tid] = make_uint4(sv, sv, sv, sv);
                reinterpret_cast<uint4*>(planes + VOXELS_IN_AREA / 16)[tid] = make_uint4(~sv, ~sv, ~sv, ~sv);
            }
        }
        const int wid = tid >> 5;
        uint32_t* s_words = reinterpret_cast<uint32_t*>(planes) + wid;
        uint32_t* t_words = reinterpret_cast<uint32_t*>(planes + VOXELS_IN_AREA / 16) + wid;
        int top = z_hi;  // first z from the top that is not transparent
        for (int z = z_hi - 1; z >= z_lo; --z) {
            const unsigned v = vox[tid + 256 * z];
            const bool transparent = is_transparent(v);
            const unsigned sb = __ballot_sync(0xFFFFFFFFu, v != 0u);
            const unsigned tb = __ballot_sync(0xFFFFFFFFu, transparent);
            if (!transparent) top = z;
            if ((tid & 31) == 0) {  // rows r = 16 z + 2 wid and r + 1 are one 32-bit word
                s_words[8 * z] = sb;
                t_words[8 * z] = tb;
            }
        }
        heights[tid] = (uint8_t)top;
    }

    // The 32 KiB area leaves as ONE TMA bulk copy (shared -> global): every thread publishes its
    // shared-memory writes to the async proxy, one thread issues the copy and stays until the
    // area has been read; the other warps are done.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        uint8_t* dst = pool_voxels + (size_t)slot * VOXELS_IN_AREA;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(sm.vox)), "r"((uint32_t)VOXELS_IN_AREA)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ------------------------------------------------------------------------------------------- K1b
#ifndef VX_CAVE_CHUNK
#define VX_CAVE_CHUNK 8
#endif
#ifndef VX_CAVE_TAIL
#define VX_CAVE_TAIL 4  // columns per claim near the end of the list (0: always VX_CAVE_CHUNK)
#endif
// Cave columns a warp claims at a time (lane c < CAVE_CHUNK describes column c).  A column is 8..78
// voxels and a region has ~87 000 cave columns for 3 552 warps: with 32 columns per claim most
// warps got exactly one chunk (a quarter of them none) and the kernel lasted as long as the
// heaviest chunk; small chunks let the global counter even the warps out.
constexpr int CAVE_CHUNK = VX_CAVE_CHUNK;
constexpr int CAVE_SLOTS = 32;  // table entries per warp (one per lane; those past CAVE_CHUNK stay empty)
static_assert(CAVE_CHUNK >= 1 && CAVE_CHUNK <= CAVE_SLOTS && (CAVE_CHUNK & (CAVE_CHUNK - 1)) == 0,
              "one column per lane; a power of two for the column search");
// The 14 cave octaves are evaluated in rounds: round k = octave k of noise2 (8 octaves, weights
// .650 .228 .080 .028 ...) and, for k < 6, octave k of noise1 (6 octaves, .553 .249 .112 .050 ...).

// 128-thread CTAs, 7 per SM: 72 registers per thread without spills, 28 resident warps per SM (256 x 3
// at 79 registers gives 24; 128 x 8 needs a 64-register cap that spills) -- profiles/r02_ab_caves.log
#ifndef VX_CAVE_CTAS_PER_SM
#define VX_CAVE_CTAS_PER_SM 7
#endif
#ifndef VX_CAVE_THREADS
#define VX_CAVE_THREADS 128
#endif
struct CaveWarp {  // per-warp description of the claimed chunk
    uint32_t first[CAVE_SLOTS + 1];  // exclusive prefix of per-column voxel counts
    uint32_t item[CAVE_SLOTS];       // slot * 256 + column
    double2 oa[CAVE_CHUNK];          // the column's x, y scaled by the base frequency of noise1 ...
    double2 ob[CAVE_CHUNK];          // ... and of noise2 (octave 0 coordinates)
};
struct CaveSmem {
    alignas(16) uint8_t bytes[TAB_BYTES];       // the (doubled) permutation tables (vx_noise.cuh)
    NoiseImage::Code grad3[2][TAB_STRIDE];      // 3-D gradient codes of the two cave tables
    double amp[2][8];              // [0] = noise1, [1] = noise2: running amplitudes
#if VX_NOISE_VARIANT & 1
    double freq[2][8];             // running frequencies (the other Fbm form)
#endif
    // Per evaluated-octave count s and Fbm f: the raw running sum x of f decides "norm_f * x + rest_f[s] < c"
    // and "norm_f * x - rest_f[s] >= c" for c = 0.4, 0.8 by comparing the HIGH WORD of x, mapped to a
    // monotone integer key, with the key of a precomputed threshold: {below 0.4, below 0.8, at least
    // 0.4, at least 0.8}.  key(x) < key(T) implies x < T, key(x) > key(T) implies x > T; equal high words
    // decide nothing, which only postpones a verdict (the bound is 2^-20 wider).
    int4 thr[2][9];
    double2 zf[AREA_HEIGHT];       // zi * frequency of noise1 / noise2: octave-0 z coordinates
    CaveWarp warp[VX_CAVE_THREADS / 32];
};

// Monotone integer key of the high word of a double: a < b as doubles whenever key(a) < key(b).
__device__ __forceinline__ int hi_key(double x) {
    const int h = __double2hiint(x);
    return h ^ ((h >> 31) & 0x7FFFFFFF);
}

// Max<Fbm, Fbm>::sample + threshold (cave_generator.rs:27-30,43-64) for every voxel of the listed
// cave-zone columns:  cave  <=>  (70..90).contains(normalise(max(a, b)) as i32)  <=>  0.4 <= m < 0.8
// up to rounding at the ends.
//   * Exact early exit.  The octaves are evaluated in rounds of one octave per Fbm (each Fbm keeps
//     its own octave order, so its running sum is the crate's); after every round the not yet
//     evaluated octaves are bounded by |simplex| <= SIMPLEX_BOUND, and a voxel stops as soon as
//     the verdict can no longer change.  Undecided voxels run all 14 octaves and use the crate's
//     expression, so the result never depends on the bound being tight, only on it being valid
//     (slack 1e-9 against rounding).  ~5 of 14 octaves are needed on average.
//   * Warp-autonomous persistent lanes.  A warp claims 32 columns from a global counter and
//     flattens their (column, z) voxels; a lane that has decided its voxel takes the next unclaimed
//     one at once (ballot + popc, no atomics, no CTA barrier) instead of idling until the slowest
//     lane is done.  Lanes of a warp sit at different octaves of different voxels, which costs
//     nothing because the octave index only selects table rows.
//   * Measured and rejected (round 2; both bit-exact, the code is in the history at e074ab7): an Fbm
//     whose upper bound has fallen below 0.4 cannot change the verdict any more, so evaluating one
//     octave per iteration, of the Fbm that still matters, needs 3.49 instead of 4.47 evaluations per
//     voxel -- but one evaluation in flight per thread is slower than the two of a round (0.304 ->
//     0.320 ms, profiles/r02_ab_caves_single.log); and two voxels per lane, one octave of each per
//     iteration with the per-voxel state in shared memory, pays the refill and the verdict once per
//     evaluation instead of once per two (0.329 ms at 72 registers, profiles/r02_ab_caves_dual.log).
__global__ void __launch_bounds__(VX_CAVE_THREADS, VX_CAVE_CTAS_PER_SM) gen_caves_kernel(const __grid_constant__ NoiseTables nt,
                                                        const CaveItem* __restrict__ cave_items,
                                                        const uint32_t* __restrict__ cave_count,
                                                        uint8_t* __restrict__ pool_voxels,
                                                        const uint8_t* __restrict__ pool_heights,
                                                        const uint2* __restrict__ slot_xy,
                                                        unsigned long long* __restrict__ stats) {
    __shared__ __align__(256) CaveSmem sm;  // table image first: 256-byte aligned for simplex3
    const int tid = threadIdx.x, lane = tid & 31;
    const FbmParams& q1 = nt.fbm[FBM_CAVE1];
    const FbmParams& q2 = nt.fbm[FBM_CAVE2];
    static_assert(FBM_TABLE<FBM_CAVE1> == 0 && FBM_TABLE<FBM_CAVE2> == 1, "grad3[0..1] are the cave tables");
    load_noise_image(sm.bytes, nt, 0, TAB_BYTES, tid, VX_CAVE_THREADS);
    load_noise_image(sm.grad3, nt, (unsigned)offsetof(NoiseImage, grad3), (unsigned)sizeof(sm.grad3), tid, VX_CAVE_THREADS);
    if (tid < 8) {
        sm.amp[0][tid] = tid < 6 ? q1.amp[tid] : 0.0;
        sm.amp[1][tid] = q2.amp[tid];
#if VX_NOISE_VARIANT & 1
        sm.freq[0][tid] = tid < 6 ? q1.freq[tid] : 0.0;
        sm.freq[1][tid] = q2.freq[tid];
#endif
    }
    if (tid < 18) {
        const int* t = &nt.cave_thr[0][0][0] + 4 * tid;  // thresholds as integer keys, made by the host (vx_tables.cpp)
        (&sm.thr[0][0])[tid] = make_int4(t[0], t[1], t[2], t[3]);
    }
    if (tid < AREA_HEIGHT) {
#if VX_NOISE_VARIANT & 1
        sm.zf[tid] = make_double2((double)tid, (double)tid);
#else
        sm.zf[tid] = make_double2((double)tid * q1.frequency, (double)tid * q2.frequency);
#endif
    }
    __syncthreads();
    const uint32_t n_items = *cave_count;
    // (a claim size that shrinks with the list -- 4, 2, 1 columns for the short lists of a sharded step --
    // was measured and lost: 0.069 -> 0.083 ms at N = 8; a column here has ~16 cave voxels, so a small
    // claim is used up in one pass of the warp and the claiming itself becomes the work)
    constexpr uint32_t claim = CAVE_CHUNK;
    const double f1 = q1.frequency, f2 = q2.frequency, l1 = q1.lacunarity, l2 = q2.lacunarity;
    const double n1 = q1.norm, n2 = q2.norm;
    const unsigned sb = smem_address(sm.bytes), sg1 = smem_address(sm.grad3[0]), sg2 = smem_address(sm.grad3[1]);
    CaveWarp& cw = sm.warp[tid >> 5];
    unsigned* chunk_counter = reinterpret_cast<unsigned*>(&stats[4]);
    const unsigned lanes_below = (1u << lane) - 1u;
    unsigned my_evals = 0;

    uint32_t next = 0, total = 0;  // unclaimed voxels [next, total) of the warp's current chunk
    bool more = true;              // the global list still has unclaimed chunks
    bool busy = false;
    double xa = 0, ya = 0, za = 0, xb = 0, yb = 0, zb = 0, a = 0, bsum = 0;
    unsigned step = 0;
    size_t out_index = 0;
    while (true) {
        const unsigned want = __ballot_sync(0xFFFFFFFFu, !busy);
        if (want != 0 && next >= total && more) {
            // The current chunk is used up: claim the next columns right away.  Lanes that are
            // still evaluating keep their state in registers, so the warp never drains.
#if VX_CAVE_TAIL
            // Guided claims: CAVE_CHUNK columns while plenty are left, VX_CAVE_TAIL once the rest of the
            // list is less than one full claim per warp of the grid -- the last claims are what the
            // kernel's end waits for, and smaller ones even the warps out.  The counter counts columns;
            // the size is picked from a plain (possibly stale) read of it, which can only pick a size.
            // Tail claims of 4: 0.3030 -> 0.2975 ms for the whole region, no change for a rank's strip
            // of eight; of 2 or 1: 0.319 / 0.379 -- a claim is a chain of four dependent global accesses
            // that the whole warp waits for (profiles/r02_ab_cave_claims.log; sizing the claim from the
            // warp's previous claim instead of the extra read never shrinks it: the counter moves by
            // one whole tail between two claims of a warp; fetching the NEXT claim one access per
            // iteration while the current one is worked on, so that no access is waited for: 0.298 ->
            // 0.330 ms -- the per-iteration stage logic costs more than the waits it removes).
            unsigned first_col = 0, seen = 0;
            if (lane == 0) seen = *reinterpret_cast<volatile unsigned*>(chunk_counter);
            seen = __shfl_sync(0xFFFFFFFFu, seen, 0);
            const unsigned tail = gridDim.x * (VX_CAVE_THREADS / 32) * CAVE_CHUNK;
            const unsigned size = (seen < n_items && n_items - seen > tail) ? (unsigned)CAVE_CHUNK : (unsigned)VX_CAVE_TAIL;
            if (lane == 0) first_col = atomicAdd(chunk_counter, size);
            first_col = __shfl_sync(0xFFFFFFFFu, first_col, 0);
            (void)claim;
            if (first_col >= n_items) {
                more = false;
            } else {
                const uint32_t idx = first_col + lane;
                uint32_t item = 0, cnt = 0;
                double px = 0.0, py = 0.0;
                if (lane < (int)size && idx < n_items) {
#else
            unsigned chunk = 0;
            if (lane == 0) chunk = atomicAdd(chunk_counter, 1u);
            chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
            if ((unsigned long long)chunk * claim >= n_items) {
                more = false;
            } else {
                const uint32_t idx = chunk * claim + lane;
                uint32_t item = 0, cnt = 0;
                double px = 0.0, py = 0.0;
                if (lane < (int)claim && idx < n_items) {
#endif
                    item = cave_items[idx];
                    const unsigned H = pool_heights[item] & 0x7Fu;
                    const uint2 axy = slot_xy[item >> 8];
                    // algorithms.rs:19-26: u32 sum, then cast
                    px = (double)((item & 15u) + axy.x * AREA_SIZE);
                    py = (double)(((item >> 4) & 15u) + axy.y * AREA_SIZE);
                    // should_be_cave evaluates the noise for CAVE_MIN_ZI <= zi < CAVE_MAX_ZI, zi <= H
                    // (in a cave zone lake_depth is 0, so the voxel is never None/WaterSource)
                    const unsigned top = H < CAVE_MAX_ZI - 1 ? H : CAVE_MAX_ZI - 1;
                    cnt = top >= CAVE_MIN_ZI ? top - CAVE_MIN_ZI + 1 : 0;
                }
                uint32_t v = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
                    if (lane >= d) v += o;
                }
                __syncwarp();  // nobody is still reading the previous chunk's tables
                cw.item[lane] = item;
                if (lane < CAVE_CHUNK) {
#if VX_NOISE_VARIANT & 1
                    cw.oa[lane] = make_double2(px, py);  // raw: every octave multiplies by its own frequency
                    cw.ob[lane] = make_double2(px, py);
#else
                    cw.oa[lane] = make_double2(px * f1, py * f1);
                    cw.ob[lane] = make_double2(px * f2, py * f2);
#endif
                }
                cw.first[lane + 1] = v;
                if (lane == 0) cw.first[0] = 0;
                __syncwarp();
                total = cw.first[CAVE_SLOTS];
                next = 0;
                if (lane == 0) atomicAdd(&stats[2], (unsigned long long)total);  // cave voxels evaluated
            }
        }
        // idle lanes take the next voxels of the chunk, in lane order
        if (!busy) {
            const uint32_t w = next + __popc(want & lanes_below);
            if (w < total) {
                int lo = 0, hi = CAVE_CHUNK;  // column with first[lo] <= w < first[lo + 1]
#pragma unroll
                for (int s = 0; (1 << s) < CAVE_CHUNK; ++s) {
                    const int mid = (lo + hi) >> 1;
                    if (cw.first[mid] <= w) lo = mid; else hi = mid;
                }
                const uint32_t item = cw.item[lo];
                const unsigned zi = CAVE_MIN_ZI + (w - cw.first[lo]);
                const double2 oa = cw.oa[lo], ob = cw.ob[lo], oz = sm.zf[zi];
                xa = oa.x; ya = oa.y; za = oz.x;
                xb = ob.x; yb = ob.y; zb = oz.y;
                a = 0.0; bsum = 0.0;
                step = 0;
                out_index = (size_t)(item >> 8) * VOXELS_IN_AREA + (item & 255u) + 256u * (AREA_HEIGHT - zi);
                busy = true;
            }
        }
        next = min(total, next + (uint32_t)__popc(want));
        if (!__any_sync(0xFFFFFFFFu, busy)) {
            if (!more && next >= total) break;  // list exhausted and nothing in flight
            continue;                           // an empty chunk (or lanes waiting for the next one)
        }
        if (busy) {
            // one octave of each Fbm per round (noise1 has 6, noise2 has 8): no per-lane selection
            // of coordinates, and two independent evaluations in flight
            const unsigned k = step;
#if VX_NOISE_VARIANT & 1
            const double fb = sm.freq[1][k];
            const double svb = simplex3<FBM_TABLE<FBM_CAVE2>>(sb, sg2, xb * fb, yb * fb, zb * fb);
            bsum += sm.amp[1][k] * svb;
            ++my_evals;
            if (k < 6) {
                const double fa = sm.freq[0][k];
                const double sva = simplex3<FBM_TABLE<FBM_CAVE1>>(sb, sg1, xa * fa, ya * fa, za * fa);
                a += sm.amp[0][k] * sva;
                ++my_evals;
            }
            (void)l1, (void)l2, (void)f1, (void)f2;
#else
            const double svb = simplex3<FBM_TABLE<FBM_CAVE2>>(sb, sg2, xb, yb, zb);
            bsum += sm.amp[1][k] * svb;
            xb *= l2; yb *= l2; zb *= l2;
            ++my_evals;
            if (k < 6) {
                const double sva = simplex3<FBM_TABLE<FBM_CAVE1>>(sb, sg1, xa, ya, za);
                a += sm.amp[0][k] * sva;
                xa *= l1; ya *= l1; za *= l1;
                ++my_evals;
            }
#endif
            ++step;
            // the verdict on integer keys of the two running sums (CaveSmem::thr): no f64 arithmetic
            const int ka = hi_key(a), kb = hi_key(bsum);
            const int4 ta = sm.thr[0][step < 6 ? step : 6], tb = sm.thr[1][step];
            int verdict = -1;  // -1 undecided, 0 solid, 1 cave
            if ((ka < ta.x && kb < tb.x) || ka > ta.w || kb > tb.w) verdict = 0;
            else if ((ka > ta.z || kb > tb.z) && ka < ta.y && kb < tb.y) verdict = 1;
            if (verdict < 0 && step == 8) {  // all octaves done: the crate's own expression
                const int c = f64_as_i32(normalise(fmax(a * n1, bsum * n2)));
                verdict = (c >= 70 && c < 90) ? 1 : 0;  // cave_generator.rs:10-11,63
            }
            if (verdict >= 0) {
                if (verdict) pool_voxels[out_index] = V_NONE;  // leave air
                busy = false;
            }
        }
    }
    my_evals = __reduce_add_sync(0xFFFFFFFFu, my_evals);
    if (lane == 0 && my_evals) atomicAdd(&stats[3], (unsigned long long)my_evals);
}

// ------------------------------------------------------------------------------------------- K2
// K2: areas that contain cave-zone columns (K1 finished all others in shared memory), from the
// list K1 wrote.  Persistent CTAs: one launch wave instead of one (mostly empty) CTA per area.
__global__ void __launch_bounds__(256) gen_finish_kernel(const uint4* __restrict__ jobs,
                                                         const uint32_t* __restrict__ cave_areas,
                                                         const uint32_t* __restrict__ cave_area_count,
                                                         uint8_t* __restrict__ pool_voxels,
                                                         uint8_t* __restrict__ pool_heights,
                                                         uint16_t* __restrict__ pool_planes,
                                                         uint64_t seed) {
    __shared__ uint4 s_vox[VOXELS_IN_AREA / 16];
    __shared__ uint16_t s_tree[COLUMNS_IN_AREA];  // indexed x*16 + y (generator.rs:54-58 order)
    __shared__ int s_min_z;
    const int tid = threadIdx.x;
    const uint32_t n_list = *cave_area_count;
    for (uint32_t item = blockIdx.x; item < n_list; item += gridDim.x) {
    const uint4 job = jobs[cave_areas[item]];
    const uint32_t ax = job.x, ay = job.y, slot = job.z;
    uint4* area = reinterpret_cast<uint4*>(pool_voxels + (size_t)slot * VOXELS_IN_AREA);
    uint8_t* heights = pool_heights + (size_t)slot * COLUMNS_IN_AREA;
    const unsigned x = tid & 15, y = tid >> 4;

    // the carved area comes into shared memory in one sweep; the column walks below would
    // otherwise be chains of dependent HBM reads.  Only the slabs that can differ from Stone: zi <=
    // ALWAYS_STONE_ZI is Stone in every column of every area (the lowest surface is 32, the deepest
    // lake bed 34 - 10; caves start at zi = 25), nothing below reads or changes those rows.
    constexpr int ALWAYS_STONE_ZI = 23;
    static_assert(ALWAYS_STONE_ZI < (int)CAVE_MIN_ZI && ALWAYS_STONE_ZI <= 32 - 4 && ALWAYS_STONE_ZI <= (int)LAKE_TOP_ZI - 10 - 1, "");
    constexpr int ROWS_USED = (AREA_HEIGHT - ALWAYS_STONE_ZI) * 16;
    for (int i = tid; i < ROWS_USED; i += 256) s_vox[i] = area[i];
    if (tid == 0) s_min_z = AREA_HEIGHT;
    __syncthreads();
    uint8_t* vox = reinterpret_cast<uint8_t*>(s_vox);

    // max_generated_height = topmost generated non-None voxel (generator.rs:91-93,100)
    const unsigned H = heights[tid] & 0x7Fu;
    unsigned z = AREA_HEIGHT - H;
    atomicMin(&s_min_z, (int)z);
    uint8_t v = vox[tid + 256u * z];
    while (v == V_NONE && z < AREA_HEIGHT - 1) {
        ++z;
        v = vox[tid + 256u * z];
    }
    const int tree = pick_tree(v, seed, ax, ay, x, y, z);
    s_tree[x * 16 + y] = tree ? (uint16_t)(tree | (z << 8)) : (uint16_t)0;
    __syncthreads();
    if (tid < 32) {
        const int tree_top = place_trees(vox, s_tree, tid);
        if (tid == 0) s_min_z = min(s_min_z, tree_top);  // topmost slab that holds anything
    }
    __syncthreads();

    // Column heights (update_all_column_heights, area.rs:34-56) and bit planes in one walk, as in K1.
    // Everything above s_min_z is None, everything from z_hi on is Stone.
    const int z_lo = s_min_z, z_hi = AREA_HEIGHT - ALWAYS_STONE_ZI;
    uint16_t* planes = pool_planes + (size_t)slot * (VOXELS_IN_AREA / 8);
    {
        const int zt = tid >> 1;  // thread i owns rows 8i .. 8i+7 = half of slab z = i / 2
        if (zt < z_lo || zt >= z_hi) {
            const uint32_t sv = zt < z_lo ? 0u : 0xFFFFFFFFu;
            reinterpret_cast<uint4*>(planes)[tid] = make_uint4(sv, sv, sv, sv);
            reinterpret_cast<uint4*>(planes + VOXELS_IN_AREA / 16)[tid] = make_uint4(~sv, ~sv, ~sv, ~sv);
        }
    }
    const int wid = tid >> 5;
    uint32_t* s_words = reinterpret_cast<uint32_t*>(planes) + wid;
    uint32_t* t_words = reinterpret_cast<uint32_t*>(planes + VOXELS_IN_AREA / 16) + wid;
    int top = z_hi;  // first z from the top that is not transparent
    for (int zz = z_hi - 1; zz >= z_lo; --zz) {
        const unsigned vv = vox[tid + 256 * zz];
        const bool transparent = is_transparent(vv);
        const unsigned sb = __ballot_sync(0xFFFFFFFFu, vv != 0u);
        const unsigned tb = __ballot_sync(0xFFFFFFFFu, transparent);
        if (!transparent) top = zz;
        if ((tid & 31) == 0) {  // rows r = 16 z + 2 wid and r + 1 are one 32-bit word
            s_words[8 * zz] = sb;
            t_words[8 * zz] = tb;
        }
    }
    heights[tid] = (uint8_t)top;
    // trees only touch the slabs from s_min_z down to the surface; write those back
    const int first_row = s_min_z * 16;
    for (int i = first_row + tid; i < ROWS_USED; i += 256) area[i] = s_vox[i];
    __syncthreads();  // shared memory is reused by the next area of this CTA
    }
}

// ------------------------------------------------------------------------------------------- L1
// Area::update_all_column_heights (area.rs:34-56) + the S/T bit planes of an uploaded / decoded area.
// The 32 KiB of voxels are read ONCE, as 16-byte rows (eight independent loads per thread, coalesced);
// the T plane is kept in shared memory, the slabs above the first one that holds anything opaque are
// skipped by every column (one ballot per row), and a column then walks T bits, not voxel bytes.
__global__ void __launch_bounds__(256) column_heights_kernel(const uint4* __restrict__ jobs,
                                                             const uint8_t* __restrict__ pool_voxels,
                                                             uint8_t* __restrict__ pool_heights,
                                                             uint2* __restrict__ slot_xy,
                                                             uint16_t* __restrict__ pool_planes) {
    __shared__ uint16_t s_t[VOXELS_IN_AREA / 16];
    __shared__ int s_first[8];
    const int tid = threadIdx.x, lane = tid & 31;
    const uint4 job = jobs[blockIdx.x];
    const uint32_t slot = job.z;
    const uint4* rows = reinterpret_cast<const uint4*>(pool_voxels + (size_t)slot * VOXELS_IN_AREA);
    if (tid == 0 && slot_xy) slot_xy[slot] = make_uint2(job.x, job.y);
    uint4 q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = rows[tid + 256 * k];  // row r = tid + 256 k: y = tid & 15, z = (tid >> 4) + 16 k
    uint16_t* planes = pool_planes ? pool_planes + (size_t)slot * (VOXELS_IN_AREA / 8) : nullptr;
    int first = AREA_HEIGHT;  // first slab (from the top) of this thread's rows that is not all transparent
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t sb, tb;
        planes_of_row(q[k], sb, tb);
        const int r = tid + 256 * k;
        s_t[r] = (uint16_t)tb;
        if (planes) {
            planes[r] = (uint16_t)sb;
            planes[VOXELS_IN_AREA / 16 + r] = (uint16_t)tb;
        }
        // the 16 rows of a slab sit in one half-warp
        const unsigned clear = __ballot_sync(0xFFFFFFFFu, tb == 0xFFFFu);
        const bool slab_clear = ((clear >> (lane & 16)) & 0xFFFFu) == 0xFFFFu;
        if (!slab_clear) first = min(first, (tid >> 4) + 16 * k);
    }
    first = __reduce_min_sync(0xFFFFFFFFu, first);
    if (lane == 0) s_first[tid >> 5] = first;
    __syncthreads();
    int z = s_first[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) z = min(z, s_first[w]);
    const int x = tid & 15, y = tid >> 4;  // heights are indexed x + 16 y
    while (z < AREA_HEIGHT && ((s_t[16 * z + y] >> x) & 1u)) ++z;
    pool_heights[(size_t)slot * COLUMNS_IN_AREA + tid] = (uint8_t)(z < AREA_HEIGHT ? z : AREA_HEIGHT - 1);
}

// ------------------------------------------------------------------------------------------- cost hint
// Cave-zone columns per area (CaveGenerator::is_cave_zone, cave_generator.rs:36-41) and the voxels of
// those columns that run the cave noise (should_be_cave's range, cave_generator.rs:43-64), without
// generating anything: a cave voxel costs up to 14 3-D octaves against 16-18 2-D octaves for a whole
// ordinary column, so these are the weights by which a region is cut into equal-cost strips for several
// GPUs.  One CTA per area, one thread per column; only cave-zone columns evaluate their terrain height.
__global__ void __launch_bounds__(256) cave_columns_kernel(const __grid_constant__ NoiseTables nt,
                                                           const uint2* __restrict__ area_xy,
                                                           uint32_t* __restrict__ columns,
                                                           uint32_t* __restrict__ voxels) {
    __shared__ __align__(256) NoiseShared sm;
    __shared__ unsigned s_n[2];
    const int tid = threadIdx.x;
    load_noise_shared<FBM_TABLE<FBM_ALT>>(&sm, nt, tid, 256);
    if (tid < 2) s_n[tid] = 0;
    __syncthreads();
    const uint2 a = area_xy[blockIdx.x];
    const double px = (double)((tid & 15) + a.x * AREA_SIZE), py = (double)((tid >> 4) + a.y * AREA_SIZE);
    const int cave_cuts[1] = {80};
    unsigned evals = 0;
    const bool cave_zone = fbm2_region<FBM_CAVE_CHECK>(nt, &sm, px, py, cave_cuts, evals) == 1;
    unsigned cnt = 0;
    if (cave_zone) {  // TerrainTypeGenerator::sample (terrain_type.rs:26-32), as in gen_columns_kernel
        const double n1 = fbm2<FBM_HEIGHT1>(nt, &sm, px, py), n2 = fbm2<FBM_HEIGHT2>(nt, &sm, px, py);
        double value = fmax(n1, n2);
        value = value < 0.1 ? 0.1 : (value > 1.0 ? 1.0 : value);
        const double height_mod = 1.0 - (fbm2<FBM_HEIGHT_MOD>(nt, &sm, px, py) + 1.0) / 2.0;
        const unsigned H = f64_as_u32(height_mod * value * 0.55 * 128.0) + 32u;
        const unsigned top = H < CAVE_MAX_ZI - 1 ? H : CAVE_MAX_ZI - 1;
        cnt = top >= CAVE_MIN_ZI ? top - CAVE_MIN_ZI + 1 : 0;
    }
    const unsigned n = __popc(__ballot_sync(0xFFFFFFFFu, cave_zone));
    const unsigned v = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((tid & 31) == 0 && n) {
        atomicAdd(&s_n[0], n);
        atomicAdd(&s_n[1], v);
    }
    __syncthreads();
    if (tid == 0) {
        columns[blockIdx.x] = s_n[0];
        if (voxels) voxels[blockIdx.x] = s_n[1];
    }
}

// ------------------------------------------------------------------------------------------- diag
// DADD + DMUL issue-rate probe: 8 independent chains per thread, one add and one multiply per
// chain per step, nothing contractable (-fmad=false).  2 * 8 * iters ops per thread.
__global__ void __launch_bounds__(256) fp64_rate_kernel(double* sink, int iters) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.125, a2 = a0 + 0.25, a3 = a0 + 0.375;
    double a4 = a0 + 0.5, a5 = a0 + 0.625, a6 = a0 + 0.75, a7 = a0 + 0.875;
    const double m = 0.9999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m; a1 = a1 * m; a2 = a2 * m; a3 = a3 * m;
        a4 = a4 * m; a5 = a5 * m; a6 = a6 * m; a7 = a7 * m;
        a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c;
        a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c;
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 12345.678) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace

void launch_gen_columns(const NoiseTables& nt, const uint4* jobs, int n, uint8_t* pool_voxels, uint8_t* pool_heights,
                        uint2* slot_xy, CaveItem* cave_items, uint32_t* cave_count,
                        unsigned long long* stats, uint16_t* pool_planes, uint32_t* cave_areas,
                        cudaStream_t stream) {
    if (n <= 0) return;
    gen_columns_kernel<<<n, 256, 0, stream>>>(nt, jobs, pool_voxels, pool_heights, slot_xy, cave_items,
                                              cave_count, stats, pool_planes, cave_areas);
}

void launch_gen_caves(const NoiseTables& nt, const CaveItem* cave_items, const uint32_t* cave_count, uint8_t* pool_voxels,
                      const uint8_t* pool_heights, const uint2* slot_xy, unsigned long long* stats,
                      int sm_count, cudaStream_t stream) {
    gen_caves_kernel<<<sm_count * VX_CAVE_CTAS_PER_SM, VX_CAVE_THREADS, 0, stream>>>(nt, cave_items, cave_count, pool_voxels,
                                                       pool_heights, slot_xy, stats);
}

void launch_gen_finish(const uint4* jobs, const uint32_t* cave_areas, const uint32_t* cave_area_count, int n,
                       uint8_t* pool_voxels, uint8_t* pool_heights, uint16_t* pool_planes, uint64_t seed,
                       int sm_count, cudaStream_t stream) {
    if (n <= 0) return;
    const int grid = n < sm_count * 6 ? n : sm_count * 6;  // 33 KiB of shared memory each: 6 per SM
    gen_finish_kernel<<<grid, 256, 0, stream>>>(jobs, cave_areas, cave_area_count, pool_voxels, pool_heights,
                                                pool_planes, seed);
}

void launch_column_heights(const uint4* jobs, int n, const uint8_t* pool_voxels,
                           uint8_t* pool_heights, uint2* slot_xy, uint16_t* pool_planes,
                           cudaStream_t stream) {
    if (n <= 0) return;
    column_heights_kernel<<<n, 256, 0, stream>>>(jobs, pool_voxels, pool_heights, slot_xy, pool_planes);
}

void launch_cave_columns(const NoiseTables& nt, const uint2* area_xy, int n, uint32_t* columns, uint32_t* voxels,
                         cudaStream_t stream) {
    if (n <= 0) return;
    cave_columns_kernel<<<n, 256, 0, stream>>>(nt, area_xy, columns, voxels);
}

void launch_fp64_rate(double* sink, int blocks, int iters, cudaStream_t stream) {
    fp64_rate_kernel<<<blocks, 256, 0, stream>>>(sink, iters);
}

}  // namespace vx
