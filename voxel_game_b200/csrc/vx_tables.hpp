// vx_tables.hpp -- seed-dependent noise tables, built once per seed on the host: the Fbm
// parameters travel to the kernels as a launch parameter, the lookup tables as a 3.4 KiB image in
// device memory that every CTA copies into shared memory.  Replaces the nine Simplex::new(..).fbm(..) constructions that the
// reference repeats for every area (generator.rs:135-146; terrain_type.rs:15-24;
// biome_type.rs:19-23; cave_generator.rs:26-34; voxel_type_generator.rs:24-29;
// lake_generator.rs:21-25).
#pragma once
#include <cstdint>

// The choices of libnoise 1.1.2 that nothing on the build machine pins are compile-time switches, bit
// for bit the flags the repository's CPU checker understands (tests/test_gpu_variants.py);
// vx_noise_variant() reports the mask a library was built with.  0 = the recalled crate behaviour.
//   1  Fbm samples point * freq with a running freq *= lacunarity (default: point *= lacunarity)
//   2  2-D middle corner: x0 == y0 goes to (0,1)                  (default: (1,0))
//   4  3-D traversal: strict >                                    (default: >=)
//   8  2-D skew constants: f64 nearest to the exact values        (default: the 15-digit literals)
// The order of the gradient tables is data: vx_set_gradient_order.
#ifndef VX_NOISE_VARIANT
#define VX_NOISE_VARIANT 0
#endif

namespace vx {

// One Fbm<Simplex> object: which permutation table it reads and its octave schedule.
struct FbmParams {
    double frequency;
    double lacunarity;
    double norm;    // 1 / sum_{i<octaves} persistence^i   (libnoise Fbm::new)
    double amp[8];  // amp[0] = 1, amp[i+1] = amp[i] * persistence  (the running `amp *= p`)
    double freq[8]; // freq[0] = frequency, freq[i+1] = freq[i] * lacunarity (VX_NOISE_VARIANT & 1 only)
    int octaves;
    int table;      // index into NoiseTables::perm
    int pad[2];
};

enum FbmId {
    FBM_HEIGHT1 = 0,  // terrain_type.rs:16  seed        6 oct
    FBM_HEIGHT2,      // terrain_type.rs:17  seed^MAX    4 oct
    FBM_HEIGHT_MOD,   // terrain_type.rs:18  seed+1      2 oct
    FBM_BIOME,        // biome_type.rs:21    seed        2 oct
    FBM_CAVE1,        // cave_generator.rs:27 seed       6 oct (3-D)
    FBM_CAVE2,        // cave_generator.rs:28 seed^MAX   8 oct (3-D)
    FBM_CAVE_CHECK,   // cave_generator.rs:32 seed       2 oct
    FBM_ALT,          // voxel_type_generator.rs:25-27 seed+1 2 oct (3-D)
    FBM_LAKE,         // lake_generator.rs:22 seed+10    2 oct
    FBM_COUNT
};

enum TableId { TAB_SEED = 0, TAB_NOT_SEED, TAB_SEED_PLUS_1, TAB_SEED_PLUS_10, TAB_COUNT };

// Which permutation table an Fbm reads (the seed its Simplex::new got).  Fixed by the reference's
// constructors, so the kernels use it as a compile-time constant: the table base folds into the
// address of every lookup.
template <int ID>
constexpr int FBM_TABLE = ID == FBM_HEIGHT1 || ID == FBM_BIOME || ID == FBM_CAVE1 || ID == FBM_CAVE_CHECK ? TAB_SEED
                          : ID == FBM_HEIGHT2 || ID == FBM_CAVE2                                         ? TAB_NOT_SEED
                          : ID == FBM_HEIGHT_MOD || ID == FBM_ALT                                        ? TAB_SEED_PLUS_1
                                                                                                          : TAB_SEED_PLUS_10;

struct NoiseTables {
    uint64_t seed;
    uint8_t perm[TAB_COUNT][256];  // the shuffled identity tables; NoiseImage holds what the kernels read
    FbmParams fbm[FBM_COUNT];
    // cave_rest[f][k]: upper bound of |norm * sum_{i >= k} amp_i * simplex_i| for the two cave
    // Fbms after k evaluated octaves (SIMPLEX_BOUND * norm * sum of the remaining amplitudes)
    double cave_rest[2][9];
    // cave_thr[f][k] = integer keys (high word of a double, made monotone) of the thresholds that the
    // raw running sum x of cave Fbm f is compared with after k octaves: x below [0] / [1] implies
    // norm * x + rest < 0.4 / 0.8, x above [2] / [3] implies norm * x - rest >= 0.4 / 0.8 (rest =
    // cave_rest[f][k] plus slack).  key(a) < key(b) implies a < b; equal keys decide nothing.
    int cave_thr[2][9][4];
    const void* image;  // device copy of the NoiseImage of this seed (set by the owner of the tables)
};

// What the kernels keep in shared memory, byte for byte: the four permutation tables, doubled to
// 512 entries (entry i = perm[i & 255], the crate's arrangement), their "% 24" copies (the last
// lookup of a 2-D corner only feeds that remainder), the 24 2-D gradients, and for the three
// tables that 3-D noise reads the 3-D gradient CODE of every entry: the last lookup of a 3-D
// corner only feeds "% 12", i.e. one of the 12 edge-midpoint gradients, and g . (x, y, z) for such
// a gradient is +-a +-b with a in {x, y}, b in {y, z}.  A code is two words
//   .x = byte-permute selector of a (0x3210: x, 0x7654: y) | sign of a << 31
//   .y = byte-permute selector of b (0x3210: y, 0x7654: z) | sign of b << 31
// so a corner picks and signs its two terms without a compare.
constexpr unsigned TAB_STRIDE = 512;
constexpr int GRAD3_TABLES = 3;  // TAB_SEED, TAB_NOT_SEED, TAB_SEED_PLUS_1
struct alignas(16) NoiseImage {
    uint8_t bytes[2 * TAB_COUNT * TAB_STRIDE];  // perm (4 x 512) | perm % 24 (4 x 512)
    double grad2[24][2];
    struct Code { uint32_t x, y; } grad3[GRAD3_TABLES][TAB_STRIDE];
};

// |simplex3| <= 1.00004 for this gradient set and normalisation constant (1/76.8808 is the true
// supremum of the raw sum, the crate scales by 76.88375854620836); DESIGN.md "Oracle and parity"
constexpr double SIMPLEX_BOUND = 1.0001;

void build_noise_tables(uint64_t seed, NoiseTables* out);
// order2 / order3: permutations of the 24 2-D / 12 3-D gradients (entry i of the crate's table =
// entry order[i] of the canonical tables in vx_tables.cpp), nullptr = canonical order
void build_noise_image(const NoiseTables& tables, NoiseImage* out, const uint8_t* order2 = nullptr,
                       const uint8_t* order3 = nullptr);
uint64_t hash_world_name(const char* name);

}  // namespace vx
