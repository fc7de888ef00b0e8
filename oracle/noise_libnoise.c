/*
 * noise_libnoise.c -- ORACLE (test infrastructure only; see vg_oracle.h).
 *
 * Restatement of the third-party arithmetic the reference calls but does not vendor:
 *   libnoise 1.1.2   (Cargo.lock:281-289)  Simplex<2>, Simplex<3>, Fbm, Max, PermutationTable
 *   rand 0.8.5       (Cargo.lock:483-492)  SliceRandom::shuffle, UniformInt<u32>::sample_single
 *   rand_chacha 0.3.1(Cargo.lock:494-501)  ChaCha12Rng
 *   rand_core 0.6.4  (Cargo.lock:503-505)  SeedableRng::seed_from_u64 (PCG32 key expansion)
 *   Rust std         DefaultHasher = SipHash-1-3, keys (0,0)
 * Reference call sites: src/service/area_generation/terrain_type.rs:16-18,28-29;
 * biome_type.rs:21,27; cave_generator.rs:27-32,38,61; lake_generator.rs:22,44;
 * voxel_type_generator.rs:27,186; generator.rs:24-28.
 *
 * *** PARITY UNPINNED ***  None of those crates' sources exist on this machine and no reference
 * test pins a noise value, so this file restates the PUBLISHED algorithms:
 *   anchored   : ChaCha (IETF quarter-round, 12 rounds, 64-bit counter) -- checked against the
 *                RFC 7539 block function with 20 rounds in tests; SipHash -- checked against the
 *                SipHash-2-4 reference vectors with (c,d) = (2,4); PCG32 seed expansion and
 *                Lemire-style widening-multiply rejection sampling as published in rand 0.8.5;
 *                2-D gradient set: the crate's 2-D normalisation constant 99.83685446303647 equals
 *                1/sup|noise| for the 24-direction OpenSimplex2 gradient set to 15 digits
 *                (tests/test_oracle_noise.py::test_simplex2_normalisation_pins_gradient_set);
 *                3-D: 12 edge-midpoint gradients with r^2 = 0.5 give 1/sup = 76.8808 vs the
 *                crate's 76.88375854620836 (agreement 4e-5, not exact).
 *   unanchored : order of entries inside the gradient tables, tie-breaking of the simplex
 *                traversal, literal digits of the skew constants, exact expression grouping.
 * Everything downstream (block IDs) is therefore "bit-exact vs this restatement", not vs a
 * Rust build.  The file is deliberately self-contained so it can be diffed against the crate.
 *
 * The unanchored choices are SWITCHES (vgo_set_noise_variant, flags VGO_VAR_* in vg_oracle.h; the
 * CUDA side has the same switches as the compile-time mask VX_NOISE_VARIANT and
 * vx_set_gradient_order): Fbm coordinate form, tie-breaks of the 2-D / 3-D simplex traversal,
 * digits of the 2-D skew constants, order of the gradient tables.  tests/test_reference_dump.py
 * loads a dump written by the reference itself (rust/parity_dump/) when one is present and reports
 * which setting reproduces it.
 *
 * Built with -DVGO_COUNT_OPS the f64 additions and multiplications of the simplex / Fbm arithmetic
 * go through counting wrappers (the "counting scalar" of SURVEY.md 8d): vgo_op_counts() then returns
 * the f64 add and multiply counts of the crate form, which tests/test_op_counts.py pins next to the
 * counts the kernels issue (bench.py F2_OPS / F3_OPS).
 */
#include <math.h>
#include <string.h>

#include "vg_oracle.h"

/* ------------------------------------------------------------------ counting scalar (-DVGO_COUNT_OPS) */
#ifdef VGO_COUNT_OPS
static _Thread_local uint64_t t_adds, t_muls;
static inline double op_add(double a, double b) { ++t_adds; return a + b; }
static inline double op_sub(double a, double b) { ++t_adds; return a - b; }
static inline double op_mul(double a, double b) { ++t_muls; return a * b; }
void vgo_op_counts(uint64_t* adds, uint64_t* muls, int reset) {
    *adds = t_adds;
    *muls = t_muls;
    if (reset) t_adds = t_muls = 0;
}
#else
#define op_add(a, b) ((a) + (b))
#define op_sub(a, b) ((a) - (b))
#define op_mul(a, b) ((a) * (b))
void vgo_op_counts(uint64_t* adds, uint64_t* muls, int reset) {
    (void)reset;
    *adds = *muls = 0; /* not a counting build */
}
#endif

/* ------------------------------------------------------------------ variant switches */
static unsigned g_variant = 0;
static int g_any = 0; /* any switch set */
static int g_order_set = 0;
static uint8_t g_order2[24], g_order3[12];
static double g_skew2 = 0.366025403784439, g_unskew2 = 0.211324865405187; /* the crate's literals as recalled */

/* flags: VGO_VAR_*; order2 / order3: gradient-table permutations (entry i of the crate's table =
 * entry order[i] of the tables below), NULL = as listed.  Not thread-safe: set before sampling. */
void vgo_set_noise_variant(unsigned flags, const uint8_t* order2, const uint8_t* order3) {
    g_variant = flags;
    g_any = flags != 0 || order2 != NULL || order3 != NULL;
    g_order_set = (order2 != NULL) || (order3 != NULL);
    for (int i = 0; i < 24; i++) g_order2[i] = order2 ? order2[i] : (uint8_t)i;
    for (int i = 0; i < 12; i++) g_order3[i] = order3 ? order3[i] : (uint8_t)i;
    if (flags & VGO_VAR_SKEW_EXACT) { /* f64 nearest to (sqrt(3) - 1) / 2 and (3 - sqrt(3)) / 6 */
        g_skew2 = 0x1.76cf5d0b09955p-2;
        g_unskew2 = 0x1.b0cb174df99c7p-3;
    } else {
        g_skew2 = 0.366025403784439;
        g_unskew2 = 0.211324865405187;
    }
}
unsigned vgo_get_noise_variant(void) { return g_variant; }

/* ------------------------------------------------------------------ SipHash-c-d */
static inline uint64_t rotl64(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }

#define SIPROUND(v0, v1, v2, v3) \
    do {                         \
        v0 += v1;                \
        v1 = rotl64(v1, 13);     \
        v1 ^= v0;                \
        v0 = rotl64(v0, 32);     \
        v2 += v3;                \
        v3 = rotl64(v3, 16);     \
        v3 ^= v2;                \
        v0 += v3;                \
        v3 = rotl64(v3, 21);     \
        v3 ^= v0;                \
        v2 += v1;                \
        v1 = rotl64(v1, 17);     \
        v1 ^= v2;                \
        v2 = rotl64(v2, 32);     \
    } while (0)

uint64_t vgo_siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t* data,
                     size_t len) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL;
    uint64_t v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL;
    uint64_t v3 = k1 ^ 0x7465646279746573ULL;
    size_t full = len & ~(size_t)7;
    for (size_t i = 0; i < full; i += 8) {
        uint64_t m = 0;
        for (int b = 0; b < 8; b++) m |= (uint64_t)data[i + b] << (8 * b);
        v3 ^= m;
        for (int r = 0; r < c_rounds; r++) SIPROUND(v0, v1, v2, v3);
        v0 ^= m;
    }
    uint64_t b = (uint64_t)len << 56;
    for (size_t i = full; i < len; i++) b |= (uint64_t)data[i] << (8 * (i - full));
    v3 ^= b;
    for (int r = 0; r < c_rounds; r++) SIPROUND(v0, v1, v2, v3);
    v0 ^= b;
    v2 ^= 0xff;
    for (int r = 0; r < d_rounds; r++) SIPROUND(v0, v1, v2, v3);
    return v0 ^ v1 ^ v2 ^ v3;
}

/* generator.rs:24-28: DefaultHasher::new(); world_name.hash(&mut hasher) feeds the UTF-8 bytes
 * followed by one 0xFF byte (impl Hash for str); finish(). */
uint64_t vgo_hash_world_name(const char* world_name) {
    uint8_t buf[256];
    size_t n = strlen(world_name);
    if (n > 254) n = 254;
    memcpy(buf, world_name, n);
    buf[n] = 0xFF;
    return vgo_siphash(1, 3, 0, 0, buf, n + 1);
}

/* ------------------------------------------------------------------ ChaCha12Rng::seed_from_u64 */
static inline uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }
static inline uint32_t rotr32(uint32_t v, int r) { return (v >> r) | (v << ((32 - r) & 31)); }

#define QR(a, b, c, d)     \
    do {                   \
        a += b;            \
        d ^= a;            \
        d = rotl32(d, 16); \
        c += d;            \
        b ^= c;            \
        b = rotl32(b, 12); \
        a += b;            \
        d ^= a;            \
        d = rotl32(d, 8);  \
        c += d;            \
        b ^= c;            \
        b = rotl32(b, 7);  \
    } while (0)

static void chacha_block(const uint32_t key[8], uint64_t counter, int double_rounds,
                         uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    for (int i = 0; i < double_rounds; i++) {
        QR(x[0], x[4], x[8], x[12]);
        QR(x[1], x[5], x[9], x[13]);
        QR(x[2], x[6], x[10], x[14]);
        QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]);
        QR(x[1], x[6], x[11], x[12]);
        QR(x[2], x[7], x[8], x[13]);
        QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

/* rand_core 0.6.4 SeedableRng::seed_from_u64: PCG32 stream fills the 32-byte key */
static void seed_from_u64(uint64_t state, uint32_t key[8]) {
    const uint64_t MUL = 6364136223846793005ULL;
    const uint64_t INC = 11634580027462260723ULL;
    for (int i = 0; i < 8; i++) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        key[i] = rotr32(xorshifted, (int)rot); /* to_le_bytes -> key word i (little endian) */
    }
}

typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[16];
    int idx;
} chacha12_rng;

static void rng_init(chacha12_rng* r, uint64_t seed) {
    seed_from_u64(seed, r->key);
    r->counter = 0;
    r->idx = 16;
}

/* BlockRng::next_u32: keystream words in order (rand_chacha refills 4 blocks at a time, which
 * yields the same word sequence as block-by-block generation). */
static uint32_t rng_next_u32(chacha12_rng* r) {
    if (r->idx >= 16) {
        chacha_block(r->key, r->counter++, 6, r->buf);
        r->idx = 0;
    }
    return r->buf[r->idx++];
}

void vgo_chacha12_words(uint64_t seed_u64, uint32_t* out, int n_words) {
    chacha12_rng r;
    rng_init(&r, seed_u64);
    for (int i = 0; i < n_words; i++) out[i] = rng_next_u32(&r);
}

/* exposed for the RFC 7539 anchor test: 20-round block with explicit key and 64-bit counter */
void vgo_chacha_block_raw(const uint32_t key[8], uint64_t counter, int double_rounds,
                          uint32_t out[16]) {
    chacha_block(key, counter, double_rounds, out);
}

/* rand 0.8.5 UniformInt<u32>::sample_single(0, ubound): widening multiply + zone rejection */
static uint32_t gen_index(chacha12_rng* r, uint32_t ubound) {
    uint32_t range = ubound; /* high - low, low = 0 */
    uint32_t zone = (range << __builtin_clz(range)) - 1u;
    for (;;) {
        uint32_t v = rng_next_u32(r);
        uint64_t m = (uint64_t)v * (uint64_t)range;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo <= zone) return hi;
    }
}

/* libnoise PermutationTable::new(seed, 256, true): table = 0..256; shuffle with
 * ChaCha12Rng::seed_from_u64(seed) (rand SliceRandom::shuffle: for i in (1..len).rev()
 * swap(i, gen_index(rng, i + 1))); then the table is appended to itself. */
void vgo_simplex_new(vgo_simplex* s, uint64_t seed) {
    chacha12_rng r;
    rng_init(&r, seed);
    for (int i = 0; i < 256; i++) s->perm[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; i--) {
        uint32_t j = gen_index(&r, i + 1);
        uint8_t t = s->perm[i];
        s->perm[i] = s->perm[j];
        s->perm[j] = t;
    }
    memcpy(s->perm + 256, s->perm, 256);
}

/* ------------------------------------------------------------------ simplex noise */
/* V = 1: the switches are consulted; V = 0 (nothing is switched, the usual case and the one the CPU
 * baseline is timed with): compile-time constants, no table indirection -- the body is compiled twice */
#define SKEW_2D (V ? g_skew2 : 0.366025403784439)     /* (sqrt(3)-1)/2 */
#define UNSKEW_2D (V ? g_unskew2 : 0.211324865405187) /* (3-sqrt(3))/6 */
#define ALWAYS_INLINE static inline __attribute__((always_inline))
static const double SKEW_3D = 1.0 / 3.0;
static const double UNSKEW_3D = 1.0 / 6.0;
static const double R_SQUARED = 0.5;
static const double NORM_2D = 99.83685446303647;
static const double NORM_3D = 76.88375854620836;

/* 24 unit vectors at 7.5 deg + k * 15 deg (the OpenSimplex2 2-D gradient set, in its order) */
#define GRAD2_N 24
static const double GRAD2[GRAD2_N][2] = {
    {0.130526192220052, 0.99144486137381},    {0.38268343236509, 0.923879532511287},
    {0.608761429008721, 0.793353340291235},   {0.793353340291235, 0.608761429008721},
    {0.923879532511287, 0.38268343236509},    {0.99144486137381, 0.130526192220051},
    {0.99144486137381, -0.130526192220051},   {0.923879532511287, -0.38268343236509},
    {0.793353340291235, -0.60876142900872},   {0.608761429008721, -0.793353340291235},
    {0.38268343236509, -0.923879532511287},   {0.130526192220052, -0.99144486137381},
    {-0.130526192220052, -0.99144486137381},  {-0.38268343236509, -0.923879532511287},
    {-0.608761429008721, -0.793353340291235}, {-0.793353340291235, -0.60876142900872},
    {-0.923879532511287, -0.38268343236509},  {-0.99144486137381, -0.130526192220052},
    {-0.99144486137381, 0.130526192220051},   {-0.923879532511287, 0.38268343236509},
    {-0.793353340291235, 0.608761429008721},  {-0.608761429008721, 0.793353340291235},
    {-0.38268343236509, 0.923879532511287},   {-0.130526192220052, 0.99144486137381}};

/* 12 edge-midpoint gradients */
#define GRAD3_N 12
static const double GRAD3[GRAD3_N][3] = {
    {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1},  {-1, 0, 1},
    {1, 0, -1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}};

ALWAYS_INLINE double contribution2(double x, double y, unsigned gi, const int V) {
    double t = op_sub(op_sub(R_SQUARED, op_mul(x, x)), op_mul(y, y));
    if (t <= 0.0) return 0.0;
    const double* g = GRAD2[(V && g_order_set) ? g_order2[gi] : gi];
    t = op_mul(t, t);
    return op_mul(op_mul(t, t), op_add(op_mul(g[0], x), op_mul(g[1], y)));
}

ALWAYS_INLINE double contribution3(double x, double y, double z, unsigned gi, const int V) {
    double t = op_sub(op_sub(op_sub(R_SQUARED, op_mul(x, x)), op_mul(y, y)), op_mul(z, z));
    if (t <= 0.0) return 0.0;
    const double* g = GRAD3[(V && g_order_set) ? g_order3[gi] : gi];
    t = op_mul(t, t);
    return op_mul(op_mul(t, t), op_add(op_add(op_mul(g[0], x), op_mul(g[1], y)), op_mul(g[2], z)));
}

ALWAYS_INLINE double simplex2_impl(const vgo_simplex* s, double x, double y, const int V) {
    const uint8_t* p = s->perm;
    /* transform into lattice space and floor for the cell origin */
    double skew = op_mul(op_add(x, y), SKEW_2D);
    double ix = floor(op_add(x, skew)), iy = floor(op_add(y, skew));
    /* input point relative to the unskewed cell origin */
    double unskew = op_mul(op_add(ix, iy), UNSKEW_2D);
    double x0 = op_add(op_sub(x, ix), unskew), y0 = op_add(op_sub(y, iy), unskew);
    /* middle corner */
    int i1 = 1, j1 = 0;
    if ((V && (g_variant & VGO_VAR_TIE2_OTHER)) ? (x0 <= y0) : (x0 < y0)) {
        i1 = 0;
        j1 = 1;
    }
    double x1 = op_add(op_sub(x0, (double)i1), UNSKEW_2D), y1 = op_add(op_sub(y0, (double)j1), UNSKEW_2D);
    double x2 = op_add(op_sub(x0, 1.0), (2.0 * UNSKEW_2D)), y2 = op_add(op_sub(y0, 1.0), (2.0 * UNSKEW_2D));
    /* hashed gradient indices (rem_euclid 256; doubled table makes +1 safe) */
    unsigned i = (unsigned)((long long)ix & 255), j = (unsigned)((long long)iy & 255);
    unsigned gi0 = p[j + p[i]] % GRAD2_N;
    unsigned gi1 = p[j + j1 + p[i + i1]] % GRAD2_N;
    unsigned gi2 = p[j + 1 + p[i + 1]] % GRAD2_N;
    double n0 = contribution2(x0, y0, gi0, V);
    double n1 = contribution2(x1, y1, gi1, V);
    double n2 = contribution2(x2, y2, gi2, V);
    return op_mul(op_add(op_add(n0, n1), n2), NORM_2D);
}

ALWAYS_INLINE double simplex3_impl(const vgo_simplex* s, double x, double y, double z, const int V) {
    const uint8_t* p = s->perm;
    double skew = op_mul(op_add(op_add(x, y), z), SKEW_3D);
    double ix = floor(op_add(x, skew)), iy = floor(op_add(y, skew)), iz = floor(op_add(z, skew));
    double unskew = op_mul(op_add(op_add(ix, iy), iz), UNSKEW_3D);
    double x0 = op_add(op_sub(x, ix), unskew), y0 = op_add(op_sub(y, iy), unskew), z0 = op_add(op_sub(z, iz), unskew);
    /* simplex traversal: second and third corner offsets; ties count as "greater or equal" in the
     * order x, y, z (VGO_VAR_TIE3_OTHER: as "less") */
    const int strict = V && (g_variant & VGO_VAR_TIE3_OTHER) != 0;
    const int xy = strict ? x0 > y0 : x0 >= y0, yz = strict ? y0 > z0 : y0 >= z0, xz = strict ? x0 > z0 : x0 >= z0;
    int i1, j1, k1, i2, j2, k2;
    if (xy) {
        if (yz) {
            i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0;
        } else if (xz) {
            i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1;
        } else {
            i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1;
        }
    } else {
        if (!yz) {
            i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1;
        } else if (!xz) {
            i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1;
        } else {
            i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0;
        }
    }
    double x1 = op_add(op_sub(x0, (double)i1), UNSKEW_3D), y1 = op_add(op_sub(y0, (double)j1), UNSKEW_3D),
           z1 = op_add(op_sub(z0, (double)k1), UNSKEW_3D);
    double x2 = op_add(op_sub(x0, (double)i2), (2.0 * UNSKEW_3D)), y2 = op_add(op_sub(y0, (double)j2), (2.0 * UNSKEW_3D)),
           z2 = op_add(op_sub(z0, (double)k2), (2.0 * UNSKEW_3D));
    double x3 = op_add(op_sub(x0, 1.0), (3.0 * UNSKEW_3D)), y3 = op_add(op_sub(y0, 1.0), (3.0 * UNSKEW_3D)),
           z3 = op_add(op_sub(z0, 1.0), (3.0 * UNSKEW_3D));
    unsigned i = (unsigned)((long long)ix & 255), j = (unsigned)((long long)iy & 255),
             k = (unsigned)((long long)iz & 255);
    unsigned gi0 = p[k + p[j + p[i]]] % GRAD3_N;
    unsigned gi1 = p[k + k1 + p[j + j1 + p[i + i1]]] % GRAD3_N;
    unsigned gi2 = p[k + k2 + p[j + j2 + p[i + i2]]] % GRAD3_N;
    unsigned gi3 = p[k + 1 + p[j + 1 + p[i + 1]]] % GRAD3_N;
    double n0 = contribution3(x0, y0, z0, gi0, V);
    double n1 = contribution3(x1, y1, z1, gi1, V);
    double n2 = contribution3(x2, y2, z2, gi2, V);
    double n3 = contribution3(x3, y3, z3, gi3, V);
    return op_mul(op_add(op_add(op_add(n0, n1), n2), n3), NORM_3D);
}

double vgo_simplex2(const vgo_simplex* s, double x, double y) {
    return g_any ? simplex2_impl(s, x, y, 1) : simplex2_impl(s, x, y, 0);
}
double vgo_simplex3(const vgo_simplex* s, double x, double y, double z) {
    return g_any ? simplex3_impl(s, x, y, z, 1) : simplex3_impl(s, x, y, z, 0);
}

/* ------------------------------------------------------------------ Fbm */
/* compiler-rt __powidf2, which Rust's f64::powi lowers to */
static double powi(double a, int b) {
    const int recip = b < 0;
    double r = 1;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1 / r : r;
}

void vgo_fbm_new(vgo_fbm* f, const vgo_simplex* src, int octaves, double frequency,
                 double lacunarity, double persistence) {
    f->src = src;
    f->octaves = octaves;
    f->frequency = frequency;
    f->lacunarity = lacunarity;
    f->persistence = persistence;
    double acc = 0.0;
    for (int o = 0; o < octaves; o++) acc = acc + powi(persistence, o);
    f->normalization_factor = 1.0 / acc;
}

/* Fbm::sample.  Default: the point is scaled by the frequency once and by the lacunarity after
 * every octave.  VGO_VAR_FBM_RUNNING_FREQ: every octave samples point * freq with a running
 * freq *= lacunarity (the two differ in the last bits: (p * f) * l vs p * (f * l)). */
double vgo_fbm2(const vgo_fbm* f, double x, double y) {
    double noise = 0.0, amp = 1.0;
    if (g_variant & VGO_VAR_FBM_RUNNING_FREQ) {
        double freq = f->frequency;
        for (int o = 0; o < f->octaves; o++) {
            noise = op_add(noise, op_mul(amp, vgo_simplex2(f->src, op_mul(x, freq), op_mul(y, freq))));
            freq = op_mul(freq, f->lacunarity);
            amp = op_mul(amp, f->persistence);
        }
        return op_mul(noise, f->normalization_factor);
    }
    x = op_mul(x, f->frequency);
    y = op_mul(y, f->frequency);
    if (!g_any) {
        for (int o = 0; o < f->octaves; o++) {
            noise = op_add(noise, op_mul(amp, simplex2_impl(f->src, x, y, 0)));
            x = op_mul(x, f->lacunarity);
            y = op_mul(y, f->lacunarity);
            amp = op_mul(amp, f->persistence);
        }
        return op_mul(noise, f->normalization_factor);
    }
    for (int o = 0; o < f->octaves; o++) {
        noise = op_add(noise, op_mul(amp, vgo_simplex2(f->src, x, y)));
        x = op_mul(x, f->lacunarity);
        y = op_mul(y, f->lacunarity);
        amp = op_mul(amp, f->persistence);
    }
    return op_mul(noise, f->normalization_factor);
}

double vgo_fbm3(const vgo_fbm* f, double x, double y, double z) {
    double noise = 0.0, amp = 1.0;
    if (g_variant & VGO_VAR_FBM_RUNNING_FREQ) {
        double freq = f->frequency;
        for (int o = 0; o < f->octaves; o++) {
            noise = op_add(noise, op_mul(amp, vgo_simplex3(f->src, op_mul(x, freq), op_mul(y, freq), op_mul(z, freq))));
            freq = op_mul(freq, f->lacunarity);
            amp = op_mul(amp, f->persistence);
        }
        return op_mul(noise, f->normalization_factor);
    }
    x = op_mul(x, f->frequency);
    y = op_mul(y, f->frequency);
    z = op_mul(z, f->frequency);
    if (!g_any) {
        for (int o = 0; o < f->octaves; o++) {
            noise = op_add(noise, op_mul(amp, simplex3_impl(f->src, x, y, z, 0)));
            x = op_mul(x, f->lacunarity);
            y = op_mul(y, f->lacunarity);
            z = op_mul(z, f->lacunarity);
            amp = op_mul(amp, f->persistence);
        }
        return op_mul(noise, f->normalization_factor);
    }
    for (int o = 0; o < f->octaves; o++) {
        noise = op_add(noise, op_mul(amp, vgo_simplex3(f->src, x, y, z)));
        x = op_mul(x, f->lacunarity);
        y = op_mul(y, f->lacunarity);
        z = op_mul(z, f->lacunarity);
        amp = op_mul(amp, f->persistence);
    }
    return op_mul(noise, f->normalization_factor);
}

/* gradient tables, exposed so that tests can pin them against the crate's normalisation constants */
void vgo_grad2_table(double* out48) { memcpy(out48, GRAD2, sizeof GRAD2); }
void vgo_grad3_table(double* out36) { memcpy(out36, GRAD3, sizeof GRAD3); }
void vgo_simplex_perm(const vgo_simplex* s, uint8_t* out512) { memcpy(out512, s->perm, 512); }
