// vx_tables.cpp -- host-side construction of the per-seed noise tables (integer work, done once
// per vx_set_seed; it is set-up, not a compute fallback: no voxel is ever produced on the CPU).
//
// What is restated here (crates are not vendored in the reference, see DESIGN.md "libnoise"):
//   libnoise 1.1.2  PermutationTable::new(seed, 256, true): identity table shuffled by
//                   rand 0.8.5 SliceRandom::shuffle driven by rand_chacha 0.3.1 ChaCha12Rng
//                   seeded with rand_core 0.6.4 SeedableRng::seed_from_u64.
//   libnoise 1.1.2  Fbm::new normalisation 1 / sum persistence.powi(i)  (powi = compiler-rt
//                   __powidf2 square-and-multiply).
//   Rust std        DefaultHasher (SipHash-1-3, zero keys) over str bytes + 0xFF.
#include "vx_tables.hpp"

#include <array>
#include <cstring>

namespace vx {
namespace {

struct ChaCha12 {
    std::array<uint32_t, 16> state{};
    std::array<uint32_t, 16> block{};
    unsigned cursor = 16;

    static uint32_t rol(uint32_t v, unsigned n) { return (v << n) | (v >> (32u - n)); }

    static void quarter(std::array<uint32_t, 16>& s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rol(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rol(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rol(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rol(s[b] ^ s[c], 7);
    }

    explicit ChaCha12(uint64_t seed_u64) {
        // seed_from_u64: eight PCG32 (XSH-RR) outputs form the 256-bit key
        uint64_t pcg = seed_u64;
        state[0] = 0x61707865u; state[1] = 0x3320646eu; state[2] = 0x79622d32u; state[3] = 0x6b206574u;
        for (int w = 0; w < 8; ++w) {
            pcg = pcg * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xs = static_cast<uint32_t>(((pcg >> 18) ^ pcg) >> 27);
            unsigned rot = static_cast<unsigned>(pcg >> 59);
            state[4 + w] = (xs >> rot) | (xs << ((32u - rot) & 31u));
        }
        state[12] = state[13] = 0;  // 64-bit block counter
        state[14] = state[15] = 0;  // stream id
    }

    void refill() {
        std::array<uint32_t, 16> w = state;
        for (int r = 0; r < 6; ++r) {  // 12 rounds = 6 column+diagonal pairs
            quarter(w, 0, 4, 8, 12); quarter(w, 1, 5, 9, 13);
            quarter(w, 2, 6, 10, 14); quarter(w, 3, 7, 11, 15);
            quarter(w, 0, 5, 10, 15); quarter(w, 1, 6, 11, 12);
            quarter(w, 2, 7, 8, 13); quarter(w, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) block[i] = w[i] + state[i];
        if (++state[12] == 0) ++state[13];
        cursor = 0;
    }

    uint32_t next_u32() {
        if (cursor == 16) refill();
        return block[cursor++];
    }

    // UniformInt<u32>::sample_single(0, bound): widening multiply, reject when the low half
    // exceeds the zone (bound << leading_zeros) - 1
    uint32_t below(uint32_t bound) {
        const uint32_t zone = (bound << __builtin_clz(bound)) - 1u;
        while (true) {
            const uint64_t wide = static_cast<uint64_t>(next_u32()) * bound;
            if (static_cast<uint32_t>(wide) <= zone) return static_cast<uint32_t>(wide >> 32);
        }
    }
};

void shuffled_identity(uint64_t seed, uint8_t out[256]) {
    ChaCha12 rng(seed);
    for (int i = 0; i < 256; ++i) out[i] = static_cast<uint8_t>(i);
    for (uint32_t i = 255; i > 0; --i) {  // Fisher-Yates from the back
        const uint32_t j = rng.below(i + 1);
        const uint8_t tmp = out[i];
        out[i] = out[j];
        out[j] = tmp;
    }
}

double powi(double base, int exp) {  // exp >= 0 here
    double result = 1.0;
    while (true) {
        if (exp & 1) result *= base;
        exp /= 2;
        if (exp == 0) return result;
        base *= base;
    }
}

FbmParams make_fbm(int table, int octaves, double f, double l, double p) {
    FbmParams q{};
    q.frequency = f;
    q.lacunarity = l;
    q.octaves = octaves;
    q.table = table;
    double sum = 0.0, amp = 1.0, freq = f;
    for (int i = 0; i < octaves; ++i) {
        sum = sum + powi(p, i);
        q.amp[i] = amp;
        amp *= p;
        q.freq[i] = freq;  // the running `freq *= lacunarity` of the other Fbm form (VX_NOISE_VARIANT & 1)
        freq *= l;
    }
    q.norm = 1.0 / sum;
    return q;
}

inline uint64_t rol64(uint64_t v, unsigned n) { return (v << n) | (v >> (64u - n)); }

// the device's hi_key (vx_gen.cu): high word of the double, sign-magnitude order made two's-complement order
inline int high_word_key(double x) {
    uint64_t b;
    std::memcpy(&b, &x, sizeof b);
    const int32_t h = (int32_t)(b >> 32);
    return h ^ ((h >> 31) & 0x7FFFFFFF);
}

}  // namespace

void build_noise_tables(uint64_t seed, NoiseTables* out) {
    std::memset(out, 0, sizeof *out);
    out->seed = seed;
    shuffled_identity(seed, out->perm[TAB_SEED]);
    shuffled_identity(~seed, out->perm[TAB_NOT_SEED]);  // seed ^ u64::MAX
    shuffled_identity(seed + 1, out->perm[TAB_SEED_PLUS_1]);
    shuffled_identity(seed + 10, out->perm[TAB_SEED_PLUS_10]);
    out->fbm[FBM_HEIGHT1] = make_fbm(FBM_TABLE<FBM_HEIGHT1>, 6, 0.008, 1.8, 0.45);
    out->fbm[FBM_HEIGHT2] = make_fbm(FBM_TABLE<FBM_HEIGHT2>, 4, 0.004, 1.8, 0.45);
    out->fbm[FBM_HEIGHT_MOD] = make_fbm(FBM_TABLE<FBM_HEIGHT_MOD>, 2, 0.004, 0.004, 0.2);
    out->fbm[FBM_BIOME] = make_fbm(FBM_TABLE<FBM_BIOME>, 2, 0.0015, 1.5, 0.3);
    out->fbm[FBM_CAVE1] = make_fbm(FBM_TABLE<FBM_CAVE1>, 6, 0.08, 1.3, 0.45);
    out->fbm[FBM_CAVE2] = make_fbm(FBM_TABLE<FBM_CAVE2>, 8, 0.04, 1.2, 0.35);
    out->fbm[FBM_CAVE_CHECK] = make_fbm(FBM_TABLE<FBM_CAVE_CHECK>, 2, 0.003, 1.3, 0.3);
    out->fbm[FBM_ALT] = make_fbm(FBM_TABLE<FBM_ALT>, 2, 0.07, 2.0, 0.3);
    out->fbm[FBM_LAKE] = make_fbm(FBM_TABLE<FBM_LAKE>, 2, 0.004, 1.3, 0.35);
    const int cave_ids[2] = {FBM_CAVE1, FBM_CAVE2};
    for (int f = 0; f < 2; ++f) {
        const FbmParams& q = out->fbm[cave_ids[f]];
        for (int k = 0; k <= 8; ++k) {
            double rest = 0.0;
            for (int i = k; i < q.octaves; ++i) rest += q.amp[i];
            out->cave_rest[f][k] = SIMPLEX_BOUND * q.norm * rest;
            const double r = out->cave_rest[f][k] + 1e-9;  // the slack the kernel adds to the bound
            const double t[4] = {(0.4 - r) / q.norm - 1e-9, (0.8 - r) / q.norm - 1e-9, (0.4 + r) / q.norm + 1e-9,
                                 (0.8 + r) / q.norm + 1e-9};
            for (int i = 0; i < 4; ++i) out->cave_thr[f][k][i] = high_word_key(t[i]);
        }
    }
}

// 24 unit gradients at 7.5 deg + k*15 deg (digits as published with OpenSimplex2)
static const double GRAD2[24][2] = {
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

void build_noise_image(const NoiseTables& tables, NoiseImage* out, const uint8_t* order2, const uint8_t* order3) {
    const unsigned n = TAB_COUNT * TAB_STRIDE;
    for (unsigned i = 0; i < n; ++i) {
        const unsigned v = tables.perm[i / TAB_STRIDE][i & 255u];
        out->bytes[i] = (uint8_t)v;
        out->bytes[n + i] = (uint8_t)(v % 24u);
    }
    for (int i = 0; i < 24; ++i) {
        const int src = order2 ? order2[i] % 24 : i;
        out->grad2[i][0] = GRAD2[src][0];
        out->grad2[i][1] = GRAD2[src][1];
    }
    // gradient gi = perm % 12 of the edge-midpoint set: 0..3 (+-1,+-1,0), 4..7 (+-1,0,+-1),
    // 8..11 (0,+-1,+-1); bit 0 negates the first non-zero component, bit 1 the second
    for (int t = 0; t < GRAD3_TABLES; ++t)
        for (unsigned i = 0; i < TAB_STRIDE; ++i) {
            const unsigned hashed = tables.perm[t][i & 255u] % 12u;
            const unsigned gi = order3 ? order3[hashed] % 12u : hashed, grp = gi >> 2;
            out->grad3[t][i].x = (grp == 2u ? 0x7654u : 0x3210u) | ((gi & 1u) << 31);
            out->grad3[t][i].y = (grp == 0u ? 0x3210u : 0x7654u) | (((gi >> 1) & 1u) << 31);
        }
}

uint64_t hash_world_name(const char* name) {
    // SipHash-1-3, k0 = k1 = 0, message = name bytes followed by 0xFF
    uint64_t v0 = 0x736f6d6570736575ull, v1 = 0x646f72616e646f6dull, v2 = 0x6c7967656e657261ull,
             v3 = 0x7465646279746573ull;
    auto round = [&]() {
        v0 += v1; v1 = rol64(v1, 13) ^ v0; v0 = rol64(v0, 32);
        v2 += v3; v3 = rol64(v3, 16) ^ v2;
        v0 += v3; v3 = rol64(v3, 21) ^ v0;
        v2 += v1; v1 = rol64(v1, 17) ^ v2; v2 = rol64(v2, 32);
    };
    const size_t name_len = std::strlen(name);
    const size_t total = name_len + 1;
    auto byte_at = [&](size_t i) -> uint64_t {
        return i < name_len ? static_cast<uint8_t>(name[i]) : 0xFFull;
    };
    size_t pos = 0;
    for (; pos + 8 <= total; pos += 8) {
        uint64_t m = 0;
        for (int b = 0; b < 8; ++b) m |= byte_at(pos + b) << (8 * b);
        v3 ^= m; round(); v0 ^= m;
    }
    uint64_t tail = static_cast<uint64_t>(total & 0xFF) << 56;
    for (int b = 0; pos + b < total; ++b) tail |= byte_at(pos + b) << (8 * b);
    v3 ^= tail; round(); v0 ^= tail;
    v2 ^= 0xFF;
    round(); round(); round();
    return v0 ^ v1 ^ v2 ^ v3;
}

}  // namespace vx
