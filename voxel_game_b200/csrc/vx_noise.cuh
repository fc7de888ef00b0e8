// vx_noise.cuh -- f64 simplex / fBm evaluation on the device.
//
// Follows libnoise 1.1.2 (the crate the reference calls at terrain_type.rs:28-29,
// biome_type.rs:27, cave_generator.rs:38,61, lake_generator.rs:44, voxel_type_generator.rs:186):
// every f64 multiply and add below is a separately rounded IEEE operation in the crate's order
// (the library is compiled with -fmad=false), so results are bit-identical to a scalar CPU
// evaluation.  What is *re-organised* for the GPU, all value-preserving:
//   * floor + integer cast are two FP64 adds (round-down add of 2^52 + 2^51, then subtract)
//     instead of FRND/F2I on the conversion pipe;
//   * the permutation table is doubled (512 B) like the crate's, and the integer lattice
//     coordinates are reduced modulo 256 once per evaluation, so no lookup needs a mask;
//   * the 3-D gradient dot product g.x (components 0/+-1) is one signed add instead of
//     3 multiplies + 2 adds (x*1 = x, x*0 = +-0, y + +-0 = y); which two coordinates and which
//     signs comes from the table as a code (vx_tables.hpp) applied with byte permutes;
//   * Fbm amplitudes amp_k are precomputed on the host with the same running product;
//   * the last permutation lookup of a corner only feeds "% 24" (2-D) or "% 12" (3-D), so it reads
//     a pre-reduced copy of the table, or the gradient code directly;
//   * a corner with t <= 0 contributes 0 in the crate, i.e. leaves the running sum unchanged.  Here
//     the high word of t is clamped with one signed integer max (a negative or -0 t becomes a
//     number below 2^-1022 whose square is exactly +0), so t^4 * d = +-0 is added unconditionally:
//     no FP64 compare and no selects (sum is never -0: it starts at +0.0, and x + (-x) = +0);
//   * "c - (step ? 1.0 : 0.0)" is a select between c - 1.0 and c (c - 0.0 == c bit for bit), and
//     c - 1.0 is shared with the last corner.
// Which table an Fbm reads is a template argument, so the table base is part of each load's
// constant offset; a lookup is one add and one LDS.
#pragma once
#include <cstddef>
#include <cstdint>

#include "vx_tables.hpp"

namespace vx {

// The Fbm parameters travel as a __grid_constant__ kernel parameter and the lookup tables as a
// per-context image in device memory: contexts with different seeds in one process cannot disturb
// each other, which a module-wide __constant__ symbol would allow.  Lookups are random per lane,
// so the tables are read from a shared-memory copy (the constant cache would serialise them); the
// copy is 536 (K1) / 640 (K1b) coalesced 16-byte loads per CTA.
constexpr unsigned TAB_BYTES = TAB_COUNT * TAB_STRIDE;
constexpr unsigned OFF_P24 = TAB_BYTES;      // byte offset of the % 24 copies
// K1's copy: everything 2-D noise reads plus the 3-D gradient codes of ONE table (the surface
// noise's); the leading members mirror NoiseImage
struct alignas(16) NoiseShared {
    uint8_t bytes[2 * TAB_BYTES];
    double grad2[24][2];
    NoiseImage::Code grad3[TAB_STRIDE];
};
static_assert(offsetof(NoiseImage, grad3) == offsetof(NoiseShared, grad3), "image and shared copy share the 2-D part");

// Numeric constants whose low word is not zero cannot be FP64 immediates; as literals ptxas
// re-materialises each of them with two moves per use.  From the constant bank they are plain
// instruction operands.
#if VX_NOISE_VARIANT & 8
#define VX_SKEW_2D 0x1.76cf5d0b09955p-2    // f64 nearest to (sqrt(3) - 1) / 2
#define VX_UNSKEW_2D 0x1.b0cb174df99c7p-3  // f64 nearest to (3 - sqrt(3)) / 6
#else
#define VX_SKEW_2D 0.366025403784439       // the crate's literals as recalled
#define VX_UNSKEW_2D 0.211324865405187
#endif
__constant__ double c_noise_k[8] = {
    VX_SKEW_2D,               // 0: 2-D skew
    VX_UNSKEW_2D,             // 1: 2-D unskew
    2.0 * VX_UNSKEW_2D,       // 2
    99.83685446303647,        // 3: 2-D normalisation
    1.0 / 3.0,                // 4: 3-D skew
    1.0 / 6.0,                // 5: 3-D unskew
    2.0 * (1.0 / 6.0),        // 6
    76.88375854620836,        // 7: 3-D normalisation
};

// `bytes` bytes (a multiple of 16) of the image, from byte `offset` on
__device__ __forceinline__ void load_noise_image(void* dst, const NoiseTables& nt, unsigned offset, unsigned bytes,
                                                 int tid, int nthreads) {
    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(nt.image) + offset);
    for (unsigned i = tid; i < bytes / 16u; i += nthreads) static_cast<uint4*>(dst)[i] = __ldg(src + i);
}

// GRAD3_TABLE = the table whose 3-D gradient codes the kernel needs
template <int GRAD3_TABLE>
__device__ __forceinline__ void load_noise_shared(NoiseShared* s, const NoiseTables& nt, int tid, int nthreads) {
    load_noise_image(s, nt, 0, (unsigned)offsetof(NoiseShared, grad3), tid, nthreads);
    load_noise_image(s->grad3, nt, (unsigned)(offsetof(NoiseImage, grad3) + GRAD3_TABLE * sizeof(s->grad3)),
                     (unsigned)sizeof(s->grad3), tid, nthreads);
}

// floor(x) as a double and as an int, |x| < 2^31, in two FP64 adds and nothing else: adding
// 2^52 + 2^51 with round-toward-minus-infinity leaves exactly floor(x) in the low mantissa bits
// (the ulp of the sum is 1), and subtracting the constant again is exact.
__device__ __forceinline__ double floor_split(double x, int& i) {
    const double MAGIC = 6755399441055744.0;  // 2^52 + 2^51
    const double t = __dadd_rd(x, MAGIC);
    i = __double2loint(t);
    return t - MAGIC;
}

// t for t > 0; for t <= 0 (sign bit set, or +0) a value whose square is exactly +0.0: the high
// word becomes 0, which leaves either +0 or a subnormal below 2^-1022.
__device__ __forceinline__ double clamp_radius(double t) {
    return __hiloint2double(max(__double2hiint(t), 0), __double2loint(t));
}

// contribution of one simplex corner, added to `sum`; gi = gradient index.  A corner outside its
// radius (t <= 0) contributes 0 in the crate, i.e. leaves the running sum unchanged: here it adds
// (+0 * +0) * d = +-0 (d is finite), and sum + +-0 == sum because sum is never -0.
__device__ __forceinline__ void corner2(const NoiseShared* s, double x, double y, unsigned gi, double& sum) {
    const double t = clamp_radius(0.5 - x * x - y * y);
    const double2 g = *reinterpret_cast<const double2*>(&s->grad2[gi][0]);
    const double d = g.x * x + g.y * y;
    const double t2 = t * t;
    sum += t2 * t2 * d;
}

// TABLE = which (doubled) permutation table of s->bytes; its % 24 copy sits OFF_P24 bytes further.
// The table is a template argument so that its base is part of each load's constant offset.
template <int TABLE>
__device__ __forceinline__ double simplex2(const NoiseShared* s, double x, double y) {
    const uint8_t* tb = s->bytes + TABLE * TAB_STRIDE;
    const double SKEW = c_noise_k[0], UNSKEW = c_noise_k[1], UNSKEW_2 = c_noise_k[2];
    const double sk = (x + y) * SKEW;
    int i, j;
    const double fx = floor_split(x + sk, i), fy = floor_split(y + sk, j);
    const double un = (fx + fy) * UNSKEW;
    const double x0 = x - fx + un, y0 = y - fy + un;
#if VX_NOISE_VARIANT & 2
    const bool lower = !(x0 <= y0);  // middle corner (1,0) unless x0 <= y0 (the other tie-break)
#else
    const bool lower = !(x0 < y0);  // middle corner (1,0) unless x0 < y0
#endif
    const double xm = x0 - 1.0, ym = y0 - 1.0;
    const double x1 = (lower ? xm : x0) + UNSKEW, y1 = (lower ? y0 : ym) + UNSKEW;
    const double x2 = xm + UNSKEW_2, y2 = ym + UNSKEW_2;
    const unsigned im = (unsigned)i & 255u, jm = (unsigned)j & 255u;
    const unsigned pi0 = tb[im], pi1 = tb[im + 1];
    const uint8_t* tg = tb + OFF_P24;
    const unsigned h0 = tg[pi0 + jm];
    const unsigned h1 = lower ? tg[pi1 + jm] : tg[pi0 + jm + 1];
    const unsigned h2 = tg[pi1 + jm + 1];
    double sum = 0.0;
    corner2(s, x0, y0, h0, sum);
    corner2(s, x1, y1, h1, sum);
    corner2(s, x2, y2, h2, sum);
    return sum * c_noise_k[3];
}

// prmt.b32 in its default mode reads selector bits 0..15 only (__byte_perm would mask them first)
__device__ __forceinline__ unsigned prmt(int lo, int hi, unsigned selector) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(selector));
    return d;
}

// One corner of the 3-D simplex; e = gradient code of the corner's hash (vx_tables.hpp): the dot
// product with an edge-midpoint gradient is +-a +-b, a in {x, y}, b in {y, z}, picked word by word
// with the code's byte-permute selector and signed with its top bit.
__device__ __forceinline__ void corner3(double x, double y, double z, uint2 e, double& sum) {
    const double t = clamp_radius(0.5 - x * x - y * y - z * z);
    const double a = __hiloint2double((int)(prmt(__double2hiint(x), __double2hiint(y), e.x) ^ (e.x & 0x80000000u)),
                                      (int)prmt(__double2loint(x), __double2loint(y), e.x));
    const double b = __hiloint2double((int)(prmt(__double2hiint(y), __double2hiint(z), e.y) ^ (e.y & 0x80000000u)),
                                      (int)prmt(__double2loint(y), __double2loint(z), e.y));
    const double d = a + b;
    const double t2 = t * t;
    sum += t2 * t2 * d;
}

// One table byte through an explicit shared-memory address: `addr` already holds table image +
// index, the table and the copy are the constant OFF that folds into the load.
template <int OFF>
__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint2 lds_u32x2(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ unsigned smem_address(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// sb = shared-memory address of the permutation tables (256-byte aligned, so that
// "sb | coordinate & 255" is one logic instruction), TABLE = which (doubled) table; sg = address of
// that table's 3-D gradient codes.  Indices are sums of table bytes and coordinates reduced to 8
// bits (at most 255 + 255 + 1), which the doubled tables serve without masking -- the crate's own
// arrangement: one add and one load per lookup.
template <int TABLE>
__device__ __forceinline__ double simplex3(unsigned sb, unsigned sg, double x, double y, double z) {
    constexpr int T = TABLE * (int)TAB_STRIDE;
    const double SKEW = c_noise_k[4], UNSKEW = c_noise_k[5], UNSKEW_2 = c_noise_k[6];
    const double sk = (x + y + z) * SKEW;
    int i, j, k;
    const double fx = floor_split(x + sk, i), fy = floor_split(y + sk, j),
                 fz = floor_split(z + sk, k);
    const double un = (fx + fy + fz) * UNSKEW;
    const double x0 = x - fx + un, y0 = y - fy + un, z0 = z - fz + un;
    // rank the three coordinates: second corner steps along the largest, third corner along the
    // two largest (ties: x before y before z, i.e. x0 >= y0, y0 >= z0, x0 >= z0 count as "greater")
#if VX_NOISE_VARIANT & 4
    const bool xy = x0 > y0, yz = y0 > z0, xz = x0 > z0;  // the other tie-break: ties count as "less"
#else
    const bool xy = x0 >= y0, yz = y0 >= z0, xz = x0 >= z0;
#endif
    const bool i1 = xy && xz, j1 = !xy && yz, k1 = !xz && !yz;
    const bool i2 = xy || xz, j2 = !xy || yz, k2 = !(xz && yz);
    // "c - step" as a select between c - 1.0 and c (c - 0.0 == c); c - 1.0 also serves corner 3
    const double xm = x0 - 1.0, ym = y0 - 1.0, zm = z0 - 1.0;
    const double x1 = (i1 ? xm : x0) + UNSKEW, y1 = (j1 ? ym : y0) + UNSKEW, z1 = (k1 ? zm : z0) + UNSKEW;
    const double x2 = (i2 ? xm : x0) + UNSKEW_2, y2 = (j2 ? ym : y0) + UNSKEW_2, z2 = (k2 ? zm : z0) + UNSKEW_2;
    const double x3 = xm + 0.5, y3 = ym + 0.5, z3 = zm + 0.5;  // 3.0 * UNSKEW == 0.5 exactly
    const unsigned ib = sb | ((unsigned)i & 255u), jb = sb | ((unsigned)j & 255u);
    unsigned kg = sg + ((unsigned)k & 255u) * 8u;  // 8-byte codes
    asm("" : "+r"(kg));                            // kept as one value: "kg + 8 * hash" is one shift-add
    const unsigned pi0 = lds_u8<T>(ib), pi1 = lds_u8<T + 1>(ib);
    // a step along j or k is the load's constant offset, chosen by predicate
    const unsigned a1 = jb + (i1 ? pi1 : pi0), a2 = jb + (i2 ? pi1 : pi0);
    const unsigned q0 = lds_u8<T>(jb + pi0), q1 = j1 ? lds_u8<T + 1>(a1) : lds_u8<T>(a1);
    const unsigned q2 = j2 ? lds_u8<T + 1>(a2) : lds_u8<T>(a2), q3 = lds_u8<T + 1>(jb + pi1);
    const unsigned g1 = kg + q1 * 8u, g2 = kg + q2 * 8u;
    const uint2 h0 = lds_u32x2<0>(kg + q0 * 8u);
    const uint2 h1 = k1 ? lds_u32x2<8>(g1) : lds_u32x2<0>(g1);
    const uint2 h2 = k2 ? lds_u32x2<8>(g2) : lds_u32x2<0>(g2);
    const uint2 h3 = lds_u32x2<8>(kg + q3 * 8u);
    double sum = 0.0;
    corner3(x0, y0, z0, h0, sum);
    corner3(x1, y1, z1, h1, sum);
    corner3(x2, y2, z2, h2, sum);
    corner3(x3, y3, z3, h3, sum);
    return sum * c_noise_k[7];
}

template <int ID>
__device__ __forceinline__ double fbm2(const NoiseTables& nt, const NoiseShared* s, double x, double y) {
    const FbmParams& q = nt.fbm[ID];
    double acc = 0.0;
#if VX_NOISE_VARIANT & 1
#pragma unroll 1
    for (int o = 0; o < q.octaves; ++o) acc += q.amp[o] * simplex2<FBM_TABLE<ID>>(s, x * q.freq[o], y * q.freq[o]);
#else
    x *= q.frequency;
    y *= q.frequency;
#pragma unroll 1
    for (int o = 0; o < q.octaves; ++o) {
        acc += q.amp[o] * simplex2<FBM_TABLE<ID>>(s, x, y);
        x *= q.lacunarity;
        y *= q.lacunarity;
    }
#endif
    return acc * q.norm;
}

template <int ID>
__device__ __forceinline__ double fbm3(const NoiseTables& nt, const NoiseShared* s, double x, double y,
                                       double z) {
    const FbmParams& q = nt.fbm[ID];
    double acc = 0.0;
#if VX_NOISE_VARIANT & 1
#pragma unroll 1
    for (int o = 0; o < q.octaves; ++o)
        acc += q.amp[o] * simplex3<FBM_TABLE<ID>>(smem_address(s->bytes), smem_address(s->grad3), x * q.freq[o],
                                                  y * q.freq[o], z * q.freq[o]);
#else
    x *= q.frequency;
    y *= q.frequency;
    z *= q.frequency;
#pragma unroll 1
    for (int o = 0; o < q.octaves; ++o) {
        acc += q.amp[o] * simplex3<FBM_TABLE<ID>>(smem_address(s->bytes), smem_address(s->grad3), x, y, z);
        x *= q.lacunarity;
        y *= q.lacunarity;
        z *= q.lacunarity;
    }
#endif
    return acc * q.norm;
}

// Three Fbms over the same point with two simplex evaluations in flight per iteration instead of one:
// A (OA octaves) runs beside B (OB < OA octaves), and A's remaining OA - OB octaves run beside C
// (OA - OB octaves).  Every Fbm keeps its own octave order and its own running sum, so all three
// results are the crate's; the pairing only gives the scheduler two independent chains (K1: -1.9 %).
template <int IDA, int IDB, int IDC, int OA, int OB>
__device__ __forceinline__ void fbm2_three(const NoiseTables& nt, const NoiseShared* s, double x, double y, double& ra,
                                           double& rb, double& rc) {
    static_assert(OA > OB, "A is the longest; C has OA - OB octaves");
    const FbmParams& qa = nt.fbm[IDA];
    const FbmParams& qb = nt.fbm[IDB];
    const FbmParams& qc = nt.fbm[IDC];
    double acc_a = 0.0, acc_b = 0.0, acc_c = 0.0;
#if VX_NOISE_VARIANT & 1
#pragma unroll 1
    for (int o = 0; o < OB; ++o) {
        acc_a += qa.amp[o] * simplex2<FBM_TABLE<IDA>>(s, x * qa.freq[o], y * qa.freq[o]);
        acc_b += qb.amp[o] * simplex2<FBM_TABLE<IDB>>(s, x * qb.freq[o], y * qb.freq[o]);
    }
#pragma unroll 1
    for (int o = OB; o < OA; ++o) {
        acc_a += qa.amp[o] * simplex2<FBM_TABLE<IDA>>(s, x * qa.freq[o], y * qa.freq[o]);
        acc_c += qc.amp[o - OB] * simplex2<FBM_TABLE<IDC>>(s, x * qc.freq[o - OB], y * qc.freq[o - OB]);
    }
#else
    double xa = x * qa.frequency, ya = y * qa.frequency, xb = x * qb.frequency, yb = y * qb.frequency;
#pragma unroll 1
    for (int o = 0; o < OB; ++o) {
        acc_a += qa.amp[o] * simplex2<FBM_TABLE<IDA>>(s, xa, ya);
        acc_b += qb.amp[o] * simplex2<FBM_TABLE<IDB>>(s, xb, yb);
        xa *= qa.lacunarity;
        ya *= qa.lacunarity;
        xb *= qb.lacunarity;
        yb *= qb.lacunarity;
    }
    xb = x * qc.frequency;  // B is done: its registers carry C
    yb = y * qc.frequency;
#pragma unroll 1
    for (int o = OB; o < OA; ++o) {
        acc_a += qa.amp[o] * simplex2<FBM_TABLE<IDA>>(s, xa, ya);
        acc_c += qc.amp[o - OB] * simplex2<FBM_TABLE<IDC>>(s, xb, yb);
        xa *= qa.lacunarity;
        ya *= qa.lacunarity;
        xb *= qc.lacunarity;
        yb *= qc.lacunarity;
    }
#endif
    ra = acc_a * qa.norm;
    rb = acc_b * qb.norm;
    rc = acc_c * qc.norm;
}

// Rust `f64 as i32` / `as u32` for the value ranges that occur here (|v| < 2^31, never NaN:
// noise outputs are finite); cvt.rzi saturates like Rust does.
__device__ __forceinline__ int f64_as_i32(double v) { return __double2int_rz(v); }
__device__ __forceinline__ unsigned f64_as_u32(double v) { return __double2uint_rz(v); }

}  // namespace vx
