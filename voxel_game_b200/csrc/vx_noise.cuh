// vx_noise.cuh -- f64 simplex / fBm evaluation on the device.
//
// Follows libnoise 1.1.2 (the crate the reference calls at terrain_type.rs:28-29,
// biome_type.rs:27, cave_generator.rs:38,61, lake_generator.rs:44, voxel_type_generator.rs:186):
// every f64 multiply and add below is a separately rounded IEEE operation in the crate's order
// (the library is compiled with -fmad=false), so results are bit-identical to a scalar CPU
// evaluation.  What is *re-organised* for the GPU, all value-preserving:
//   * floor + integer cast are two FP64 adds (round-down add of 2^52 + 2^51, then subtract)
//     instead of FRND/F2I on the conversion pipe;
//   * the permutation table is kept un-doubled (256 B, at most 2-way bank conflicts) and indexed
//     modulo 256, which equals indexing the crate's doubled 512-entry table;
//   * the 3-D gradient dot product g.x (components 0/+-1) is one signed add instead of
//     3 multiplies + 2 adds (x*1 = x, x*0 = +-0, y + +-0 = y);
//   * Fbm amplitudes amp_k are precomputed on the host with the same running product;
//   * the last permutation lookup only feeds "% 12" / "% 24", so it reads pre-reduced copies of
//     the table; a corner with t <= 0 contributes 0, i.e. leaves the running sum unchanged.
// Tables are addressed as (shared-memory base, byte offset) and never through a selected pointer,
// so every lookup stays an LDS even when the table is chosen per lane.
#pragma once
#include <cstdint>

#include "vx_tables.hpp"

namespace vx {

// The Fbm parameters travel as a __grid_constant__ kernel parameter and the lookup tables as a
// per-context image in device memory: contexts with different seeds in one process cannot disturb
// each other, which a module-wide __constant__ symbol would allow.  Lookups are random per lane,
// so the tables are read from a shared-memory copy (the constant cache would serialise them); the
// copy is 216 coalesced 16-byte loads per CTA.
constexpr unsigned TAB_BYTES = TAB_COUNT * 256;
constexpr unsigned OFF_P24 = TAB_BYTES;      // byte offset of the % 24 copies
constexpr unsigned OFF_P12 = 2 * TAB_BYTES;  // byte offset of the % 12 copies
using NoiseShared = NoiseImage;

// first `bytes` bytes of the image (a multiple of 16): the cave kernel only needs the tables
__device__ __forceinline__ void load_noise_image(void* dst, const NoiseTables& nt, unsigned bytes, int tid,
                                                 int nthreads) {
    const uint4* src = static_cast<const uint4*>(nt.image);
    for (unsigned i = tid; i < bytes / 16u; i += nthreads) static_cast<uint4*>(dst)[i] = __ldg(src + i);
}

__device__ __forceinline__ void load_noise_shared(NoiseShared* s, const NoiseTables& nt, int tid,
                                                  int nthreads) {
    load_noise_image(s, nt, (unsigned)sizeof(NoiseImage), tid, nthreads);
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

// contribution of one simplex corner, added to `sum`; gi = gradient index.  A corner outside its
// radius (t <= 0) contributes 0 in the crate, i.e. leaves the running sum unchanged (sum starts at
// +0.0 and 0.0 + v == v).
__device__ __forceinline__ void corner2(const NoiseShared* s, double x, double y, unsigned gi,
                                        double& sum) {
    const double t = 0.5 - x * x - y * y;
    const double2 g = *reinterpret_cast<const double2*>(&s->grad2[gi][0]);
    const double d = g.x * x + g.y * y;
    const double t2 = t * t;
    const double v = t2 * t2 * d;
    if (t > 0.0) sum += v;
}

// p = byte offset of the permutation table inside s->bytes
__device__ __forceinline__ double simplex2(const NoiseShared* s, unsigned p, double x, double y) {
    const double SKEW = 0.366025403784439, UNSKEW = 0.211324865405187;
    const uint8_t* b = s->bytes;
    const double sk = (x + y) * SKEW;
    int i, j;
    const double fx = floor_split(x + sk, i), fy = floor_split(y + sk, j);
    const double un = (fx + fy) * UNSKEW;
    const double x0 = x - fx + un, y0 = y - fy + un;
    const bool lower = !(x0 < y0);  // middle corner (1,0) unless x0 < y0
    const double x1 = x0 - (lower ? 1.0 : 0.0) + UNSKEW, y1 = y0 - (lower ? 0.0 : 1.0) + UNSKEW;
    const double x2 = x0 - 1.0 + 2.0 * UNSKEW, y2 = y0 - 1.0 + 2.0 * UNSKEW;
    const unsigned pi0 = b[p + (i & 255)], pi1 = b[p + ((i + 1) & 255)];
    const unsigned g = p + OFF_P24;
    const unsigned h0 = b[g + ((pi0 + j) & 255)];
    const unsigned h1 = b[g + (((lower ? pi1 : pi0) + j + (lower ? 0 : 1)) & 255)];
    const unsigned h2 = b[g + ((pi1 + j + 1) & 255)];
    double sum = 0.0;
    corner2(s, x0, y0, h0, sum);
    corner2(s, x1, y1, h1, sum);
    corner2(s, x2, y2, h2, sum);
    return sum * 99.83685446303647;
}

// gradient gi (0..11) of the edge-midpoint set
//   0..3: (+-1,+-1,0)   4..7: (+-1,0,+-1)   8..11: (0,+-1,+-1); bit0 negates the first non-zero
//   component, bit1 the second
__device__ __forceinline__ double flip_sign(double v, unsigned sign_bit31) {
    return __hiloint2double(__double2hiint(v) ^ (int)sign_bit31, __double2loint(v));
}
__device__ __forceinline__ void corner3(double x, double y, double z, unsigned gi, double& sum) {
    const double t = 0.5 - x * x - y * y - z * z;
    const unsigned grp = gi >> 2;
    const double a = flip_sign(grp == 2u ? y : x, gi << 31);
    const double b = flip_sign(grp == 0u ? y : z, (gi << 30) & 0x80000000u);
    const double d = a + b;
    const double t2 = t * t;
    const double v = t2 * t2 * d;
    if (t > 0.0) sum += v;
}

// Table bytes through explicit shared-memory addresses.  With the tables 256-byte aligned,
// "table address | (index & 255)" is one logic instruction and the % 12 copy sits at a fixed
// +2048 that folds into the load: two instructions per lookup.
__device__ __forceinline__ unsigned smem_address(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u8_p12(unsigned addr) {  // same index in the % 12 copy
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+2048];" : "=r"(v) : "r"(addr));
    return v;
}
static_assert(OFF_P12 == 2048, "lds_u8_p12 hard-codes the offset");

template <class Lookup>
__device__ __forceinline__ double simplex3_body(double x, double y, double z, Lookup look);

// b = shared-memory base of the table bytes, p = byte offset of the permutation table in it
__device__ __forceinline__ double simplex3(const uint8_t* b, unsigned p, double x, double y, double z) {
    struct {
        const uint8_t* b;
        unsigned p;
        __device__ __forceinline__ unsigned perm(unsigned idx) const { return b[p + (idx & 255u)]; }
        __device__ __forceinline__ unsigned perm12(unsigned idx) const { return b[p + OFF_P12 + (idx & 255u)]; }
    } look{b, p};
    return simplex3_body(x, y, z, look);
}

// sp = shared-memory address of the permutation table, a multiple of 256
__device__ __forceinline__ double simplex3_sa(unsigned sp, double x, double y, double z) {
    struct {
        unsigned sp;
        __device__ __forceinline__ unsigned perm(unsigned idx) const { return lds_u8(sp | (idx & 255u)); }
        __device__ __forceinline__ unsigned perm12(unsigned idx) const { return lds_u8_p12(sp | (idx & 255u)); }
    } look{sp};
    return simplex3_body(x, y, z, look);
}

template <class Lookup>
__device__ __forceinline__ double simplex3_body(double x, double y, double z, Lookup look) {
    const double SKEW = 1.0 / 3.0, UNSKEW = 1.0 / 6.0;
    const double sk = (x + y + z) * SKEW;
    int i, j, k;
    const double fx = floor_split(x + sk, i), fy = floor_split(y + sk, j),
                 fz = floor_split(z + sk, k);
    const double un = (fx + fy + fz) * UNSKEW;
    const double x0 = x - fx + un, y0 = y - fy + un, z0 = z - fz + un;
    // rank the three coordinates: second corner steps along the largest, third corner along the
    // two largest (ties: x before y before z, i.e. x0 >= y0, y0 >= z0, x0 >= z0 count as "greater")
    const bool xy = x0 >= y0, yz = y0 >= z0, xz = x0 >= z0;
    const bool i1 = xy && xz, j1 = !xy && yz, k1 = !xz && !yz;
    const bool i2 = xy || xz, j2 = !xy || yz, k2 = !(xz && yz);
    // selects, not int->double conversions (those run on the slow conversion pipe)
    const double x1 = x0 - (i1 ? 1.0 : 0.0) + UNSKEW, y1 = y0 - (j1 ? 1.0 : 0.0) + UNSKEW,
                 z1 = z0 - (k1 ? 1.0 : 0.0) + UNSKEW;
    const double x2 = x0 - (i2 ? 1.0 : 0.0) + 2.0 * UNSKEW, y2 = y0 - (j2 ? 1.0 : 0.0) + 2.0 * UNSKEW,
                 z2 = z0 - (k2 ? 1.0 : 0.0) + 2.0 * UNSKEW;
    const double x3 = x0 - 1.0 + 3.0 * UNSKEW, y3 = y0 - 1.0 + 3.0 * UNSKEW,
                 z3 = z0 - 1.0 + 3.0 * UNSKEW;
    const unsigned pi0 = look.perm((unsigned)i), pi1 = look.perm((unsigned)(i + 1));
    const unsigned h0 = look.perm12(look.perm(pi0 + j) + k);
    const unsigned h1 = look.perm12(look.perm((i1 ? pi1 : pi0) + j + (int)j1) + k + (int)k1);
    const unsigned h2 = look.perm12(look.perm((i2 ? pi1 : pi0) + j + (int)j2) + k + (int)k2);
    const unsigned h3 = look.perm12(look.perm(pi1 + j + 1) + k + 1);
    double sum = 0.0;
    corner3(x0, y0, z0, h0, sum);
    corner3(x1, y1, z1, h1, sum);
    corner3(x2, y2, z2, h2, sum);
    corner3(x3, y3, z3, h3, sum);
    return sum * 76.88375854620836;
}

template <int ID>
__device__ __forceinline__ double fbm2(const NoiseTables& nt, const NoiseShared* s, double x, double y) {
    const FbmParams& q = nt.fbm[ID];
    const unsigned p = (unsigned)q.table * 256u;
    x *= q.frequency;
    y *= q.frequency;
    double acc = 0.0;
#pragma unroll 1
    for (int o = 0; o < q.octaves; ++o) {
        acc += q.amp[o] * simplex2(s, p, x, y);
        x *= q.lacunarity;
        y *= q.lacunarity;
    }
    return acc * q.norm;
}

template <int ID>
__device__ __forceinline__ double fbm3(const NoiseTables& nt, const NoiseShared* s, double x, double y,
                                       double z) {
    const FbmParams& q = nt.fbm[ID];
    const unsigned p = (unsigned)q.table * 256u;
    x *= q.frequency;
    y *= q.frequency;
    z *= q.frequency;
    double acc = 0.0;
#pragma unroll 1
    for (int o = 0; o < q.octaves; ++o) {
        acc += q.amp[o] * simplex3(s->bytes, p, x, y, z);
        x *= q.lacunarity;
        y *= q.lacunarity;
        z *= q.lacunarity;
    }
    return acc * q.norm;
}

// Rust `f64 as i32` / `as u32` for the value ranges that occur here (|v| < 2^31, never NaN:
// noise outputs are finite); cvt.rzi saturates like Rust does.
__device__ __forceinline__ int f64_as_i32(double v) { return __double2int_rz(v); }
__device__ __forceinline__ unsigned f64_as_u32(double v) { return __double2uint_rz(v); }

}  // namespace vx
