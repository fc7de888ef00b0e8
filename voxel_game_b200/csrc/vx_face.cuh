// vx_face.cuh -- face descriptors and vertex assembly shared by the batch mesher (vx_mesh.cu) and the
// incremental updater (vx_update.cu): MeshGenerator::get_verticies_for_voxel
// (mesh_generator.rs:200-488) as corner codes + a handful of f32 operations.
#pragma once
#include "vx_common.cuh"

namespace vx {

// Corner code per (direction, vertex): bit0 = +x side, bit1 = +y side, bit2 = bottom z (z + 0.5),
// bits 3..4 = uv corner.  Transcribed from the six vec![] blocks of mesh_generator.rs:222-487.
//   Left(x+1):  (x+h,y-h,zb;s1) (x+h,y-h,zt;s2) (x+h,y+h,zt;s3) (x+h,y+h,zb;s0)
//   Right(x-1): (x-h,y-h,zt;s3) (x-h,y-h,zb;s0) (x-h,y+h,zb;s1) (x-h,y+h,zt;s2)
//   Front(y+1): (x-h,y+h,zb;s0) (x+h,y+h,zb;s1) (x+h,y+h,zt;s2) (x-h,y+h,zt;s3)
//   Back(y-1):  (x-h,y-h,zt;s3) (x+h,y-h,zt;s2) (x+h,y-h,zb;s1) (x-h,y-h,zb;s0)
//   Down(z+1):  (x-h,y-h,zb;b0) (x+h,y-h,zb;b1) (x+h,y+h,zb;b2) (x-h,y+h,zb;b3)
//   Up(z-1):    (x+h,y-h,zt;t0) (x-h,y-h,zt;t1) (x-h,y+h,zt;t2) (x+h,y+h,zt;t3)
// The four 5-bit codes of a direction are packed into 20 bits; three directions per 64-bit word,
// so the lookup is two shifts on immediates instead of a (lane-divergent) memory access.
#define CORNER(px, py, bottom, uv) ((unsigned long long)((px) | ((py) << 1) | ((bottom) << 2) | ((uv) << 3)))
#define CORNER4(a, b, c, d) ((a) | ((b) << 5) | ((c) << 10) | ((d) << 15))
constexpr unsigned long long CORNERS_LRF =
    CORNER4(CORNER(1, 0, 1, 1), CORNER(1, 0, 0, 2), CORNER(1, 1, 0, 3), CORNER(1, 1, 1, 0)) |
    (CORNER4(CORNER(0, 0, 0, 3), CORNER(0, 0, 1, 0), CORNER(0, 1, 1, 1), CORNER(0, 1, 0, 2)) << 20) |
    (CORNER4(CORNER(0, 1, 1, 0), CORNER(1, 1, 1, 1), CORNER(1, 1, 0, 2), CORNER(0, 1, 0, 3)) << 40);
constexpr unsigned long long CORNERS_BDU =
    CORNER4(CORNER(0, 0, 0, 3), CORNER(1, 0, 0, 2), CORNER(1, 0, 1, 1), CORNER(0, 0, 1, 0)) |
    (CORNER4(CORNER(0, 0, 1, 0), CORNER(1, 0, 1, 1), CORNER(1, 1, 1, 2), CORNER(0, 1, 1, 3)) << 20) |
    (CORNER4(CORNER(1, 0, 0, 0), CORNER(0, 0, 0, 1), CORNER(0, 1, 0, 2), CORNER(1, 1, 0, 3)) << 40);

// TextureManager::VOXELS_WITH_DIFFERENT_FACES (texture_manager.rs:95-102)
constexpr uint32_t DIFFERENT_FACES_BITS = (1u << V_GRASS) | (1u << V_TRAMPOLINE) | (1u << V_WOOD) |
                                          (1u << V_STONE_PILLAR) | (1u << V_BOMB) |
                                          (1u << V_ACTIVE_BOMB);

// face descriptor: row (11) | x (4) << 11 | dir (3) << 15 | ordinal in voxel (3) << 18 | voxel << 21
__device__ __forceinline__ uint32_t make_desc(int r, int x, int dir, int ord, uint32_t voxel) {
    return (uint32_t)r | ((uint32_t)x << 11) | ((uint32_t)dir << 15) | ((uint32_t)ord << 18) |
           (voxel << 21);
}

// What the four vertices of a face are assembled from.  MeshGenerator::get_verticies_for_voxel
// (mesh_generator.rs:200-488); all f32 operations are the reference's: offset -/+ 0.5,
// "offset_z - 0.5 + height_offset", "side_uv.y -= height_offset".
struct FaceParts {
    float xm, xp, ym, yp, zt, zb;  // the two candidate values of each coordinate
    float vhi, vlo;                // v of uv corners {0,1} and {2,3}
    uint32_t fnx, fny, fnz;        // normal (bit patterns)
    uint32_t codes;                // four 5-bit corner codes (CORNERS_*)
};

__device__ __forceinline__ FaceParts decode_face(uint32_t desc, uint32_t gx0, uint32_t gy0) {
    FaceParts p;
    const int rr = desc & 2047, x = (desc >> 11) & 15, dir = (desc >> 15) & 7;
    const uint32_t voxel = desc >> 21;
    // location.rs:19-27: (internal - 1_000_000) as f32; mesh_generator.rs:120-123
    const float cx = (float)((int)(gx0 + x) - LOCATION_OFFSET);
    const float cy = (float)((int)(gy0 + (rr & 15)) - LOCATION_OFFSET);
    const float cz = (float)(rr >> 4);
    // get_partial_height_offset (mesh_generator.rs:490-498): k / 5 in f32.  A correctly rounded
    // f32 division k.0 / 5.0 is the f32 nearest to k/5, i.e. the literal below: no divide needed.
    const float ho = voxel == V_WATER1 ? 0.2f : voxel == V_WATER2 ? 0.4f : voxel == V_WATER3 ? 0.6f
                   : voxel == V_WATER4 ? 0.8f : 0.0f;
    p.xm = cx - 0.5f; p.xp = cx + 0.5f; p.ym = cy - 0.5f; p.yp = cy + 0.5f;
    p.zt = cz - 0.5f + ho; p.zb = cz + 0.5f;
    // UV_REPEATING 1/0; TOP_UV 0.625/0.375, SIDE_UV 1/0.75, BOTTOM_UV 0.25/0 for voxels with
    // different faces (mesh_generator.rs:61-99)
    float vhi = 1.0f, vlo = 0.0f;
    if ((DIFFERENT_FACES_BITS >> voxel) & 1u) {
        vhi = dir == 5 ? 0.625f : (dir == 4 ? 0.25f : 1.0f);
        vlo = dir == 5 ? 0.375f : (dir == 4 ? 0.0f : 0.75f);
    }
    if (dir < 4 && ho > 0.0f) {  // mesh_generator.rs:214-220, side faces only
        if (vhi > 0.0f) vhi -= ho;
        if (vlo > 0.0f) vlo -= ho;
    }
    p.vhi = vhi; p.vlo = vlo;
    // normals (mesh_generator.rs:45-50)
    p.fnx = __float_as_uint(dir == 0 ? 1.0f : (dir == 1 ? -1.0f : 0.0f));
    p.fny = __float_as_uint(dir == 2 ? 1.0f : (dir == 3 ? -1.0f : 0.0f));
    p.fnz = __float_as_uint(dir == 4 ? 1.0f : (dir == 5 ? -1.0f : 0.0f));
    const unsigned long long packed = dir < 3 ? CORNERS_LRF : CORNERS_BDU;
    p.codes = (uint32_t)(packed >> (20 * (dir < 3 ? dir : dir - 3))) & 0xFFFFFu;
    return p;
}

// Two vertices (2*half, 2*half + 1) of a face = 20 words = 5 uint4 written to the staging tile.
__device__ __forceinline__ void build_half_face(uint4* dst, const FaceParts& p, int half) {
    const uint32_t codes = p.codes >> (10 * half);
    const uint32_t c0 = codes & 31u, c1 = (codes >> 5) & 31u;
    const uint32_t uv0 = c0 >> 3, uv1 = c1 >> 3;
    const uint32_t COLOR = 0xFFFFFFFFu;  // mesh_generator.rs:43
    dst[0] = make_uint4(__float_as_uint((c0 & 1u) ? p.xp : p.xm), __float_as_uint((c0 & 2u) ? p.yp : p.ym),
                        __float_as_uint((c0 & 4u) ? p.zb : p.zt),
                        __float_as_uint((uv0 == 1u || uv0 == 2u) ? 1.0f : 0.0f));
    dst[1] = make_uint4(__float_as_uint(uv0 < 2u ? p.vhi : p.vlo), COLOR, p.fnx, p.fny);
    dst[2] = make_uint4(p.fnz, 0u, __float_as_uint((c1 & 1u) ? p.xp : p.xm),
                        __float_as_uint((c1 & 2u) ? p.yp : p.ym));
    dst[3] = make_uint4(__float_as_uint((c1 & 4u) ? p.zb : p.zt),
                        __float_as_uint((uv1 == 1u || uv1 == 2u) ? 1.0f : 0.0f),
                        __float_as_uint(uv1 < 2u ? p.vhi : p.vlo), COLOR);
    dst[4] = make_uint4(p.fnx, p.fny, p.fnz, 0u);
}


__device__ __forceinline__ bool partial_height(uint32_t v) { return v >= V_WATER1 && v <= V_WATER4; }

}  // namespace vx
