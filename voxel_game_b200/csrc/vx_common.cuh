// vx_common.cuh -- shared constants and small device helpers of the chunk pipeline.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vx {

constexpr int AREA_SIZE = 16;           // area.rs:7
constexpr int AREA_HEIGHT = 128;        // area.rs:8
constexpr int VOXELS_IN_AREA = 32768;   // area.rs:9
constexpr int COLUMNS_IN_AREA = 256;
constexpr int LOCATION_OFFSET = 1000000;  // location.rs:6

// voxel.rs:8-36
enum Voxel : uint8_t {
    V_NONE = 0, V_COBBLESTONE, V_SAND, V_GRASS, V_WOOD, V_LEAVES, V_BRICK, V_DIRT, V_BOARDS,
    V_STONE, V_CLAY, V_LAMP, V_TRAMPOLINE, V_CACTUS, V_WATER_SOURCE, V_WATER_DOWN, V_WATER1,
    V_WATER2, V_WATER3, V_WATER4, V_STONE_BRICK, V_STONE_PILLAR, V_SNOW, V_ICE, V_BOMB,
    V_ACTIVE_BOMB, V_GLASS, V_VARIANTS
};

// Voxel::TRANSPARENT (voxel.rs:39-49) as a bit set over the discriminants
constexpr uint32_t TRANSPARENT_BITS = (1u << V_NONE) | (1u << V_GLASS) | (1u << V_WATER_SOURCE) |
                                      (1u << V_WATER_DOWN) | (1u << V_WATER1) | (1u << V_WATER2) |
                                      (1u << V_WATER3) | (1u << V_WATER4) | (1u << V_ICE);
static_assert(TRANSPARENT_BITS == 0x048FC001u, "TRANSPARENT set");

__host__ __device__ __forceinline__ bool is_transparent(unsigned v) {
    return (TRANSPARENT_BITS >> v) & 1u;
}

// Resident-area lookup: dense slot map over the bounding box of the loaded areas.
struct AreaMap {
    const int32_t* slots;  // w*h entries, -1 = not resident
    uint32_t ax0, ay0, w, h;
};

__device__ __forceinline__ int area_slot(const AreaMap& m, uint32_t ax, uint32_t ay) {
    const uint32_t ix = ax - m.ax0, iy = ay - m.ay0;  // wraps to huge when below the origin
    if (ix >= m.w || iy >= m.h) return -1;
    return m.slots[ix + m.w * iy];
}

}  // namespace vx
