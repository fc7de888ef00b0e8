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

// Resident-area lookup: open-addressing hash table (AreaLocation -> pool slot), built on the host
// whenever the resident set changes.  Entry = {ax, ay, slot, used}; capacity is a power of two and
// at most half full, so probes are short.  Replaces World.areas: HashMap<AreaLocation, Area>
// (world.rs:16-21) for device-side neighbour lookups.
struct AreaMap {
    const uint4* entries;
    uint32_t mask;  // capacity - 1; 0 entries when the table is empty
};

__host__ __device__ __forceinline__ uint32_t area_hash(uint32_t ax, uint32_t ay) {
    uint32_t h = ax * 0x9E3779B1u ^ (ay * 0x85EBCA77u);
    h ^= h >> 15;
    h *= 0xC2B2AE3Du;
    h ^= h >> 13;
    return h;
}

__device__ __forceinline__ int area_slot(const AreaMap& m, uint32_t ax, uint32_t ay) {
    if (m.entries == nullptr) return -1;
    uint32_t h = area_hash(ax, ay) & m.mask;
    while (true) {
        const uint4 e = m.entries[h];
        if (e.w == 0u) return -1;
        if (e.x == ax && e.y == ay) return (int)e.z;
        h = (h + 1u) & m.mask;
    }
}

}  // namespace vx
