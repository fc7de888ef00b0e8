/*
 * light_ext.c -- ORACLE (test infrastructure only; see vg_oracle.h).
 *
 * L-ext: voxel light propagation.  THE REFERENCE HAS NO COUNTERPART (lighting is
 * done per fragment in GLSL from the column height map and lamp uniforms), so this is a
 * self-specified extension named by the north star, and its parity is "CUDA vs this flood fill",
 * never "vs the reference".  Specification (DESIGN.md "L-ext"):
 *   - levels 0..15, one byte per voxel, same layout as the voxels;
 *   - a voxel is a light CARRIER iff it is in Voxel::TRANSPARENT (voxel.rs:39-49);
 *   - sources: a carrier above its column's max_height (area.rs:34-56; z < max_height) has
 *     level 15 (sky);  a Lamp voxel has level 15 (emitter, does not carry);
 *   - every other carrier has level max(0, max over its 6 neighbours (level - 1)), where
 *     neighbours outside the loaded grid or outside 0 <= z < 128 contribute nothing;
 *   - every other non-carrier has level 0.
 * This file computes that least fixed point with a classic sequential BFS flood fill.
 */
#include <stdlib.h>
#include <string.h>

#include "vg_oracle.h"

static int carrier(uint8_t v) {
    return v == VGO_NONE || v == VGO_GLASS || v == VGO_ICE ||
           (v >= VGO_WATER_SOURCE && v <= VGO_WATER4);
}

void vgo_light_flood_world(const vgo_world* w, uint8_t* light) {
    const uint32_t W = w->nx * VGO_AREA_SIZE, Hh = w->ny * VGO_AREA_SIZE;
    size_t total = (size_t)W * Hh * VGO_AREA_HEIGHT;
    memset(light, 0, total);
    /* queue of (x, y, z) in grid-local voxel coordinates */
    uint32_t* queue = (uint32_t*)malloc(total * sizeof(uint32_t));
    size_t head = 0, tail = 0;
#define SLOT(x, y) ((size_t)((x) / VGO_AREA_SIZE) + (size_t)w->nx * ((y) / VGO_AREA_SIZE))
#define IDX(x, y, z)                                                        \
    (SLOT(x, y) * VGO_VOXELS_IN_AREA + ((x) % VGO_AREA_SIZE) +              \
     (size_t)((y) % VGO_AREA_SIZE) * VGO_AREA_SIZE + (size_t)(z) * VGO_AREA_SIZE * VGO_AREA_SIZE)
#define PACK(x, y, z) (((uint32_t)(z) << 24) | ((uint32_t)(y) << 12) | (uint32_t)(x))
    for (uint32_t y = 0; y < Hh; y++)
        for (uint32_t x = 0; x < W; x++) {
            uint8_t mh = w->max_height[SLOT(x, y) * VGO_COLUMNS_IN_AREA + (x % VGO_AREA_SIZE) +
                                       (y % VGO_AREA_SIZE) * VGO_AREA_SIZE];
            for (uint32_t z = 0; z < VGO_AREA_HEIGHT; z++) {
                uint8_t v = w->voxels[IDX(x, y, z)];
                if ((carrier(v) && z < mh) || v == VGO_LAMP) {
                    light[IDX(x, y, z)] = 15;
                    queue[tail++] = PACK(x, y, z);
                }
            }
        }
    static const int D[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    while (head < tail) {
        uint32_t p = queue[head++];
        uint32_t x = p & 0xFFF, y = (p >> 12) & 0xFFF, z = p >> 24;
        uint8_t l = light[IDX(x, y, z)];
        if (l <= 1) continue;
        for (int d = 0; d < 6; d++) {
            int nx = (int)x + D[d][0], ny = (int)y + D[d][1], nz = (int)z + D[d][2];
            if (nx < 0 || ny < 0 || nz < 0 || nx >= (int)W || ny >= (int)Hh ||
                nz >= VGO_AREA_HEIGHT)
                continue;
            size_t ni = IDX((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
            if (!carrier(w->voxels[ni])) continue;
            if (light[ni] + 1 < l) { /* BFS in non-increasing level order: first visit is final */
                light[ni] = (uint8_t)(l - 1);
                queue[tail++] = PACK(nx, ny, nz);
            }
        }
    }
    free(queue);
}
