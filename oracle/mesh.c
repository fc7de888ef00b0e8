/*
 * mesh.c -- ORACLE (test infrastructure only; see vg_oracle.h).
 * Face culling, vertex emission, the edit path and the height-map image, following
 * src/model/world.rs, src/model/area.rs, src/graphics/renderer.rs,
 * src/graphics/mesh_generator.rs and src/graphics/height_map.rs of the reference.
 * Written voxel-at-a-time with explicit neighbour lookups, the way the reference does it.
 */
#include <string.h>

#include "vg_oracle.h"

/* voxel.rs:39-49 */
static int transparent(uint8_t v) {
    switch (v) {
    case VGO_NONE:
    case VGO_GLASS:
    case VGO_WATER_SOURCE:
    case VGO_WATER_DOWN:
    case VGO_WATER1:
    case VGO_WATER2:
    case VGO_WATER3:
    case VGO_WATER4:
    case VGO_ICE: return 1;
    default: return 0;
    }
}
/* voxel.rs:71 */
static int partial_height(uint8_t v) { return v >= VGO_WATER1 && v <= VGO_WATER4; }

/* mesh_generator.rs:501-513, the three-clause form kept literally */
int vgo_should_generate_face(uint8_t cur, uint8_t nbr) {
    if (cur == nbr) return 0;
    return nbr == VGO_NONE || (!transparent(cur) && transparent(nbr)) ||
           (transparent(cur) && transparent(nbr) && cur != nbr);
}
/* mesh_generator.rs:515-518 */
int vgo_should_generate_top_face(uint8_t cur, uint8_t nbr) {
    return vgo_should_generate_face(cur, nbr) || partial_height(cur);
}

static inline size_t vidx(uint32_t x, uint32_t y, uint32_t z) {
    return (size_t)x + (size_t)y * VGO_AREA_SIZE + (size_t)z * VGO_AREA_SIZE * VGO_AREA_SIZE;
}

/* grid slot of an area, or -1 when the area is not loaded */
static long area_slot(const vgo_world* w, uint32_t ax, uint32_t ay) {
    if (ax < w->ax0 || ay < w->ay0) return -1;
    uint32_t ix = ax - w->ax0, iy = ay - w->ay0;
    if (ix >= w->nx || iy >= w->ny) return -1;
    return (long)ix + (long)w->nx * (long)iy;
}

/* World::get_without_loading (world.rs:176-182): -1 when the area is not loaded */
static int world_get(const vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz) {
    long slot = area_slot(w, gx / VGO_AREA_SIZE, gy / VGO_AREA_SIZE);
    if (slot < 0) return -1;
    return w->voxels[(size_t)slot * VGO_VOXELS_IN_AREA +
                     vidx(gx % VGO_AREA_SIZE, gy % VGO_AREA_SIZE, gz)];
}

/* Area::has_non_transparent_neighbours (area.rs:126-201) */
static int has_non_transparent_neighbours(const uint8_t* a, uint32_t x, uint32_t y, uint32_t z) {
    if (x == 0 || x + 1 >= VGO_AREA_SIZE || y == 0 || y + 1 >= VGO_AREA_SIZE || z == 0 ||
        z + 1 >= VGO_AREA_HEIGHT)
        return 0;
    uint8_t cur = a[vidx(x, y, z)];
    if (vgo_should_generate_face(cur, a[vidx(x + 1, y, z)])) return 0;
    if (vgo_should_generate_face(cur, a[vidx(x - 1, y, z)])) return 0;
    if (vgo_should_generate_face(cur, a[vidx(x, y + 1, z)])) return 0;
    if (vgo_should_generate_face(cur, a[vidx(x, y - 1, z)])) return 0;
    if (vgo_should_generate_face(cur, a[vidx(x, y, z + 1)])) return 0;
    if (vgo_should_generate_face(cur, a[vidx(x, y, z - 1)])) return 0; /* plain face on purpose */
    return 1;
}

/* Renderer::generate_mesh_for_voxel (renderer.rs:126-195): face directions in push order.
 * Returns -1 if a needed neighbour area is not loaded (the reference would generate it). */
static int face_mask_for_voxel(const vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz,
                               uint8_t voxel) {
    if (voxel == VGO_NONE) return 0;
    int mask = 0, n;
    if ((n = world_get(w, gx + 1, gy, gz)) < 0) return -1;
    if (vgo_should_generate_face(voxel, (uint8_t)n)) mask |= VGO_FACE_LEFT;
    if ((n = world_get(w, gx - 1, gy, gz)) < 0) return -1;
    if (vgo_should_generate_face(voxel, (uint8_t)n)) mask |= VGO_FACE_RIGHT;
    if ((n = world_get(w, gx, gy + 1, gz)) < 0) return -1;
    if (vgo_should_generate_face(voxel, (uint8_t)n)) mask |= VGO_FACE_FRONT;
    if ((n = world_get(w, gx, gy - 1, gz)) < 0) return -1;
    if (vgo_should_generate_face(voxel, (uint8_t)n)) mask |= VGO_FACE_BACK;
    if (gz + 1 < VGO_AREA_HEIGHT &&
        vgo_should_generate_face(voxel, (uint8_t)world_get(w, gx, gy, gz + 1)))
        mask |= VGO_FACE_DOWN;
    if (gz > 0 && vgo_should_generate_top_face(voxel, (uint8_t)world_get(w, gx, gy, gz - 1)))
        mask |= VGO_FACE_UP;
    return mask;
}

long vgo_mesh_area(vgo_world* w, uint32_t ax, uint32_t ay) {
    long slot = area_slot(w, ax, ay);
    if (slot < 0) return -1;
    const uint8_t* a = w->voxels + (size_t)slot * VGO_VOXELS_IN_AREA;
    uint8_t* fm = w->face_mask + (size_t)slot * VGO_VOXELS_IN_AREA;
    memset(fm, 0, VGO_VOXELS_IN_AREA);
    long faces = 0;
    /* world.rs:129-143: z, y, x; skip None; skip enclosed voxels */
    for (uint32_t z = 0; z < VGO_AREA_HEIGHT; z++)
        for (uint32_t y = 0; y < VGO_AREA_SIZE; y++)
            for (uint32_t x = 0; x < VGO_AREA_SIZE; x++) {
                uint8_t v = a[vidx(x, y, z)];
                if (v == VGO_NONE) continue;
                if (has_non_transparent_neighbours(a, x, y, z)) continue;
                int m = face_mask_for_voxel(w, x + ax * VGO_AREA_SIZE, y + ay * VGO_AREA_SIZE, z, v);
                if (m < 0) return -1;
                fm[vidx(x, y, z)] = (uint8_t)m;
                faces += __builtin_popcount((unsigned)m);
            }
    return faces;
}

/* ------------------------------------------------------------------ vertex emission */
typedef struct {
    float px, py, pz;
    float u, v;
    uint8_t color[4];
    float nx, ny, nz, nw;
} vertex; /* macroquad::ui::Vertex, repr(C), glam scalar-math: 40 bytes */

/* mesh_generator.rs:490-498 */
static float partial_height_offset(uint8_t voxel) {
    switch (voxel) {
    case VGO_WATER1: return 1.0f / 5.0f;
    case VGO_WATER2: return 2.0f / 5.0f;
    case VGO_WATER3: return 3.0f / 5.0f;
    case VGO_WATER4: return 4.0f / 5.0f;
    default: return 0.0f;
    }
}

/* texture_manager.rs:95-102 */
static int has_different_faces(uint8_t v) {
    return v == VGO_GRASS || v == VGO_TRAMPOLINE || v == VGO_WOOD || v == VGO_STONE_PILLAR ||
           v == VGO_BOMB || v == VGO_ACTIVE_BOMB;
}

static vertex mk(float x, float y, float z, const float uv[2], const float n[4]) {
    vertex r;
    r.px = x; r.py = y; r.pz = z;
    r.u = uv[0]; r.v = uv[1];
    r.color[0] = r.color[1] = r.color[2] = r.color[3] = 255; /* mesh_generator.rs:43 */
    r.nx = n[0]; r.ny = n[1]; r.nz = n[2]; r.nw = n[3];
    return r;
}

/* MeshGenerator::get_verticies_for_voxel (mesh_generator.rs:200-488); dir is one VGO_FACE_* bit */
static void vertices_for_face(uint8_t voxel, int dir, float ox, float oy, float oz, vertex out[4]) {
    const float H = 0.5f; /* voxel.rs:72-73 */
    /* mesh_generator.rs:60-100, f32 constant arithmetic restated */
    const float GAP = 32.0f, TEX = 64.0f;
    const float TEX_H = TEX * 3.0f + GAP * 2.0f;
    const float PX = 1.0f / TEX_H;
    const float O1 = TEX * PX;
    const float O2 = TEX * PX + GAP * PX;
    const float O3 = 2.0f * TEX * PX + GAP * PX;
    const float O4 = 2.0f * TEX * PX + 2.0f * GAP * PX;
    float top[4][2], side[4][2], bottom[4][2];
    const float rep[4][2] = {{0.0f, 1.0f}, {1.0f, 1.0f}, {1.0f, 0.0f}, {0.0f, 0.0f}};
    if (has_different_faces(voxel)) {
        const float t[4][2] = {{0.0f, O3}, {1.0f, O3}, {1.0f, O2}, {0.0f, O2}};
        const float s[4][2] = {{0.0f, 1.0f}, {1.0f, 1.0f}, {1.0f, O4}, {0.0f, O4}};
        const float b[4][2] = {{0.0f, O1}, {1.0f, O1}, {1.0f, 0.0f}, {0.0f, 0.0f}};
        memcpy(top, t, sizeof top);
        memcpy(side, s, sizeof side);
        memcpy(bottom, b, sizeof bottom);
    } else {
        memcpy(top, rep, sizeof top);
        memcpy(side, rep, sizeof side);
        memcpy(bottom, rep, sizeof bottom);
    }
    float ho = partial_height_offset(voxel);
    if (ho > 0.0f)
        for (int i = 0; i < 4; i++)
            if (side[i][1] > 0.0f) side[i][1] -= ho;

    static const float FRONT_N[4] = {0.0f, 1.0f, 0.0f, 0.0f}, BACK_N[4] = {0.0f, -1.0f, 0.0f, 0.0f},
                       RIGHT_N[4] = {-1.0f, 0.0f, 0.0f, 0.0f}, LEFT_N[4] = {1.0f, 0.0f, 0.0f, 0.0f},
                       DOWN_N[4] = {0.0f, 0.0f, 1.0f, 0.0f}, UP_N[4] = {0.0f, 0.0f, -1.0f, 0.0f};
    switch (dir) {
    case VGO_FACE_FRONT:
        out[0] = mk(ox - H, oy + H, oz + H, side[0], FRONT_N);
        out[1] = mk(ox + H, oy + H, oz + H, side[1], FRONT_N);
        out[2] = mk(ox + H, oy + H, oz - H + ho, side[2], FRONT_N);
        out[3] = mk(ox - H, oy + H, oz - H + ho, side[3], FRONT_N);
        break;
    case VGO_FACE_BACK:
        out[0] = mk(ox - H, oy - H, oz - H + ho, side[3], BACK_N);
        out[1] = mk(ox + H, oy - H, oz - H + ho, side[2], BACK_N);
        out[2] = mk(ox + H, oy - H, oz + H, side[1], BACK_N);
        out[3] = mk(ox - H, oy - H, oz + H, side[0], BACK_N);
        break;
    case VGO_FACE_RIGHT:
        out[0] = mk(ox - H, oy - H, oz - H + ho, side[3], RIGHT_N);
        out[1] = mk(ox - H, oy - H, oz + H, side[0], RIGHT_N);
        out[2] = mk(ox - H, oy + H, oz + H, side[1], RIGHT_N);
        out[3] = mk(ox - H, oy + H, oz - H + ho, side[2], RIGHT_N);
        break;
    case VGO_FACE_LEFT:
        out[0] = mk(ox + H, oy - H, oz + H, side[1], LEFT_N);
        out[1] = mk(ox + H, oy - H, oz - H + ho, side[2], LEFT_N);
        out[2] = mk(ox + H, oy + H, oz - H + ho, side[3], LEFT_N);
        out[3] = mk(ox + H, oy + H, oz + H, side[0], LEFT_N);
        break;
    case VGO_FACE_DOWN:
        out[0] = mk(ox - H, oy - H, oz + H, bottom[0], DOWN_N);
        out[1] = mk(ox + H, oy - H, oz + H, bottom[1], DOWN_N);
        out[2] = mk(ox + H, oy + H, oz + H, bottom[2], DOWN_N);
        out[3] = mk(ox - H, oy + H, oz + H, bottom[3], DOWN_N);
        break;
    default: /* VGO_FACE_UP */
        out[0] = mk(ox + H, oy - H, oz - H + ho, top[0], UP_N);
        out[1] = mk(ox - H, oy - H, oz - H + ho, top[1], UP_N);
        out[2] = mk(ox - H, oy + H, oz - H + ho, top[2], UP_N);
        out[3] = mk(ox + H, oy + H, oz - H + ho, top[3], UP_N);
        break;
    }
}

long vgo_emit_area(const vgo_world* w, uint32_t ax, uint32_t ay, uint32_t first_face_base,
                   vgo_voxel_rec* recs, long* n_recs, uint8_t* vertices, uint16_t* indices) {
    long slot = area_slot(w, ax, ay);
    if (slot < 0) return -1;
    const uint8_t* a = w->voxels + (size_t)slot * VGO_VOXELS_IN_AREA;
    const uint8_t* fm = w->face_mask + (size_t)slot * VGO_VOXELS_IN_AREA;
    long faces = 0, nr = 0;
    static const uint16_t INDECIES[6] = {0, 1, 2, 0, 2, 3}; /* mesh_generator.rs:44 */
    for (uint32_t z = 0; z < VGO_AREA_HEIGHT; z++)
        for (uint32_t y = 0; y < VGO_AREA_SIZE; y++)
            for (uint32_t x = 0; x < VGO_AREA_SIZE; x++) {
                unsigned m = fm[vidx(x, y, z)];
                if (!m) continue;
                uint8_t voxel = a[vidx(x, y, z)];
                uint32_t gx = x + ax * VGO_AREA_SIZE, gy = y + ay * VGO_AREA_SIZE;
                /* mesh_generator.rs:120-123 via location.rs:19-27 */
                float mx = (float)((int32_t)gx - VGO_LOCATION_OFFSET);
                float my = (float)((int32_t)gy - VGO_LOCATION_OFFSET);
                float mz = (float)(int32_t)z;
                unsigned nf = (unsigned)__builtin_popcount(m);
                if (recs) {
                    recs[nr].gx = gx;
                    recs[nr].gy = gy;
                    recs[nr].packed = z | ((uint32_t)voxel << 8) | (m << 16) | (nf << 24);
                    recs[nr].first_face = first_face_base + (uint32_t)faces;
                }
                nr++;
                uint16_t index_offset = 0;
                for (int bit = 0; bit < 6; bit++) {
                    if (!(m & (1u << bit))) continue;
                    if (vertices) {
                        vertex fv[4];
                        vertices_for_face(voxel, 1 << bit, mx, my, mz, fv);
                        memcpy(vertices + (size_t)faces * 4 * sizeof(vertex), fv, sizeof fv);
                    }
                    if (indices)
                        for (int i = 0; i < 6; i++)
                            indices[(size_t)faces * 6 + i] = (uint16_t)(INDECIES[i] + index_offset);
                    index_offset = (uint16_t)(index_offset + 4);
                    faces++;
                }
            }
    if (n_recs) *n_recs = nr;
    return faces;
}

/* ------------------------------------------------------------------ edits */
/* Renderer::update_meshes_for_voxel (renderer.rs:221-237) with cached_area = None */
static void update_meshes_for_voxel(vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz) {
    long slot = area_slot(w, gx / VGO_AREA_SIZE, gy / VGO_AREA_SIZE);
    if (slot < 0) return; /* world.get_without_loading -> None: skipped (renderer.rs:241,281) */
    size_t i = (size_t)slot * VGO_VOXELS_IN_AREA + vidx(gx % VGO_AREA_SIZE, gy % VGO_AREA_SIZE, gz);
    int m = face_mask_for_voxel(w, gx, gy, gz, w->voxels[i]);
    if (m < 0) return; /* neighbour area unloaded: the reference would generate it; callers keep
                          edits inside the ring so this cannot happen in tests */
    w->face_mask[i] = (uint8_t)m; /* set_voxel_mesh: insert or remove (renderer.rs:197-219) */
}

/* Renderer::update_location (renderer.rs:239-286): the voxel, then x+1, x-1, y+1, y-1, z-1, z+1 */
void vgo_update_location(vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz) {
    update_meshes_for_voxel(w, gx, gy, gz);
    update_meshes_for_voxel(w, gx + 1, gy, gz);
    update_meshes_for_voxel(w, gx - 1, gy, gz);
    update_meshes_for_voxel(w, gx, gy + 1, gz);
    update_meshes_for_voxel(w, gx, gy - 1, gz);
    if (gz > 0) update_meshes_for_voxel(w, gx, gy, gz - 1);
    if (gz < VGO_AREA_HEIGHT - 1) update_meshes_for_voxel(w, gx, gy, gz + 1);
}

int vgo_apply_edit(vgo_world* w, uint32_t gx, uint32_t gy, uint32_t gz, uint8_t voxel) {
    long slot = area_slot(w, gx / VGO_AREA_SIZE, gy / VGO_AREA_SIZE);
    if (slot < 0 || gz >= VGO_AREA_HEIGHT || voxel >= VGO_VOXEL_VARIANTS) return -1;
    /* World::set (world.rs:184-191) -> Area::set (area.rs:99-111) */
    vgo_area_set(w->voxels + (size_t)slot * VGO_VOXELS_IN_AREA,
                 w->max_height + (size_t)slot * VGO_COLUMNS_IN_AREA, gx % VGO_AREA_SIZE,
                 gy % VGO_AREA_SIZE, gz, voxel);
    vgo_update_location(w, gx, gy, gz);
    return 0;
}

/* ------------------------------------------------------------------ height map */
/* HeightMap::generate_height_map (height_map.rs:41-85) + filter_distant_areas (:90-101) */
void vgo_height_map(const uint32_t* area_xy, int n, const uint8_t* max_height, uint32_t cam_area_x,
                    uint32_t cam_area_y, uint8_t* pixels) {
    const int IMAGE_SIZE = 512, MAX_DISTANCE = 16;
    for (int a = 0; a < n; a++) {
        int dx = (int)area_xy[2 * a] - (int)cam_area_x, dy = (int)area_xy[2 * a + 1] - (int)cam_area_y;
        if (!(dx > -MAX_DISTANCE && dx < MAX_DISTANCE && dy > -MAX_DISTANCE && dy < MAX_DISTANCE))
            continue;
        for (int y = 0; y < VGO_AREA_SIZE; y++)
            for (int x = 0; x < VGO_AREA_SIZE; x++) {
                uint8_t h = max_height[(size_t)a * VGO_COLUMNS_IN_AREA + x + VGO_AREA_SIZE * y];
                int image_x = IMAGE_SIZE / 2 + dx * VGO_AREA_SIZE + x;
                int image_y = IMAGE_SIZE / 2 + dy * VGO_AREA_SIZE + y;
                uint8_t* p = pixels + 4 * ((size_t)image_x + (size_t)image_y * IMAGE_SIZE);
                p[0] = p[1] = p[2] = p[3] = h;
            }
    }
}

/* ------------------------------------------------------------------ batch (CPU baseline timing)
 * Renderer::load_all_blocking (renderer.rs:468-472) over a list of areas: face masks + full vertex
 * and index emission per area (into scratch that is thrown away, like a Mesh that is built and
 * stored).  threads == 1 is the faithful setting (the reference meshes serially on the main
 * thread); threads != 1 runs areas in parallel with OpenMP ("generous"). */
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
long vgo_mesh_areas(vgo_world* w, const uint32_t* area_xy, int n, int threads, long* total_recs) {
    long faces_total = 0, recs_total = 0;
    int failed = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : faces_total, recs_total)
#endif
    for (int i = 0; i < n; i++) {
        long nf = vgo_mesh_area(w, area_xy[2 * i], area_xy[2 * i + 1]);
        if (nf < 0) {
            failed = 1;
            continue;
        }
        uint8_t* verts = (uint8_t*)malloc((size_t)(nf ? nf : 1) * 4 * sizeof(vertex));
        uint16_t* idx = (uint16_t*)malloc((size_t)(nf ? nf : 1) * 6 * sizeof(uint16_t));
        vgo_voxel_rec* recs = (vgo_voxel_rec*)malloc((size_t)(nf ? nf : 1) * sizeof(vgo_voxel_rec));
        long nr = 0;
        vgo_emit_area(w, area_xy[2 * i], area_xy[2 * i + 1], 0, recs, &nr, verts, idx);
        faces_total += nf;
        recs_total += nr;
        free(verts);
        free(idx);
        free(recs);
    }
    if (total_recs) *total_recs = recs_total;
    return failed ? -1 : faces_total;
}
