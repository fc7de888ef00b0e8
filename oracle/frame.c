/*
 * frame.c -- CPU ORACLE (test infrastructure, see vg_oracle.h) for the per-frame visibility pass:
 * which meshed areas and which of their voxel meshes the reference draws this frame.
 *
 * Follows src/graphics/renderer.rs:324-361 (is_area_visible, optimise_render_order), :365-462
 * (set_voxel_shader_and_find_visible_areas, render_voxels), :496-586 (prepare_lights,
 * prepare_visible_areas, filter_visible_voxels, calculate_area_render_distance, is_voxel_visible).
 *
 * PARITY UNPINNED at the glam boundary: Vec3 is glam 0.27.0's scalar Vec3 (through macroquad
 * 0.4.14), not vendored under /root/reference, and the reference has no test for these
 * functions.  The f32 operations below are glam's published ones, each rounded separately
 * (the oracle is built with -ffp-contract=off):
 *   dot(a, b)            (a.x * b.x) + (a.y * b.y) + (a.z * b.z)
 *   length(v)            sqrt(dot(v, v))
 *   distance(a, b)       length(a - b)
 *   normalize_or_zero(v) rcp = 1.0 / length(v); rcp finite and > 0 ? v * rcp : 0
 */
#include <math.h>
#include <string.h>

#include "vg_oracle.h"

/* renderer.rs:36-39 */
#define AREA_RENDER_THRESHOLD 0.35f
#define LOOK_DOWN_RENDER_MULTIPLIER 0.5f
#define VOXEL_RENDER_THRESHOLD 0.71f
#define VOXEL_PROXIMITY_THRESHOLD 5.5f

static float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static float length3(const float v[3]) { return sqrtf(dot3(v, v)); }

static void normalize_or_zero(const float v[3], float out[3]) {
    const float rcp = 1.0f / length3(v);
    if (isfinite(rcp) && rcp > 0.0f) {
        out[0] = v[0] * rcp;
        out[1] = v[1] * rcp;
        out[2] = v[2] * rcp;
    } else {
        out[0] = out[1] = out[2] = 0.0f;
    }
}

/* renderer.rs:549-558 */
float vgo_area_render_distance(const float look[3], uint32_t render_size) {
    const float down[3] = {0.0f, 0.0f, 1.0f};
    const float dot_product = fabsf(dot3(look, down));
    const float dot_product_smooth = 1.0f - (1.0f - dot_product) * (1.0f - dot_product);
    const float render_distance = dot_product_smooth * (float)(VGO_AREA_SIZE * render_size);
    const float scaled = render_distance * LOOK_DOWN_RENDER_MULTIPLIER;
    return scaled > (float)VGO_AREA_SIZE ? scaled : (float)VGO_AREA_SIZE; /* f32::max */
}

/* renderer.rs:324-346 */
int vgo_is_area_visible(uint32_t ax, uint32_t ay, const vgo_view* v, float render_distance) {
    const int32_t mx = (int32_t)(ax * VGO_AREA_SIZE + VGO_AREA_SIZE / 2) - VGO_LOCATION_OFFSET;
    const int32_t my = (int32_t)(ay * VGO_AREA_SIZE + VGO_AREA_SIZE / 2) - VGO_LOCATION_OFFSET;
    const float area_vec[3] = {(float)mx, (float)my, v->cam_target_z};
    /* camera.position.distance(area_vec) = (position - area_vec).length() */
    const float d[3] = {v->cam_pos[0] - area_vec[0], v->cam_pos[1] - area_vec[1], v->cam_pos[2] - area_vec[2]};
    if (length3(d) <= render_distance) return 1;
    const float to_area[3] = {area_vec[0] - v->cam_pos[0], area_vec[1] - v->cam_pos[1], area_vec[2] - v->cam_pos[2]};
    float area_look[3];
    normalize_or_zero(to_area, area_look);
    return dot3(area_look, v->look) >= AREA_RENDER_THRESHOLD;
}

/* renderer.rs:560-586 */
int vgo_is_voxel_visible(uint32_t gx, uint32_t gy, uint32_t gz, const float look[3],
                         const float cam_pos[3], float max_distance) {
    /* Location::from(InternalLocation) then Vec3::from(Location) (location.rs:19-33) */
    const float loc[3] = {(float)((int32_t)gx - VGO_LOCATION_OFFSET), (float)((int32_t)gy - VGO_LOCATION_OFFSET),
                          (float)(int32_t)gz};
    const float dir[3] = {loc[0] - cam_pos[0], loc[1] - cam_pos[1], loc[2] - cam_pos[2]};
    if (fabsf(dir[0]) <= VOXEL_PROXIMITY_THRESHOLD && fabsf(dir[1]) <= VOXEL_PROXIMITY_THRESHOLD &&
        fabsf(dir[2]) <= VOXEL_PROXIMITY_THRESHOLD)
        return 1;
    const float distance = length3(dir);
    if (distance > max_distance) return 0;
    return dot3(dir, look) > VOXEL_RENDER_THRESHOLD * distance;
}

static int is_water(uint32_t voxel) { return voxel >= VGO_WATER_SOURCE && voxel <= VGO_WATER4; } /* voxel.rs:54-61 */

long vgo_filter_visible(const vgo_voxel_rec* recs, const uint32_t* rec_first, const uint32_t* area_xy,
                        int n, const vgo_view* v, uint8_t* area_visible, uint32_t* visible,
                        uint32_t* type_first, long* n_faces, uint32_t* lamps, long* n_lamps) {
    /* renderer.rs:373-387: the map view shows every meshed area and skips the voxel filter */
    const float area_distance = vgo_area_render_distance(v->look, v->render_size);
    for (int i = 0; i < n; ++i)
        area_visible[i] = v->show_map ? 1 : (uint8_t)vgo_is_area_visible(area_xy[2 * i], area_xy[2 * i + 1], v, area_distance);
    /* renderer.rs:530 */
    const float render_distance = (float)(v->render_size * VGO_AREA_SIZE);
    /* two passes: count per voxel type, then place (optimise_render_order, renderer.rs:349-361) */
    uint32_t count[VGO_VOXEL_VARIANTS + 1];
    memset(count, 0, sizeof count);
    long lamps_found = 0, faces = 0;
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t cursor[VGO_VOXEL_VARIANTS];
        if (pass == 1) {
            uint32_t run = 0;
            for (int t = 0; t < VGO_VOXEL_VARIANTS; ++t) {
                type_first[t] = run;
                cursor[t] = run;
                run += count[t];
            }
            type_first[VGO_VOXEL_VARIANTS] = run;
        }
        for (int i = 0; i < n; ++i) {
            if (!area_visible[i]) continue;
            for (uint32_t q = rec_first[i]; q < rec_first[i + 1]; ++q) {
                const vgo_voxel_rec* r = &recs[q];
                const uint32_t gz = r->packed & 0xFFu, voxel = (r->packed >> 8) & 0xFFu;
                if (pass == 0 && voxel == VGO_LAMP) { /* RenderArea.lights (renderer.rs:60-67) */
                    if (lamps) lamps[lamps_found] = q;
                    ++lamps_found;
                }
                int drawn;
                if (v->show_map)
                    drawn = 1;
                else /* renderer.rs:536-540 */
                    drawn = !(v->head_in_water && is_water(voxel)) &&
                            vgo_is_voxel_visible(r->gx, r->gy, gz, v->look, v->cam_pos, render_distance);
                if (!drawn) continue;
                if (pass == 0) {
                    ++count[voxel];
                    faces += r->packed >> 24;
                } else if (visible) {
                    visible[cursor[voxel]++] = q;
                }
            }
        }
    }
    if (n_faces) *n_faces = faces;
    if (n_lamps) *n_lamps = lamps_found;
    return (long)type_first[VGO_VOXEL_VARIANTS];
}
