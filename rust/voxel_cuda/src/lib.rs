//! Rust side of the drop-in boundary: `extern "C"` bindings of `include/voxel_cuda.h` plus the
//! thin safe layer the reference's call sites use.
//!
//! NOT COMPILED IN THE BUILD IMAGE (no cargo/rustc there); kept in sync with the header by hand and
//! exercised through the same C ABI by the C++ mirror (`voxel_game_b200/host/`) and the Python
//! tests.  See INTEGRATION.md for the three reference functions whose bodies call into this.
//!
//! There is deliberately no CPU fallback: if `Pipeline::new` fails the game cannot generate terrain.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr};

pub const VOXELS_IN_AREA: usize = 32768; // model/area.rs:9
pub const COLUMNS_IN_AREA: usize = 256;

#[repr(C)]
pub struct vx_ctx {
    _private: [u8; 0],
}

/// One visible voxel = key + value of `RenderArea.mesh_map` (graphics/renderer.rs:45-51).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct vx_voxel_rec {
    pub gx: u32,
    pub gy: u32,
    /// gz | voxel << 8 | face_mask << 16 | n_faces << 24
    pub packed: u32,
    pub first_face: u32,
}

/// `World::set(location, voxel)` (model/world.rs:184-191).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct vx_edit {
    pub gx: u32,
    pub gy: u32,
    pub gz: u32,
    pub voxel: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct vx_mesh_info {
    pub n_faces: u64,
    pub n_recs: u64,
    pub vertex_bytes: u64,
    pub index_bytes: u64,
    pub rec_bytes: u64,
}

/// Camera state of one frame (`vx_view`), see `Pipeline::visible`.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct vx_view {
    pub cam_pos: [f32; 3],
    pub cam_target_z: f32,
    pub look: [f32; 3],
    pub render_size: u32,
    pub head_in_water: u32,
    pub show_map: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct vx_visible_info {
    pub n_visible: u64,
    pub n_faces: u64,
    pub n_areas: u32,
    pub n_lamps: u32,
    pub type_first: [u32; 28],
    pub lamps: [u32; 64],
}

/// macroquad::ui::Vertex is `#[repr(C)] { position: Vec3, uv: Vec2, color: [u8; 4], normal: Vec4 }`
/// with glam's scalar-math feature = 40 bytes; the GPU writes exactly these bytes.
pub const VERTEX_BYTES: usize = 40;

#[link(name = "voxel_cuda")]
extern "C" {
    pub fn vx_device_count() -> c_int;
    pub fn vx_create(device: c_int, out: *mut *mut vx_ctx) -> c_int;
    pub fn vx_destroy(ctx: *mut vx_ctx);
    pub fn vx_strerror(status: c_int) -> *const c_char;
    pub fn vx_last_error(ctx: *mut vx_ctx) -> *const c_char;
    pub fn vx_copy_last_error(ctx: *mut vx_ctx, buf: *mut c_char, cap: usize) -> usize;
    pub fn vx_lock(ctx: *mut vx_ctx) -> c_int;
    pub fn vx_unlock(ctx: *mut vx_ctx) -> c_int;
    pub fn vx_set_stream(ctx: *mut vx_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn vx_sync(ctx: *mut vx_ctx) -> c_int;
    pub fn vx_host_alloc(bytes: usize) -> *mut c_void;
    pub fn vx_host_free(p: *mut c_void);
    pub fn vx_set_seed(ctx: *mut vx_ctx, seed: u64) -> c_int;
    pub fn vx_hash_world_name(world_name: *const c_char) -> u64;
    pub fn vx_cave_columns(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, cave_columns_out: *mut u32,
                           cave_voxels_out: *mut u32) -> c_int;
    pub fn vx_water_tick(ctx: *mut vx_ctx, check_xyz: *const u32, n: c_int, changed_xyzv: *mut u32, n_changed: *mut c_int,
                         next_xyz: *mut u32, n_next: *mut c_int) -> c_int;
    pub fn vx_falling_check(ctx: *mut vx_ctx, check_xyz: *const u32, n: c_int, spawned_xyzv: *mut u32, cap_spawned: c_int,
                            n_spawned: *mut c_int, water_next_xyz: *mut u32, cap_next: c_int, n_next: *mut c_int) -> c_int;
    pub fn vx_falling_land(ctx: *mut vx_ctx, positions: *const f32, voxel_types: *const u8, n: c_int, keep: *mut u8,
                           placed_xyzv: *mut u32, n_placed: *mut c_int, water_next_xyz: *mut u32, n_next: *mut c_int) -> c_int;
    pub fn vx_noise_variant() -> c_int;
    pub fn vx_set_gradient_order(ctx: *mut vx_ctx, order2: *const u8, order3: *const u8) -> c_int;
    pub fn vx_generate(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, voxels_out: *mut u8,
                       max_height_out: *mut u8) -> c_int;
    pub fn vx_upload(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, voxels: *const u8,
                     max_height_out: *mut u8) -> c_int;
    pub fn vx_read_areas(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, voxels_out: *mut u8,
                         max_height_out: *mut u8) -> c_int;
    pub fn vx_unload(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int) -> c_int;
    pub fn vx_resident_count(ctx: *mut vx_ctx) -> c_int;
    pub fn vx_set_deferred_reads(ctx: *mut vx_ctx, enabled: c_int) -> c_int;
    pub fn vx_wait_reads(ctx: *mut vx_ctx) -> c_int;
    pub fn vx_pool_info(ctx: *mut vx_ctx, capacity_areas: *mut u64, committed_bytes: *mut u64, virtual_memory: *mut c_int) -> c_int;
    pub fn vx_column_heights(ctx: *mut vx_ctx, voxels: *const u8, n: c_int, max_height_out: *mut u8) -> c_int;
    pub fn vx_mesh(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, info: *mut vx_mesh_info) -> c_int;
    pub fn vx_mesh_read(ctx: *mut vx_ctx, rec_first: *mut u32, face_first: *mut u32,
                        recs: *mut vx_voxel_rec, vertices: *mut c_void, indices: *mut u16) -> c_int;
    pub fn vx_apply_edits(ctx: *mut vx_ctx, edits: *const vx_edit, n: c_int, dirty_area_xy: *mut u32,
                          n_dirty: *mut c_int) -> c_int;
    pub fn vx_heightmap(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, cam_area_x: u32,
                        cam_area_y: u32, rgba: *mut u8) -> c_int;
    pub fn vx_encode_areas(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, total_bytes: *mut u64) -> c_int;
    pub fn vx_encode_read(ctx: *mut vx_ctx, files: *mut u8, offsets: *mut u64) -> c_int;
    pub fn vx_decode_areas(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, files: *const u8,
                           offsets: *const u64, ok: *mut u8, max_height_out: *mut u8) -> c_int;
    pub fn vx_explode(ctx: *mut vx_ctx, positions: *const f32, n: c_int, removed_xyz: *mut u32,
                      n_removed: *mut c_int, bomb_xyz: *mut u32, n_bombs: *mut c_int,
                      dirty_area_xy: *mut u32, n_dirty: *mut c_int) -> c_int;
    pub fn vx_visible(ctx: *mut vx_ctx, view: *const vx_view, info: *mut vx_visible_info) -> c_int;
    pub fn vx_visible_read(ctx: *mut vx_ctx, rec_index: *mut u32, area_visible: *mut u8) -> c_int;
    pub fn vx_light(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, light_out: *mut u8,
                    iterations: *mut c_int) -> c_int;
    pub fn vx_update_locations(ctx: *mut vx_ctx, xyz: *const u32, n: c_int, info: *mut vx_mesh_info) -> c_int;
    pub fn vx_update_read(ctx: *mut vx_ctx, recs: *mut vx_voxel_rec, vertices: *mut c_void, indices: *mut u16) -> c_int;
    pub fn vx_mesh_read_compact(ctx: *mut vx_ctx, rec_first: *mut u32, packed_recs: *mut u32) -> c_int;
    pub fn vx_set_mesh_retention(ctx: *mut vx_ctx, enabled: c_int) -> c_int;
    pub fn vx_mesh_unload(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int) -> c_int;
    pub fn vx_read_face_masks(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, masks_out: *mut u8,
                              meshed_out: *mut u8) -> c_int;
    pub fn vx_visible_zone(ctx: *mut vx_ctx, area_xy: *const u32, n: c_int, view: *const vx_view,
                           info: *mut vx_visible_info) -> c_int;
}

#[derive(Debug)]
pub struct VxError {
    pub status: i32,
    pub message: String,
}

/// Owns one `vx_ctx` (one CUDA device).  `Send + Sync`: the C side locks every call internally,
/// which is what lets `AreaLoader`'s rayon workers share it
/// (service/persistence/world_persistence.rs:90-93).  Results that are fetched by a SECOND call
/// (`vx_mesh` -> `vx_mesh_read`, `vx_encode_areas` -> `vx_encode_read`, `vx_visible` ->
/// `vx_visible_read`, `vx_update_locations` -> `vx_update_read`, a failure -> its error text) belong
/// to whichever first call ran last on the context, so every method below that makes such a pair
/// holds a [`Transaction`] (the context's recursive lock, `vx_lock` / `vx_unlock`) across it: two
/// threads in `mesh()` can never read each other's sizes or buffers.
pub struct Pipeline {
    ctx: *mut vx_ctx,
}
unsafe impl Send for Pipeline {}
unsafe impl Sync for Pipeline {}

/// RAII guard of the context's recursive lock; calls made by this thread while it lives are one unit.
pub struct Transaction<'a> {
    p: &'a Pipeline,
}
impl<'a> Transaction<'a> {
    fn new(p: &'a Pipeline) -> Self {
        unsafe { vx_lock(p.ctx) };
        Transaction { p }
    }
}
impl Drop for Transaction<'_> {
    fn drop(&mut self) {
        unsafe { vx_unlock(self.p.ctx) };
    }
}

/// Result of `update_locations`: one record per touched voxel (`n_faces == 0` = removal).
pub struct UpdateBatch {
    pub recs: Vec<vx_voxel_rec>,
    pub vertices: Vec<u8>,
    pub indices: Vec<u16>,
}

/// Flat mesh of a batch of areas; `recs[rec_first[i]..rec_first[i + 1]]` belong to area `i`.
pub struct MeshBatch {
    pub rec_first: Vec<u32>,
    pub face_first: Vec<u32>,
    pub recs: Vec<vx_voxel_rec>,
    /// 4 vertices per face, `VERTEX_BYTES` each, byte-identical to `macroquad::ui::Vertex`
    pub vertices: Vec<u8>,
    /// 6 per face: `[0,1,2,0,2,3] + 4 * (face ordinal inside its voxel)`
    pub indices: Vec<u16>,
}

impl Pipeline {
    /// `world_name` is the seed, exactly as in `AreaGenerator::generate_area(_, world_name)`.
    pub fn new(device: i32, world_name: &str) -> Result<Self, VxError> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { vx_create(device, &mut ctx) };
        if st != 0 {
            return Err(VxError { status: st, message: strerror(st) });
        }
        let p = Pipeline { ctx };
        // hash_world_name stays in Rust (std DefaultHasher, generator.rs:24-28) so that the seed is
        // whatever the running std produces; vx_hash_world_name is for non-Rust hosts.
        p.check(unsafe { vx_set_seed(p.ctx, hash_world_name(world_name)) })?;
        Ok(p)
    }

    /// The calling thread's own unit of several calls (see the type's documentation).
    pub fn transaction(&self) -> Transaction<'_> {
        Transaction::new(self)
    }

    fn check(&self, st: c_int) -> Result<(), VxError> {
        if st == 0 {
            return Ok(());
        }
        // copied out under the context's lock: another thread may be assigning the string
        let mut buf = vec![0 as c_char; 1024];
        unsafe { vx_copy_last_error(self.ctx, buf.as_mut_ptr(), buf.len()) };
        let message = unsafe { CStr::from_ptr(buf.as_ptr()) }.to_string_lossy().into_owned();
        Err(VxError { status: st, message })
    }

    /// Body of `AreaLoader::load_all_blocking` for the areas that are not on disk: returns, per
    /// area, the 32 768 voxel bytes and the 256 column heights.
    pub fn generate(&self, areas: &[(u32, u32)]) -> Result<(Vec<u8>, Vec<u8>), VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut voxels = vec![0u8; areas.len() * VOXELS_IN_AREA];
        let mut heights = vec![0u8; areas.len() * COLUMNS_IN_AREA];
        self.check(unsafe {
            vx_generate(self.ctx, xy.as_ptr(), areas.len() as c_int, voxels.as_mut_ptr(), heights.as_mut_ptr())
        })?;
        Ok((voxels, heights))
    }

    /// Areas read from disk (`read_binary_object` hit in `load_blocking`): make them resident and
    /// get their column heights (`AreaDTO::into_area`).
    pub fn upload(&self, areas: &[(u32, u32)], voxels: &[u8]) -> Result<Vec<u8>, VxError> {
        assert_eq!(voxels.len(), areas.len() * VOXELS_IN_AREA);
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut heights = vec![0u8; areas.len() * COLUMNS_IN_AREA];
        self.check(unsafe {
            vx_upload(self.ctx, xy.as_ptr(), areas.len() as c_int, voxels.as_ptr(), heights.as_mut_ptr())
        })?;
        Ok(heights)
    }

    pub fn unload(&self, areas: &[(u32, u32)]) -> Result<(), VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        self.check(unsafe { vx_unload(self.ctx, xy.as_ptr(), areas.len() as c_int) })
    }

    /// Body of `Renderer::load_all_blocking` / `load_areas_in_queue`.
    pub fn mesh(&self, areas: &[(u32, u32)]) -> Result<MeshBatch, VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut info = vx_mesh_info::default();
        let _pair = self.transaction(); // vx_mesh + vx_mesh_read see the same batch
        self.check(unsafe { vx_mesh(self.ctx, xy.as_ptr(), areas.len() as c_int, &mut info) })?;
        let mut m = MeshBatch {
            rec_first: vec![0; areas.len() + 1],
            face_first: vec![0; areas.len() + 1],
            recs: vec![vx_voxel_rec::default(); info.n_recs as usize],
            vertices: vec![0u8; info.vertex_bytes as usize],
            indices: vec![0u16; info.n_faces as usize * 6],
        };
        self.check(unsafe {
            vx_mesh_read(self.ctx, m.rec_first.as_mut_ptr(), m.face_first.as_mut_ptr(), m.recs.as_mut_ptr(),
                         m.vertices.as_mut_ptr() as *mut c_void, m.indices.as_mut_ptr())
        })?;
        Ok(m)
    }

    /// Batched `World::set`; returns the areas whose meshes must be rebuilt with `mesh`.
    pub fn apply_edits(&self, edits: &[vx_edit]) -> Result<Vec<(u32, u32)>, VxError> {
        let mut dirty = vec![0u32; edits.len().max(1) * 6];
        let mut n: c_int = 0;
        self.check(unsafe {
            vx_apply_edits(self.ctx, edits.as_ptr(), edits.len() as c_int, dirty.as_mut_ptr(), &mut n)
        })?;
        Ok(dirty.chunks(2).take(n as usize).map(|c| (c[0], c[1])).collect())
    }

    /// Pixel loop of `HeightMap::generate_height_map`; `rgba` is the persistent 512x512 image.
    pub fn heightmap(&self, areas: &[(u32, u32)], cam: (u32, u32), rgba: &mut [u8]) -> Result<(), VxError> {
        assert_eq!(rgba.len(), 512 * 512 * 4);
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        self.check(unsafe { vx_heightmap(self.ctx, xy.as_ptr(), areas.len() as c_int, cam.0, cam.1, rgba.as_mut_ptr()) })
    }
}

impl Pipeline {
    /// `store_all_blocking` (world_persistence.rs:55-59) for resident areas: the bytes of every
    /// `saves/<world>/area<x>_<y>.dat`, compressed on the device.
    pub fn encode_areas(&self, areas: &[(u32, u32)]) -> Result<Vec<Vec<u8>>, VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut total = 0u64;
        let _pair = self.transaction(); // vx_encode_areas + vx_encode_read
        self.check(unsafe { vx_encode_areas(self.ctx, xy.as_ptr(), areas.len() as c_int, &mut total) })?;
        let mut blob = vec![0u8; total as usize];
        let mut offsets = vec![0u64; areas.len() + 1];
        self.check(unsafe { vx_encode_read(self.ctx, blob.as_mut_ptr(), offsets.as_mut_ptr()) })?;
        self.check(unsafe { vx_wait_reads(self.ctx) })?; // no-op unless deferred reads are on
        Ok(offsets.windows(2).map(|w| blob[w[0] as usize..w[1] as usize].to_vec()).collect())
    }

    /// `load_blocking` (world_persistence.rs:62-71) for a batch of files read from disk: returns
    /// which files were well-formed; the others are generated by the caller, as the reference does.
    pub fn decode_areas(&self, areas: &[(u32, u32)], files: &[Vec<u8>]) -> Result<(Vec<bool>, Vec<u8>), VxError> {
        assert_eq!(areas.len(), files.len());
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut offsets = vec![0u64; files.len() + 1];
        for (i, f) in files.iter().enumerate() {
            offsets[i + 1] = offsets[i] + f.len() as u64;
        }
        let blob: Vec<u8> = files.concat();
        let mut ok = vec![0u8; files.len()];
        let mut max_height = vec![0u8; files.len() * 256];
        self.check(unsafe {
            vx_decode_areas(self.ctx, xy.as_ptr(), areas.len() as c_int, blob.as_ptr(), offsets.as_ptr(),
                            ok.as_mut_ptr(), max_height.as_mut_ptr())
        })?;
        Ok((ok.into_iter().map(|b| b != 0).collect(), max_height))
    }

    /// The voxel loops of `BombSimulator::handle_explosions` for one tick (at most 32 positions):
    /// removed voxels and found bombs as internal (x, y, z), in the reference's visiting order, and
    /// the areas to re-mesh.
    #[allow(clippy::type_complexity)]
    pub fn explode(&self, positions: &[[f32; 3]]) -> Result<(Vec<[u32; 3]>, Vec<[u32; 3]>, Vec<(u32, u32)>), VxError> {
        let n = positions.len();
        let flat: Vec<f32> = positions.iter().flatten().copied().collect();
        let mut removed = vec![0u32; 3 * 1331 * n.max(1)];
        let mut bombs = vec![0u32; 3 * 1331 * n.max(1)];
        let mut dirty = vec![0u32; 2 * 9 * n.max(1)];
        let (mut nr, mut nb, mut nd) = (0 as c_int, 0 as c_int, 0 as c_int);
        self.check(unsafe {
            vx_explode(self.ctx, flat.as_ptr(), n as c_int, removed.as_mut_ptr(), &mut nr, bombs.as_mut_ptr(),
                       &mut nb, dirty.as_mut_ptr(), &mut nd)
        })?;
        let triples = |v: &[u32], k: c_int| v.chunks(3).take(k as usize).map(|c| [c[0], c[1], c[2]]).collect();
        Ok((triples(&removed, nr), triples(&bombs, nb),
            dirty.chunks(2).take(nd as usize).map(|c| (c[0], c[1])).collect()))
    }

    /// One frame of `set_voxel_shader_and_find_visible_areas` + `render_voxels`
    /// (graphics/renderer.rs:365-452) over the batch of the last `mesh` call: returns the summary
    /// and the record indices to draw, already in `optimise_render_order`.
    pub fn visible(&self, view: &vx_view) -> Result<(vx_visible_info, Vec<u32>), VxError> {
        let mut info = std::mem::MaybeUninit::<vx_visible_info>::zeroed();
        let _pair = self.transaction(); // vx_visible + vx_visible_read
        self.check(unsafe { vx_visible(self.ctx, view, info.as_mut_ptr()) })?;
        let info = unsafe { info.assume_init() };
        let mut idx = vec![0u32; info.n_visible as usize];
        self.check(unsafe { vx_visible_read(self.ctx, idx.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok((info, idx))
    }
}

impl Pipeline {
    /// Body of `Renderer::update_location` (graphics/renderer.rs:239-286) for a batch of locations
    /// whose voxels `apply_edits` / `explode` have just changed: the <= 7 voxel meshes per location.
    pub fn update_locations(&self, locations: &[[u32; 3]]) -> Result<UpdateBatch, VxError> {
        let flat: Vec<u32> = locations.iter().flatten().copied().collect();
        let mut info = vx_mesh_info::default();
        let _pair = self.transaction(); // vx_update_locations + vx_update_read
        self.check(unsafe { vx_update_locations(self.ctx, flat.as_ptr(), locations.len() as c_int, &mut info) })?;
        let mut u = UpdateBatch {
            recs: vec![vx_voxel_rec::default(); info.n_recs as usize],
            vertices: vec![0u8; info.vertex_bytes as usize],
            indices: vec![0u16; info.n_faces as usize * 6],
        };
        self.check(unsafe {
            vx_update_read(self.ctx, u.recs.as_mut_ptr(), u.vertices.as_mut_ptr() as *mut c_void, u.indices.as_mut_ptr())
        })?;
        Ok(u)
    }

    /// `mesh` with the compact transport: 4 bytes per visible voxel, `(x + 16y + 256z) | voxel << 15 |
    /// face_mask << 20`; vertices and indices are rebuilt on the host (`MeshGenerator::generate_mesh`
    /// stays as it is and is fed the mask).  Returns `(rec_first, packed)`.
    pub fn mesh_compact(&self, areas: &[(u32, u32)]) -> Result<(Vec<u32>, Vec<u32>), VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut info = vx_mesh_info::default();
        let _pair = self.transaction(); // vx_mesh + vx_mesh_read_compact
        self.check(unsafe { vx_mesh(self.ctx, xy.as_ptr(), areas.len() as c_int, &mut info) })?;
        let mut rec_first = vec![0u32; areas.len() + 1];
        let mut packed = vec![0u32; info.n_recs as usize];
        self.check(unsafe { vx_mesh_read_compact(self.ctx, rec_first.as_mut_ptr(), packed.as_mut_ptr()) })?;
        self.check(unsafe { vx_wait_reads(self.ctx) })?; // no-op unless deferred reads are on
        Ok((rec_first, packed))
    }

    /// Deferred reads (`vx_set_deferred_reads`): `encode_read_into` / `mesh_read_compact_into` return
    /// before their copies are done, so one thread can issue the next step at once; the buffers are
    /// valid after `wait_reads`.  A pre-generation tool uses this with two sets of pinned buffers
    /// (`vx_host_alloc`); the game's own paths keep the blocking wrappers above.
    pub fn set_deferred_reads(&self, enabled: bool) -> Result<(), VxError> {
        self.check(unsafe { vx_set_deferred_reads(self.ctx, enabled as c_int) })
    }

    pub fn wait_reads(&self) -> Result<(), VxError> {
        self.check(unsafe { vx_wait_reads(self.ctx) })
    }

    /// # Safety
    /// `files` (the size `vx_encode_areas` reported) and `offsets` (n + 1, or null) must stay valid and
    /// untouched until `wait_reads` returns; pinned memory, or the copy is not asynchronous.
    pub unsafe fn encode_read_into(&self, files: *mut u8, offsets: *mut u64) -> Result<(), VxError> {
        self.check(vx_encode_read(self.ctx, files, offsets))
    }

    /// # Safety
    /// As `encode_read_into`: `rec_first` (n + 1) and `packed` (`n_recs` of the last `vx_mesh`).
    pub unsafe fn mesh_read_compact_into(&self, rec_first: *mut u32, packed: *mut u32) -> Result<(), VxError> {
        self.check(vx_mesh_read_compact(self.ctx, rec_first, packed))
    }

    /// Body of `WaterSimulator::simulate_voxels` (service/physics/water_simulator.rs:55-77): `check` is
    /// `mem::take(&mut self.check_locations)` drained into a Vec -- the order of that Vec is the order
    /// applied.  Returns (changed: x, y, z, new voxel -> `update_locations`; next: what to insert into
    /// `self.check_locations`).
    #[allow(clippy::type_complexity)]
    pub fn water_tick(&self, check: &[[u32; 3]]) -> Result<(Vec<[u32; 4]>, Vec<[u32; 3]>), VxError> {
        let n = check.len();
        let flat: Vec<u32> = check.iter().flatten().copied().collect();
        let mut changed = vec![0u32; 16 * n.max(1)];
        let mut next = vec![0u32; 21 * n.max(1)];
        let (mut nc, mut nn) = (0 as c_int, 0 as c_int);
        self.check(unsafe {
            vx_water_tick(self.ctx, flat.as_ptr(), n as c_int, changed.as_mut_ptr(), &mut nc, next.as_mut_ptr(), &mut nn)
        })?;
        Ok((changed.chunks(4).take(nc as usize).map(|c| [c[0], c[1], c[2], c[3]]).collect(),
            next.chunks(3).take(nn as usize).map(|c| [c[0], c[1], c[2]]).collect()))
    }

    /// Body of `FallingVoxelSimulator::update_voxels` (falling_voxel_simulator.rs:105-150) for a batch of
    /// calls in order: the entities to create (x, y, z, voxel; position = location, velocity 0) and the
    /// locations `water_simulator.location_updated` inserts.
    #[allow(clippy::type_complexity)]
    pub fn falling_check(&self, check: &[[u32; 3]], capacity: usize) -> Result<(Vec<[u32; 4]>, Vec<[u32; 3]>), VxError> {
        let flat: Vec<u32> = check.iter().flatten().copied().collect();
        let mut spawned = vec![0u32; 4 * capacity.max(1)];
        let mut next = vec![0u32; 21 * capacity.max(1)];
        let (mut ns, mut nn) = (0 as c_int, 0 as c_int);
        self.check(unsafe {
            vx_falling_check(self.ctx, flat.as_ptr(), check.len() as c_int, spawned.as_mut_ptr(), capacity as c_int, &mut ns,
                             next.as_mut_ptr(), (7 * capacity) as c_int, &mut nn)
        })?;
        Ok((spawned.chunks(4).take(ns as usize).map(|c| [c[0], c[1], c[2], c[3]]).collect(),
            next.chunks(3).take(nn as usize).map(|c| [c[0], c[1], c[2]]).collect()))
    }

    /// The `retain` pass of `simulate_falling` (:152-166) after the positions have been advanced:
    /// which entities keep falling, which voxels were put down, what the water simulator is told.
    #[allow(clippy::type_complexity)]
    pub fn falling_land(&self, positions: &[[f32; 3]], voxel_types: &[u8]) -> Result<(Vec<bool>, Vec<[u32; 4]>, Vec<[u32; 3]>), VxError> {
        let n = positions.len();
        assert_eq!(voxel_types.len(), n);
        let flat: Vec<f32> = positions.iter().flatten().copied().collect();
        let mut keep = vec![0u8; n.max(1)];
        let mut placed = vec![0u32; 4 * n.max(1)];
        let mut next = vec![0u32; 21 * n.max(1)];
        let (mut np, mut nn) = (0 as c_int, 0 as c_int);
        self.check(unsafe {
            vx_falling_land(self.ctx, flat.as_ptr(), voxel_types.as_ptr(), n as c_int, keep.as_mut_ptr(), placed.as_mut_ptr(),
                            &mut np, next.as_mut_ptr(), &mut nn)
        })?;
        Ok((keep.into_iter().take(n).map(|k| k != 0).collect(),
            placed.chunks(4).take(np as usize).map(|c| [c[0], c[1], c[2], c[3]]).collect(),
            next.chunks(3).take(nn as usize).map(|c| [c[0], c[1], c[2]]).collect()))
    }

    /// Cost hint for sharding a region over several GPUs, from the seed alone: per area the cave-zone
    /// columns and the voxels that run the cave noise.
    pub fn cave_columns(&self, areas: &[(u32, u32)]) -> Result<(Vec<u32>, Vec<u32>), VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        let (mut cols, mut vox) = (vec![0u32; areas.len()], vec![0u32; areas.len()]);
        self.check(unsafe {
            vx_cave_columns(self.ctx, xy.as_ptr(), areas.len() as c_int, cols.as_mut_ptr(), vox.as_mut_ptr())
        })?;
        Ok((cols, vox))
    }

    /// libnoise calibration (include/voxel_cuda.h): the mask of code switches this library was built
    /// with, and the order of the crate's two gradient tables as data.
    pub fn noise_variant() -> i32 {
        unsafe { vx_noise_variant() }
    }
    pub fn set_gradient_order(&self, order2: Option<&[u8; 24]>, order3: Option<&[u8; 12]>) -> Result<(), VxError> {
        self.check(unsafe {
            vx_set_gradient_order(self.ctx, order2.map_or(std::ptr::null(), |o| o.as_ptr()),
                                  order3.map_or(std::ptr::null(), |o| o.as_ptr()))
        })
    }

    /// Device mirror of `Renderer.meshes` on / off (graphics/renderer.rs:97-101).
    pub fn set_mesh_retention(&self, enabled: bool) -> Result<(), VxError> {
        self.check(unsafe { vx_set_mesh_retention(self.ctx, enabled as c_int) })
    }

    /// `Renderer::unload_area` (graphics/renderer.rs:111-117) for a batch.
    pub fn mesh_unload(&self, areas: &[(u32, u32)]) -> Result<(), VxError> {
        let xy: Vec<u32> = areas.iter().flat_map(|a| [a.0, a.1]).collect();
        self.check(unsafe { vx_mesh_unload(self.ctx, xy.as_ptr(), areas.len() as c_int) })
    }

    /// One frame over the resident meshes of the render zone: the summary and, per drawn voxel,
    /// `zone index << 15 | (x + 16y + 256z)` in `optimise_render_order`.
    pub fn visible_zone(&self, zone: &[(u32, u32)], view: &vx_view) -> Result<(vx_visible_info, Vec<u32>), VxError> {
        let xy: Vec<u32> = zone.iter().flat_map(|a| [a.0, a.1]).collect();
        let mut info = std::mem::MaybeUninit::<vx_visible_info>::zeroed();
        let _pair = self.transaction(); // vx_visible_zone + vx_visible_read
        self.check(unsafe { vx_visible_zone(self.ctx, xy.as_ptr(), zone.len() as c_int, view, info.as_mut_ptr()) })?;
        let info = unsafe { info.assume_init() };
        let mut codes = vec![0u32; info.n_visible as usize];
        self.check(unsafe { vx_visible_read(self.ctx, codes.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok((info, codes))
    }
}

impl Drop for Pipeline {
    fn drop(&mut self) {
        unsafe { vx_destroy(self.ctx) }
    }
}

fn strerror(st: c_int) -> String {
    unsafe { CStr::from_ptr(vx_strerror(st)) }.to_string_lossy().into_owned()
}

/// service/area_generation/generator.rs:24-28, unchanged.
pub fn hash_world_name(world_name: &str) -> u64 {
    use std::hash::{DefaultHasher, Hash, Hasher};
    let mut hasher = DefaultHasher::new();
    world_name.hash(&mut hasher);
    hasher.finish()
}
