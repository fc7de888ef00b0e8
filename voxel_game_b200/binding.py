"""ctypes binding of libvoxel_cuda.so (the C ABI declared in include/voxel_cuda.h).

Python is plumbing here: every voxel, height and vertex is produced by the CUDA kernels in
voxel_game_b200/csrc.  There is no CPU fallback -- if the shared library is missing or no CUDA
device is usable, importing/creating fails loudly.
"""
from __future__ import annotations

import contextlib
import functools
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: VX_LIB_PATH selects another build of the library (the VX_NOISE_VARIANT builds under variants/)
LIB_PATH = os.environ.get("VX_LIB_PATH") or os.path.join(_HERE, "libvoxel_cuda.so")
CSRC = os.path.join(_HERE, "csrc")

VOXELS_IN_AREA = 32768
COLUMNS_IN_AREA = 256

REC_DTYPE = np.dtype([("gx", "<u4"), ("gy", "<u4"), ("packed", "<u4"), ("first_face", "<u4")])
VERTEX_DTYPE = np.dtype(
    [("position", "<f4", 3), ("uv", "<f4", 2), ("color", "u1", 4), ("normal", "<f4", 4)]
)
EDIT_DTYPE = np.dtype([("gx", "<u4"), ("gy", "<u4"), ("gz", "<u4"), ("voxel", "<u4")])

VX_OK = 0
VX_ERR_NO_DEVICE = -1
VX_ERR_CUDA = -2
VX_ERR_INVALID = -3
VX_ERR_NO_SEED = -4
VX_ERR_NOT_LOADED = -5
VX_ERR_OOM = -6
VX_ERR_NO_MESH = -7

#: every symbol include/voxel_cuda.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "vx_device_count", "vx_create", "vx_destroy", "vx_strerror", "vx_last_error", "vx_set_stream",
    "vx_sync", "vx_set_timing", "vx_get_timings", "vx_host_alloc", "vx_host_free", "vx_set_seed",
    "vx_hash_world_name", "vx_generate", "vx_upload", "vx_read_areas", "vx_unload",
    "vx_resident_count", "vx_column_heights", "vx_mesh", "vx_mesh_read", "vx_apply_edits",
    "vx_heightmap", "vx_visible", "vx_visible_read", "vx_encode_areas", "vx_encode_read", "vx_decode_areas",
    "vx_explode", "vx_light", "vx_diag_fp64_rate", "vx_update_locations", "vx_update_read",
    "vx_mesh_read_compact", "vx_set_mesh_retention", "vx_mesh_unload", "vx_read_face_masks", "vx_visible_zone",
    "vx_lock", "vx_unlock", "vx_copy_last_error", "vx_noise_variant", "vx_set_gradient_order",
    "vx_cave_columns", "vx_water_tick", "vx_falling_check", "vx_falling_land",
    "vx_pool_info", "vx_set_deferred_reads", "vx_wait_reads",
]


class VxError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"vx status {status}: {message}")
        self.status = status


class MeshInfo(C.Structure):
    _fields_ = [("n_faces", C.c_uint64), ("n_recs", C.c_uint64), ("vertex_bytes", C.c_uint64),
                ("index_bytes", C.c_uint64), ("rec_bytes", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [("gen_columns_ms", C.c_float), ("gen_caves_ms", C.c_float),
                ("gen_finish_ms", C.c_float), ("mesh_count_ms", C.c_float),
                ("mesh_emit_ms", C.c_float), ("edit_ms", C.c_float), ("heightmap_ms", C.c_float),
                ("light_ms", C.c_float), ("visible_ms", C.c_float), ("gen_launches", C.c_uint32),
                ("mesh_launches", C.c_uint32), ("other_launches", C.c_uint32),
                ("cave_columns", C.c_uint32), ("codec_ms", C.c_float), ("simplex2_evals", C.c_uint64),
                ("simplex3_evals", C.c_uint64), ("cave_voxels", C.c_uint64), ("cave_evals", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class View(C.Structure):
    """vx_view: camera state of one frame (renderer.rs:365-462)."""
    _fields_ = [("cam_pos", C.c_float * 3), ("cam_target_z", C.c_float), ("look", C.c_float * 3),
                ("render_size", C.c_uint32), ("head_in_water", C.c_uint32), ("show_map", C.c_uint32)]


class VisibleInfo(C.Structure):
    _fields_ = [("n_visible", C.c_uint64), ("n_faces", C.c_uint64), ("n_areas", C.c_uint32),
                ("n_lamps", C.c_uint32), ("type_first", C.c_uint32 * 28), ("lamps", C.c_uint32 * 64)]


def make_view(cam_pos, cam_target, render_size=8, head_in_water=False, show_map=False) -> View:
    """Camera3D position / target -> vx_view; look = (target - position).normalize_or_zero() in
    f32, as render_voxels computes it (renderer.rs:425)."""
    p = np.asarray(cam_pos, np.float32)
    t = np.asarray(cam_target, np.float32)
    d = (t - p).astype(np.float32)
    length = np.sqrt(np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]), dtype=np.float32)
    with np.errstate(divide="ignore"):
        rcp = np.float32(1.0) / length
    look = (d * rcp).astype(np.float32) if np.isfinite(rcp) and rcp > 0 else np.zeros(3, np.float32)
    v = View()
    v.cam_pos[:] = [float(x) for x in p]
    v.cam_target_z = float(t[2])
    v.look[:] = [float(x) for x in look]
    v.render_size, v.head_in_water, v.show_map = int(render_size), int(bool(head_in_water)), int(bool(show_map))
    return v


#: the libnoise-calibration builds (csrc/Makefile `variants`): mask -> path
VARIANT_LIBS = {v: os.path.join(_HERE, "variants", f"libvoxel_cuda_v{v}.so") for v in (1, 6, 8)}  # Fbm form; both tie-breaks; skew digits


def build(verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a with nvcc (csrc/Makefile), the host mirror and the
    VX_NOISE_VARIANT builds. Returns the .so path."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvoxel_cuda.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "host")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvoxel_host.so failed:\n" + r.stdout + r.stderr)
    r = subprocess.run(["make", "-C", CSRC, "-j", "4", "variants"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the VX_NOISE_VARIANT libraries failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and declare the prototypes. Raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    L.vx_device_count.restype = C.c_int
    L.vx_create.restype = C.c_int
    L.vx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.vx_destroy.argtypes = [vp]
    L.vx_destroy.restype = None
    L.vx_strerror.restype = C.c_char_p
    L.vx_strerror.argtypes = [C.c_int]
    L.vx_last_error.restype = C.c_char_p
    L.vx_last_error.argtypes = [vp]
    L.vx_lock.argtypes = [vp]
    L.vx_unlock.argtypes = [vp]
    L.vx_copy_last_error.restype = C.c_size_t
    L.vx_copy_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.vx_set_stream.argtypes = [vp, vp]
    L.vx_sync.argtypes = [vp]
    L.vx_set_timing.argtypes = [vp, C.c_int]
    L.vx_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.vx_host_alloc.restype = vp
    L.vx_host_alloc.argtypes = [C.c_size_t]
    L.vx_host_free.argtypes = [vp]
    L.vx_host_free.restype = None
    L.vx_set_seed.argtypes = [vp, C.c_uint64]
    L.vx_cave_columns.argtypes = [vp, vp, C.c_int, vp, vp]
    ip = C.POINTER(C.c_int)
    L.vx_water_tick.argtypes = [vp, vp, C.c_int, vp, ip, vp, ip]
    L.vx_falling_check.argtypes = [vp, vp, C.c_int, vp, C.c_int, ip, vp, C.c_int, ip]
    L.vx_falling_land.argtypes = [vp, vp, vp, C.c_int, vp, vp, ip, vp, ip]
    L.vx_noise_variant.restype = C.c_int
    L.vx_noise_variant.argtypes = []
    L.vx_set_gradient_order.argtypes = [vp, vp, vp]
    L.vx_hash_world_name.restype = C.c_uint64
    L.vx_hash_world_name.argtypes = [C.c_char_p]
    L.vx_generate.argtypes = [vp, vp, C.c_int, vp, vp]
    L.vx_upload.argtypes = [vp, vp, C.c_int, vp, vp]
    L.vx_read_areas.argtypes = [vp, vp, C.c_int, vp, vp]
    L.vx_unload.argtypes = [vp, vp, C.c_int]
    L.vx_resident_count.argtypes = [vp]
    L.vx_pool_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    L.vx_column_heights.argtypes = [vp, vp, C.c_int, vp]
    L.vx_mesh.argtypes = [vp, vp, C.c_int, C.POINTER(MeshInfo)]
    L.vx_mesh_read.argtypes = [vp, vp, vp, vp, vp, vp]
    L.vx_apply_edits.argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_int)]
    L.vx_heightmap.argtypes = [vp, vp, C.c_int, C.c_uint32, C.c_uint32, vp]
    L.vx_explode.argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int)]
    L.vx_encode_areas.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_uint64)]
    L.vx_encode_read.argtypes = [vp, vp, vp]
    L.vx_decode_areas.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    L.vx_visible.argtypes = [vp, C.POINTER(View), C.POINTER(VisibleInfo)]
    L.vx_visible_read.argtypes = [vp, vp, vp]
    L.vx_light.argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_int)]
    L.vx_diag_fp64_rate.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.vx_update_locations.argtypes = [vp, vp, C.c_int, C.POINTER(MeshInfo)]
    L.vx_update_read.argtypes = [vp, vp, vp, vp]
    L.vx_mesh_read_compact.argtypes = [vp, vp, vp]
    L.vx_set_mesh_retention.argtypes = [vp, C.c_int]
    L.vx_set_deferred_reads.argtypes = [vp, C.c_int]
    L.vx_wait_reads.argtypes = [vp]
    L.vx_mesh_unload.argtypes = [vp, vp, C.c_int]
    L.vx_read_face_masks.argtypes = [vp, vp, C.c_int, vp, vp]
    L.vx_visible_zone.argtypes = [vp, vp, C.c_int, C.POINTER(View), C.POINTER(VisibleInfo)]
    _lib = L
    return L


def hash_world_name(name: str) -> int:
    """hash_world_name (generator.rs:24-28) as computed by the library."""
    return int(load().vx_hash_world_name(name.encode("utf-8")))


class PinnedArray:
    """numpy view over cudaMallocHost memory (vx_host_alloc)."""

    def __init__(self, shape, dtype):
        self._lib = load()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._ptr = self._lib.vx_host_alloc(max(n, 1))
        if not self._ptr:
            raise MemoryError(f"vx_host_alloc({n}) failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.vx_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def unpack_records(area_xy, rec_first, packed) -> np.ndarray:
    """4-byte records of mesh_read_compact -> the 16-byte records of mesh_read (numpy; tests and
    tools -- the product-side expander is the C++ one in host/voxel_game.hpp)."""
    xy = _xy(area_xy).astype(np.uint64)
    rec_first = np.asarray(rec_first, np.int64)
    packed = np.asarray(packed, np.uint32)
    area = np.repeat(np.arange(len(xy)), np.diff(rec_first))
    local = packed & 0x7FFF
    voxel = (packed >> 15) & 0x1F
    mask = (packed >> 20) & 0x3F
    nf = np.zeros_like(mask)
    for b in range(6):
        nf += (mask >> b) & 1
    recs = np.zeros(len(packed), REC_DTYPE)
    recs["gx"] = (xy[area, 0] * 16 + (local & 15)).astype(np.uint32)
    recs["gy"] = (xy[area, 1] * 16 + ((local >> 4) & 15)).astype(np.uint32)
    recs["packed"] = (local >> 8) | (voxel << 8) | (mask << 16) | (nf << 24)
    # first_face: running face count over the whole stream (areas in call order)
    first = np.zeros(len(packed), np.uint64)
    if len(packed):
        first[1:] = np.cumsum(nf.astype(np.uint64))[:-1]
    recs["first_face"] = first.astype(np.uint32)
    return recs


def _xy(area_xy) -> np.ndarray:
    a = np.ascontiguousarray(area_xy, dtype=np.uint32).reshape(-1, 2)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


def _locked(method):
    """Run a two-call helper (query + read of its result) under vx_lock."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with self.transaction():
            return method(self, *args, **kwargs)
    return wrapper


class Context:
    """One vx_ctx: a CUDA device, a seed and the set of resident areas."""

    def __init__(self, device: int = 0, seed: int | None = None, world_name: str | None = None):
        self._lib = load()
        h = C.c_void_p()
        st = self._lib.vx_create(device, C.byref(h))
        if st != VX_OK:
            raise VxError(st, self._lib.vx_strerror(st).decode())
        self._h = h
        self._deferred = False
        self.device = device
        if world_name is not None:
            seed = hash_world_name(world_name)
        if seed is not None:
            self.set_seed(seed)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.vx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st: int):
        if st != VX_OK:
            buf = C.create_string_buffer(1024)
            self._lib.vx_copy_last_error(self._h, buf, len(buf))
            raise VxError(st, buf.value.decode() or self._lib.vx_strerror(st).decode())

    @contextlib.contextmanager
    def transaction(self):
        """vx_lock ... vx_unlock: a call and the read of its result as one unit when several threads
        share the context (the two-call helpers below take it themselves)."""
        self._lib.vx_lock(self._h)
        try:
            yield self
        finally:
            self._lib.vx_unlock(self._h)

    def set_seed(self, seed: int):
        self.seed = seed & 0xFFFFFFFFFFFFFFFF
        self._check(self._lib.vx_set_seed(self._h, self.seed))

    def set_gradient_order(self, order2=None, order3=None):
        """libnoise calibration: permutations of the 2-D (24) / 3-D (12) gradient tables, None = canonical."""
        o2 = None if order2 is None else np.ascontiguousarray(order2, np.uint8)
        o3 = None if order3 is None else np.ascontiguousarray(order3, np.uint8)
        assert (o2 is None or o2.size == 24) and (o3 is None or o3.size == 12)
        self._check(self._lib.vx_set_gradient_order(self._h, _ptr(o2), _ptr(o3)))

    def cave_columns(self, area_xy):
        """(cave-zone columns, cave-noise voxels) per area from the seed alone: the cost hint for sharding."""
        xy = _xy(area_xy)
        cols = np.zeros(xy.shape[0], np.uint32)
        vox = np.zeros(xy.shape[0], np.uint32)
        self._check(self._lib.vx_cave_columns(self._h, xy.ctypes.data, xy.shape[0], cols.ctypes.data, vox.ctypes.data))
        return cols, vox

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.vx_set_stream(self._h, cuda_stream_ptr or None))

    def sync(self):
        self._check(self._lib.vx_sync(self._h))

    def set_timing(self, enabled: bool):
        self._check(self._lib.vx_set_timing(self._h, int(enabled)))

    def timings(self) -> dict:
        t = Timings()
        self._check(self._lib.vx_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    # -- generation / light
    def generate(self, area_xy, voxels_out=None, max_height_out=None, read=True):
        """AreaLoader::load_all_blocking for disk misses. Returns (voxels[n,32768], max_height[n,256])
        (None, None when read=False and no output buffers are given).  With deferred reads on, a call
        that asks for the heights only (read=False, max_height_out=pinned) returns before they have
        arrived: wait_reads()."""
        xy = _xy(area_xy)
        n = xy.shape[0]
        if read and voxels_out is None:
            voxels_out = np.empty((n, VOXELS_IN_AREA), np.uint8)
        if read and max_height_out is None:
            max_height_out = np.empty((n, COLUMNS_IN_AREA), np.uint8)
        self._check(self._lib.vx_generate(self._h, xy.ctypes.data, n, _ptr(voxels_out), _ptr(max_height_out)))
        return voxels_out, max_height_out

    def upload(self, area_xy, voxels, read_heights=True):
        xy = _xy(area_xy)
        n = xy.shape[0]
        voxels = np.ascontiguousarray(voxels, np.uint8).reshape(n, VOXELS_IN_AREA)
        mh = np.empty((n, COLUMNS_IN_AREA), np.uint8) if read_heights else None
        self._check(self._lib.vx_upload(self._h, xy.ctypes.data, n, voxels.ctypes.data, _ptr(mh)))
        return mh

    def read_areas(self, area_xy, voxels=True, heights=True):
        xy = _xy(area_xy)
        n = xy.shape[0]
        v = np.empty((n, VOXELS_IN_AREA), np.uint8) if voxels else None
        h = np.empty((n, COLUMNS_IN_AREA), np.uint8) if heights else None
        self._check(self._lib.vx_read_areas(self._h, xy.ctypes.data, n, _ptr(v), _ptr(h)))
        return v, h

    def unload(self, area_xy):
        xy = _xy(area_xy)
        self._check(self._lib.vx_unload(self._h, xy.ctypes.data, xy.shape[0]))

    def pool_info(self) -> dict:
        cap, committed, virt = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        self._check(self._lib.vx_pool_info(self._h, C.byref(cap), C.byref(committed), C.byref(virt)))
        return {"capacity_areas": int(cap.value), "committed_bytes": int(committed.value), "virtual_memory": bool(virt.value)}

    def resident_count(self) -> int:
        return int(self._lib.vx_resident_count(self._h))

    def column_heights(self, voxels) -> np.ndarray:
        """Area::update_all_column_heights on host data (n areas)."""
        voxels = np.ascontiguousarray(voxels, np.uint8).reshape(-1, VOXELS_IN_AREA)
        n = voxels.shape[0]
        mh = np.empty((n, COLUMNS_IN_AREA), np.uint8)
        self._check(self._lib.vx_column_heights(self._h, voxels.ctypes.data, n, mh.ctypes.data))
        return mh

    # -- mesh
    def mesh(self, area_xy) -> dict:
        """Renderer::load_all_blocking: mesh on the device; returns the sizes."""
        xy = _xy(area_xy)
        info = MeshInfo()
        self._check(self._lib.vx_mesh(self._h, xy.ctypes.data, xy.shape[0], C.byref(info)))
        self._mesh_n = xy.shape[0]
        return {k: int(getattr(info, k)) for k, _ in info._fields_}

    def mesh_read(self, info: dict, out=None):
        """Copy the last mesh to the host: (rec_first, face_first, recs, vertices, indices)."""
        n = self._mesh_n
        if out is None:
            out = (np.empty(n + 1, np.uint32), np.empty(n + 1, np.uint32),
                   np.empty(info["n_recs"], REC_DTYPE), np.empty(info["n_faces"] * 4, VERTEX_DTYPE),
                   np.empty(info["n_faces"] * 6, np.uint16))
        rf, ff, recs, verts, idx = out
        self._check(self._lib.vx_mesh_read(self._h, _ptr(rf), _ptr(ff), _ptr(recs), _ptr(verts), _ptr(idx)))
        return out

    def mesh_and_read(self, area_xy, compact: bool = False):
        """mesh() + mesh_read() / mesh_read_compact() under one lock -> (info, arrays)."""
        with self.transaction():
            info = self.mesh(area_xy)
            return info, (self.mesh_read_compact(info) if compact else self.mesh_read(info))

    def mesh_read_compact(self, info: dict, out=None):
        """The last mesh as 4-byte records: (rec_first u32[n+1], packed u32[n_recs]); see
        unpack_records / the C++ expander for the way back."""
        n = self._mesh_n
        own = out is None
        if own:
            out = (np.empty(n + 1, np.uint32), np.empty(info["n_recs"], np.uint32))
        rf, packed = out
        self._check(self._lib.vx_mesh_read_compact(self._h, _ptr(rf), _ptr(packed)))
        if self._deferred and own:
            self.wait_reads()  # buffers of our own (pageable): nothing to defer
        return out

    # -- Renderer.meshes on the device
    def set_mesh_retention(self, enabled: bool):
        self._check(self._lib.vx_set_mesh_retention(self._h, int(enabled)))

    def mesh_unload(self, area_xy):
        xy = _xy(area_xy)
        self._check(self._lib.vx_mesh_unload(self._h, xy.ctypes.data, xy.shape[0]))

    def read_face_masks(self, area_xy):
        """-> (face masks u8[n,32768], meshed u8[n]) of resident areas (mesh retention must be on)."""
        xy = _xy(area_xy)
        n = xy.shape[0]
        masks = np.empty((n, VOXELS_IN_AREA), np.uint8)
        meshed = np.empty(n, np.uint8)
        self._check(self._lib.vx_read_face_masks(self._h, xy.ctypes.data, n, masks.ctypes.data, meshed.ctypes.data))
        return masks, meshed

    @_locked
    def update_locations(self, xyz, read: bool = True):
        """Renderer::update_location for a batch of internal locations [k,3] -> info dict and, if
        read, (recs, vertices, indices); records with n_faces == 0 are removals."""
        loc = np.ascontiguousarray(xyz, np.uint32).reshape(-1, 3)
        info = MeshInfo()
        self._check(self._lib.vx_update_locations(self._h, loc.ctypes.data, loc.shape[0], C.byref(info)))
        d = {k: int(getattr(info, k)) for k, _ in info._fields_}
        if not read:
            return d
        recs = np.empty(d["n_recs"], REC_DTYPE)
        verts = np.empty(d["n_faces"] * 4, VERTEX_DTYPE)
        idx = np.empty(d["n_faces"] * 6, np.uint16)
        self._check(self._lib.vx_update_read(self._h, _ptr(recs) if recs.size else None,
                                             _ptr(verts) if verts.size else None, _ptr(idx) if idx.size else None))
        return d, recs, verts, idx

    @_locked
    def visible_zone(self, area_xy, view: View, read: bool = True):
        """Per-frame visibility over the resident meshes of a render zone -> dict and, if read,
        (codes u32[n_visible] = zone index << 15 | local voxel index, area_visible u8[n])."""
        xy = _xy(area_xy)
        info = VisibleInfo()
        self._check(self._lib.vx_visible_zone(self._h, xy.ctypes.data, xy.shape[0], C.byref(view), C.byref(info)))
        out = dict(n_visible=int(info.n_visible), n_faces=int(info.n_faces), n_areas=int(info.n_areas),
                   n_lamps=int(info.n_lamps), type_first=np.array(info.type_first[:], np.uint32),
                   lamps=np.array(info.lamps[: min(64, int(info.n_lamps))], np.uint32))
        if not read:
            return out
        idx = np.empty(out["n_visible"], np.uint32)
        flags = np.empty(xy.shape[0], np.uint8)
        self._check(self._lib.vx_visible_read(self._h, idx.ctypes.data if idx.size else None,
                                              flags.ctypes.data if flags.size else None))
        return out, idx, flags

    # -- edits
    def apply_edits(self, edits, want_dirty: bool = True):
        """World::set for a batch (array order). Returns the dirty areas as [k,2] u32, or None with
        want_dirty=False (the call then only enqueues; follow up with update_locations)."""
        e = np.ascontiguousarray(edits, dtype=EDIT_DTYPE)
        n = e.shape[0]
        if not want_dirty:
            self._check(self._lib.vx_apply_edits(self._h, e.ctypes.data, n, None, None))
            return None
        dirty = np.empty((max(3 * n, 1), 2), np.uint32)
        nd = C.c_int(0)
        self._check(self._lib.vx_apply_edits(self._h, e.ctypes.data, n, dirty.ctypes.data, C.byref(nd)))
        return dirty[: nd.value].copy()

    def explode(self, positions):
        """BombSimulator::handle_explosions for one tick -> (removed [k,3] u32, bombs [m,3] u32,
        dirty areas [d,2] u32); positions are Vec3 in player coordinates."""
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        removed = np.empty((max(1331 * n, 1), 3), np.uint32)
        bombs = np.empty((max(1331 * n, 1), 3), np.uint32)
        dirty = np.empty((max(9 * n, 1), 2), np.uint32)
        nr, nb, nd = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._lib.vx_explode(self._h, pos.ctypes.data, n, removed.ctypes.data, C.byref(nr),
                                         bombs.ctypes.data, C.byref(nb), dirty.ctypes.data, C.byref(nd)))
        return removed[: nr.value].copy(), bombs[: nb.value].copy(), dirty[: nd.value].copy()

    # -- water / falling voxels (ordered lists, applied one location after the other)
    def water_tick(self, check_xyz):
        """WaterSimulator::simulate_voxels over check_xyz [n,3] in this order -> (changed [k,4] u32:
        x, y, z, new voxel; next [m,3] u32: insertions into the next tick's check_locations)."""
        loc = np.ascontiguousarray(check_xyz, np.uint32).reshape(-1, 3)
        n = loc.shape[0]
        changed = np.zeros((max(4 * n, 1), 4), np.uint32)
        nxt = np.zeros((max(7 * n, 1), 3), np.uint32)
        nc, nn = C.c_int(0), C.c_int(0)
        st = self._lib.vx_water_tick(self._h, loc.ctypes.data, n, changed.ctypes.data, C.byref(nc), nxt.ctypes.data, C.byref(nn))
        self._check(st)
        return changed[: nc.value].copy(), nxt[: nn.value].copy()

    def falling_check(self, check_xyz, cap_spawned: int = 1 << 16, cap_next: int = 7 << 16):
        """FallingVoxelSimulator::update_voxels for each of check_xyz [n,3] in order -> (spawned [k,4]:
        x, y, z, voxel of every entity created; water_next [m,3])."""
        loc = np.ascontiguousarray(check_xyz, np.uint32).reshape(-1, 3)
        spawned = np.zeros((max(cap_spawned, 1), 4), np.uint32)
        nxt = np.zeros((max(cap_next, 1), 3), np.uint32)
        ns, nn = C.c_int(0), C.c_int(0)
        st = self._lib.vx_falling_check(self._h, loc.ctypes.data, loc.shape[0], spawned.ctypes.data, cap_spawned,
                                        C.byref(ns), nxt.ctypes.data, cap_next, C.byref(nn))
        self._check(st)
        return spawned[: ns.value].copy(), nxt[: nn.value].copy()

    def falling_land(self, positions, voxel_types):
        """The retain pass of simulate_falling over entities in Vec order (positions already advanced)
        -> (keep u8[n], placed [k,4] u32, water_next [m,3] u32)."""
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        types = np.ascontiguousarray(voxel_types, np.uint8).reshape(-1)
        n = pos.shape[0]
        assert types.shape[0] == n
        keep = np.zeros(max(n, 1), np.uint8)
        placed = np.zeros((max(n, 1), 4), np.uint32)
        nxt = np.zeros((max(7 * n, 1), 3), np.uint32)
        npl, nn = C.c_int(0), C.c_int(0)
        st = self._lib.vx_falling_land(self._h, pos.ctypes.data, types.ctypes.data, n, keep.ctypes.data, placed.ctypes.data,
                                       C.byref(npl), nxt.ctypes.data, C.byref(nn))
        self._check(st)
        return keep[:n].copy(), placed[: npl.value].copy(), nxt[: nn.value].copy()

    def heightmap(self, area_xy, cam_ax: int, cam_ay: int, rgba=None) -> np.ndarray:
        xy = _xy(area_xy)
        if rgba is None:
            rgba = np.empty((512, 512, 4), np.uint8)
        self._check(self._lib.vx_heightmap(self._h, xy.ctypes.data, xy.shape[0], cam_ax, cam_ay, rgba.ctypes.data))
        return rgba

    # -- area save files
    @_locked
    def encode_areas(self, area_xy, read: bool = True, out=None):
        """store_blocking for a batch of resident areas -> (files u8[total], offsets u64[n+1]);
        file i = files[offsets[i]:offsets[i+1]].  read=False leaves them on the device and returns
        the total size."""
        xy = _xy(area_xy)
        total = C.c_uint64(0)
        self._check(self._lib.vx_encode_areas(self._h, xy.ctypes.data, xy.shape[0], C.byref(total)))
        if not read:
            return int(total.value)
        files = out if out is not None else np.empty(int(total.value), np.uint8)
        offsets = np.empty(xy.shape[0] + 1, np.uint64)
        self._check(self._lib.vx_encode_read(self._h, files.ctypes.data if files.size else None, offsets.ctypes.data))
        if self._deferred:
            self.wait_reads()  # this form returns the files: nothing to defer
        return files[: int(total.value)], offsets

    def encode_read(self, n_areas: int, total: int, out=None, offsets_out=None):
        """The files of the last encode_areas(read=False) -> (files u8[total], offsets u64[n+1]).  The copy
        waits for the encoder only: a mesh() issued in between runs while the files cross to the host.
        With deferred reads on, `out` (and `offsets_out`, else the offsets are not fetched) must be pinned
        and are valid after wait_reads()."""
        files = out if out is not None else np.empty(int(total), np.uint8)
        offsets = offsets_out
        if offsets is None and (out is None or not self._deferred):
            offsets = np.empty(n_areas + 1, np.uint64)
        self._check(self._lib.vx_encode_read(self._h, files.ctypes.data if files.size else None, _ptr(offsets)))
        if self._deferred and out is None:
            self.wait_reads()  # a buffer of our own (pageable): nothing to defer
        return files[: int(total)], offsets

    def set_deferred_reads(self, enabled: bool):
        """vx_set_deferred_reads: encode_read / mesh_read_compact return before their copies are done."""
        self._check(self._lib.vx_set_deferred_reads(self._h, int(bool(enabled))))
        self._deferred = bool(enabled)

    def wait_reads(self):
        self._check(self._lib.vx_wait_reads(self._h))

    def decode_areas(self, area_xy, files, offsets, read_heights: bool = True):
        """load_blocking for a batch of files -> (ok u8[n], max_height u8[n,256] or None); areas with
        ok == 0 are not resident afterwards (the caller generates them)."""
        xy = _xy(area_xy)
        n = xy.shape[0]
        files = np.ascontiguousarray(files, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        assert offsets.shape[0] == n + 1
        ok = np.zeros(n, np.uint8)
        mh = np.empty((n, COLUMNS_IN_AREA), np.uint8) if read_heights else None
        self._check(self._lib.vx_decode_areas(self._h, xy.ctypes.data, n, files.ctypes.data if files.size else None,
                                              offsets.ctypes.data, ok.ctypes.data, _ptr(mh)))
        return ok, mh

    @_locked
    def visible(self, view: View, read: bool = True):
        """Per-frame visibility over the mesh of the last mesh() call.
        -> dict(n_visible, n_faces, n_areas, n_lamps, type_first[28], lamps[<=64]) and, if read,
        (rec_index u32[n_visible], area_visible u8[n areas])."""
        info = VisibleInfo()
        self._check(self._lib.vx_visible(self._h, C.byref(view), C.byref(info)))
        out = dict(n_visible=int(info.n_visible), n_faces=int(info.n_faces), n_areas=int(info.n_areas),
                   n_lamps=int(info.n_lamps), type_first=np.array(info.type_first[:], np.uint32),
                   lamps=np.array(info.lamps[: min(64, int(info.n_lamps))], np.uint32))
        if not read:
            return out
        idx = np.empty(out["n_visible"], np.uint32)
        flags = np.empty(self._mesh_n, np.uint8)
        self._check(self._lib.vx_visible_read(self._h, idx.ctypes.data if idx.size else None,
                                              flags.ctypes.data if flags.size else None))
        return out, idx, flags

    def light(self, area_xy):
        xy = _xy(area_xy)
        n = xy.shape[0]
        out = np.empty((n, VOXELS_IN_AREA), np.uint8)
        it = C.c_int(0)
        self._check(self._lib.vx_light(self._h, xy.ctypes.data, n, out.ctypes.data, C.byref(it)))
        return out, it.value

    def fp64_rate(self):
        g, ms = C.c_double(0), C.c_float(0)
        self._check(self._lib.vx_diag_fp64_rate(self._h, C.byref(g), C.byref(ms)))
        return g.value, ms.value
