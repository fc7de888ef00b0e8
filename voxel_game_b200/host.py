"""ctypes plumbing for libvoxel_host.so: the C++ host mirror of the reference's Renderer
(voxel_game_b200/host/voxel_game.hpp) behind a few C entry points (host/host_abi.cpp).  Used by the
tests and by bench.py to time the path a compiled-code host takes: device -> compact records ->
per-voxel Mesh maps on the host.  Nothing here computes meshes on the CPU from voxels: the records
come from the CUDA kernels, the host only rebuilds vertices from them."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .binding import REC_DTYPE, VERTEX_DTYPE, VxError, _xy, load

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libvoxel_host.so")
HOST_SYMBOLS = [
    "vxh_expand_records", "vxh_renderer_create", "vxh_renderer_destroy", "vxh_last_error",
    "vxh_renderer_load_all_blocking", "vxh_renderer_update_locations", "vxh_renderer_unload_area",
    "vxh_renderer_clear", "vxh_renderer_face_count", "vxh_renderer_voxel_count", "vxh_renderer_area_count",
    "vxh_renderer_area_face_counts", "vxh_renderer_voxel_mesh",
]
_lib = None


def build_host() -> str:
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "host")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvoxel_host.so failed:\n" + r.stdout + r.stderr)
    return HOST_LIB_PATH


def load_host():
    global _lib
    if _lib is not None:
        return _lib
    load()  # libvoxel_cuda.so first (the host library links against it)
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
    L = C.CDLL(HOST_LIB_PATH)
    vp = C.c_void_p
    L.vxh_expand_records.restype = None
    L.vxh_expand_records.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
    L.vxh_renderer_create.restype = vp
    L.vxh_renderer_create.argtypes = [vp]
    L.vxh_renderer_destroy.restype = None
    L.vxh_renderer_destroy.argtypes = [vp]
    L.vxh_last_error.restype = C.c_char_p
    L.vxh_last_error.argtypes = [vp]
    L.vxh_renderer_load_all_blocking.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.vxh_renderer_update_locations.argtypes = [vp, vp, C.c_int]
    L.vxh_renderer_unload_area.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.vxh_renderer_clear.restype = None
    L.vxh_renderer_clear.argtypes = [vp]
    for name in ("vxh_renderer_face_count", "vxh_renderer_voxel_count", "vxh_renderer_area_count"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [vp]
    L.vxh_renderer_area_face_counts.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.POINTER(C.c_uint32)]
    L.vxh_renderer_voxel_mesh.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    _lib = L
    return L


def expand_records(area_xy, rec_first, packed, threads: int = 0, out=None):
    """Compact records -> (face_first, recs, vertices, indices), the stream vx_mesh_read delivers,
    rebuilt on the host by the C++ expander."""
    L = load_host()
    xy = _xy(area_xy)
    n = xy.shape[0]
    rec_first = np.ascontiguousarray(rec_first, np.uint32)
    packed = np.ascontiguousarray(packed, np.uint32)
    mask = (packed >> 20) & 0x3F
    n_faces = int(sum(int(((mask >> b) & 1).sum()) for b in range(6)))
    if out is None:
        out = (np.empty(n + 1, np.uint32), np.empty(len(packed), REC_DTYPE), np.empty(n_faces * 4, VERTEX_DTYPE),
               np.empty(n_faces * 6, np.uint16))
    ff, recs, verts, idx = out
    L.vxh_expand_records(xy.ctypes.data, n, rec_first.ctypes.data, packed.ctypes.data, ff.ctypes.data,
                         recs.ctypes.data, verts.ctypes.data, idx.ctypes.data, int(threads))
    return out


class HostRenderer:
    """The C++ mirror of `Renderer` (renderer.rs:97-322) on top of a Context."""

    def __init__(self, ctx):
        self._lib = load_host()
        self._ctx = ctx
        self._h = self._lib.vxh_renderer_create(ctx._h)
        if not self._h:
            raise VxError(-3, "vxh_renderer_create failed")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vxh_renderer_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st):
        if st != 0:
            raise VxError(st, self._lib.vxh_last_error(self._h).decode())

    def load_all_blocking(self, area_xy, transport: str = "compact", threads: int = 0) -> float:
        """-> wall-clock milliseconds of the whole call (device mesh + transfer + host Mesh maps)."""
        xy = _xy(area_xy)
        ms = (C.c_double * 2)()
        self._check(self._lib.vxh_renderer_load_all_blocking(self._h, xy.ctypes.data, xy.shape[0],
                                                             1 if transport == "compact" else 0, int(threads), ms))
        return float(ms[0])

    def update_locations(self, xyz):
        loc = np.ascontiguousarray(xyz, np.uint32).reshape(-1, 3)
        self._check(self._lib.vxh_renderer_update_locations(self._h, loc.ctypes.data, loc.shape[0]))

    def unload_area(self, ax: int, ay: int):
        self._check(self._lib.vxh_renderer_unload_area(self._h, int(ax), int(ay)))

    def clear(self):
        self._lib.vxh_renderer_clear(self._h)

    def face_count(self) -> int:
        return int(self._lib.vxh_renderer_face_count(self._h))

    def voxel_count(self) -> int:
        return int(self._lib.vxh_renderer_voxel_count(self._h))

    def area_count(self) -> int:
        return int(self._lib.vxh_renderer_area_count(self._h))

    def area_face_counts(self, ax: int, ay: int):
        """-> (face count per voxel u8[32768], lamps) of the area's RenderArea, or None."""
        out = np.empty(32768, np.uint8)
        lamps = C.c_uint32(0)
        st = self._lib.vxh_renderer_area_face_counts(self._h, int(ax), int(ay), out.ctypes.data, C.byref(lamps))
        if st == -1:
            return None
        assert st == 0, "a Mesh whose vertex / index count does not match its face count"
        return out, int(lamps.value)

    def voxel_mesh(self, gx: int, gy: int, gz: int):
        verts = np.empty(24, VERTEX_DTYPE)
        idx = np.empty(36, np.uint16)
        nf = self._lib.vxh_renderer_voxel_mesh(self._h, int(gx), int(gy), int(gz), verts.ctypes.data, idx.ctypes.data)
        return verts[: 4 * nf].copy(), idx[: 6 * nf].copy()
