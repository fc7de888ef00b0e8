"""ctypes binding of the CPU ORACLE (oracle/build/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (voxel_game_b200) never imports it.
See oracle/vg_oracle.h for the parity status ("parity unpinned" at the libnoise boundary).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")

AREA_SIZE = 16
AREA_HEIGHT = 128
VOXELS_IN_AREA = AREA_SIZE * AREA_SIZE * AREA_HEIGHT
COLUMNS_IN_AREA = AREA_SIZE * AREA_SIZE
LOCATION_OFFSET = 1_000_000

VOXEL_NAMES = [
    "None", "Cobblestone", "Sand", "Grass", "Wood", "Leaves", "Brick", "Dirt", "Boards", "Stone",
    "Clay", "Lamp", "Trampoline", "Cactus", "WaterSource", "WaterDown", "Water1", "Water2",
    "Water3", "Water4", "StoneBrick", "StonePillar", "Snow", "Ice", "Bomb", "ActiveBomb", "Glass",
]
V = {n: i for i, n in enumerate(VOXEL_NAMES)}
TREE_NAMES = ["None", "Short", "Tall", "TallCactus", "DeadTree", "HugeTree", "ShortCactus", "Bush"]
T = {n: i for i, n in enumerate(TREE_NAMES)}

REC_DTYPE = np.dtype([("gx", "<u4"), ("gy", "<u4"), ("packed", "<u4"), ("first_face", "<u4")])
VERTEX_DTYPE = np.dtype(
    [("position", "<f4", 3), ("uv", "<f4", 2), ("color", "u1", 4), ("normal", "<f4", 4)]
)
assert VERTEX_DTYPE.itemsize == 40 and REC_DTYPE.itemsize == 16


def use_native_build() -> bool:
    """For the CPU-baseline timing legs of bench.py: rebuild the oracle with -march=native ON THE
    MACHINE THAT RUNS THE TIMING (the library that travels with the repo is -march=x86-64-v2 so that
    it runs anywhere) and switch this module to it.  Returns False (and changes nothing) if gcc is
    not available there."""
    global _LIB_PATH, _lib
    native = os.path.join(_HERE, "build", "liboracle_native.so")
    try:
        r = subprocess.run(["make", "-C", _HERE, "-B", "MARCH=native", "OUT=build/liboracle_native.so"],
                           capture_output=True)
    except OSError:
        return False
    if r.returncode != 0 or not os.path.exists(native):
        return False
    _LIB_PATH = native
    _lib = None
    return True


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile). Returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


class GenStats(C.Structure):
    _fields_ = [
        ("simplex2_evals", C.c_uint64),
        ("simplex3_evals", C.c_uint64),
        ("cave_voxels", C.c_uint64),
        ("cave_columns", C.c_uint64),
        ("trees_placed", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class _World(C.Structure):
    _fields_ = [
        ("ax0", C.c_uint32), ("ay0", C.c_uint32), ("nx", C.c_uint32), ("ny", C.c_uint32),
        ("voxels", C.c_void_p), ("max_height", C.c_void_p), ("face_mask", C.c_void_p),
    ]


class View(C.Structure):
    """vgo_view: the camera state the per-frame visibility pass depends on."""
    _fields_ = [
        ("cam_pos", C.c_float * 3), ("cam_target_z", C.c_float), ("look", C.c_float * 3),
        ("render_size", C.c_uint32), ("head_in_water", C.c_uint32), ("show_map", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u32p, u16p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)
        L.vgo_hash_world_name.restype = C.c_uint64
        L.vgo_hash_world_name.argtypes = [C.c_char_p]
        L.vgo_siphash.restype = C.c_uint64
        L.vgo_siphash.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_char_p, C.c_size_t]
        L.vgo_chacha12_words.argtypes = [C.c_uint64, u32p, C.c_int]
        L.vgo_chacha_block_raw.argtypes = [u32p, C.c_uint64, C.c_int, u32p]
        L.vgo_simplex_new.argtypes = [C.c_void_p, C.c_uint64]
        L.vgo_simplex2.restype = C.c_double
        L.vgo_simplex2.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.vgo_simplex3.restype = C.c_double
        L.vgo_simplex3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.vgo_grad2_table.argtypes = [C.POINTER(C.c_double)]
        L.vgo_grad3_table.argtypes = [C.POINTER(C.c_double)]
        L.vgo_simplex_perm.argtypes = [C.c_void_p, u8p]
        L.vgo_fbm_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.vgo_fbm2.restype = C.c_double
        L.vgo_fbm2.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.vgo_fbm3.restype = C.c_double
        L.vgo_fbm3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.vgo_set_noise_variant.argtypes = [C.c_uint, u8p, u8p]
        L.vgo_get_noise_variant.restype = C.c_uint
        L.vgo_op_counts.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        L.vgo_generator_new.restype = C.c_void_p
        L.vgo_generator_new.argtypes = [C.c_uint64]
        L.vgo_generator_free.argtypes = [C.c_void_p]
        L.vgo_generate_area.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u8p, u8p, C.POINTER(GenStats)]
        L.vgo_sample_column.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
        L.vgo_generate_areas.restype = C.c_int
        L.vgo_generate_areas.argtypes = [C.c_uint64, u32p, C.c_int, u8p, u8p, C.c_int, C.c_int, C.POINTER(GenStats)]
        L.vgo_should_generate_tree.restype = C.c_int
        L.vgo_should_generate_tree.argtypes = [C.c_uint8, C.c_uint64] + [C.c_uint32] * 5
        L.vgo_can_generate_tree.restype = C.c_int
        L.vgo_can_generate_tree.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.vgo_generate_tree.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.vgo_combine_seed.restype = C.c_uint64
        L.vgo_combine_seed.argtypes = [C.c_uint64] + [C.c_uint32] * 5
        L.vgo_split_mix64.restype = C.c_uint64
        L.vgo_split_mix64.argtypes = [C.c_uint64]
        L.vgo_update_all_column_heights.argtypes = [u8p, u8p]
        L.vgo_area_new.argtypes = [u8p, u8p]
        L.vgo_area_set.argtypes = [u8p, u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8]
        L.vgo_height_map.argtypes = [u32p, C.c_int, u8p, C.c_uint32, C.c_uint32, u8p]
        L.vgo_should_generate_face.restype = C.c_int
        L.vgo_should_generate_face.argtypes = [C.c_uint8, C.c_uint8]
        L.vgo_should_generate_top_face.restype = C.c_int
        L.vgo_should_generate_top_face.argtypes = [C.c_uint8, C.c_uint8]
        L.vgo_mesh_area.restype = C.c_long
        L.vgo_mesh_area.argtypes = [C.POINTER(_World), C.c_uint32, C.c_uint32]
        L.vgo_emit_area.restype = C.c_long
        L.vgo_emit_area.argtypes = [C.POINTER(_World), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.POINTER(C.c_long), C.c_void_p, C.c_void_p]
        L.vgo_mesh_areas.restype = C.c_long
        L.vgo_mesh_areas.argtypes = [C.POINTER(_World), u32p, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.vgo_apply_edit.restype = C.c_int
        L.vgo_apply_edit.argtypes = [C.POINTER(_World), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8]
        L.vgo_update_location.restype = None
        L.vgo_update_location.argtypes = [C.POINTER(_World), C.c_uint32, C.c_uint32, C.c_uint32]
        L.vgo_light_flood_world.argtypes = [C.POINTER(_World), u8p]
        L.vgo_area_render_distance.restype = C.c_float
        L.vgo_area_render_distance.argtypes = [C.POINTER(C.c_float), C.c_uint32]
        L.vgo_is_area_visible.restype = C.c_int
        L.vgo_is_area_visible.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(View), C.c_float]
        L.vgo_is_voxel_visible.restype = C.c_int
        L.vgo_is_voxel_visible.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float),
                                           C.POINTER(C.c_float), C.c_float]
        L.vgo_filter_visible.restype = C.c_long
        L.vgo_filter_visible.argtypes = [C.c_void_p, u32p, u32p, C.c_int, C.POINTER(View), u8p, u32p, u32p,
                                         C.POINTER(C.c_long), u32p, C.POINTER(C.c_long)]
        L.vgo_bincode_area.argtypes = [u8p, u8p]
        L.vgo_lz4_decompress_block.restype = C.c_long
        L.vgo_lz4_decompress_block.argtypes = [u8p, C.c_long, u8p, C.c_long]
        L.vgo_lz4_compress_block.restype = C.c_long
        L.vgo_lz4_compress_block.argtypes = [u8p, C.c_long, u8p]
        L.vgo_area_file_encode.restype = C.c_long
        L.vgo_area_file_encode.argtypes = [u8p, u8p]
        L.vgo_area_files_encode.restype = C.c_long
        L.vgo_area_files_encode.argtypes = [u8p, C.c_int, C.c_int, u8p, C.POINTER(C.c_long)]
        L.vgo_area_file_decode.restype = C.c_int
        L.vgo_area_file_decode.argtypes = [u8p, C.c_long, u8p]
        lp = C.POINTER(C.c_long)
        L.vgo_water_tick.restype = C.c_int
        L.vgo_water_tick.argtypes = [C.POINTER(_World), u32p, C.c_int, u32p, lp, u32p, lp]
        L.vgo_falling_check.restype = C.c_int
        L.vgo_falling_check.argtypes = [C.POINTER(_World), u32p, C.c_int, u32p, C.c_long, lp, u32p, C.c_long, lp]
        L.vgo_falling_land.restype = C.c_int
        L.vgo_falling_land.argtypes = [C.POINTER(_World), C.POINTER(C.c_float), u8p, C.c_int, u8p, u32p, lp, u32p, lp]
        L.vgo_explode.restype = C.c_int
        L.vgo_explode.argtypes = [C.POINTER(_World), C.POINTER(C.c_float), C.c_int, u32p, C.POINTER(C.c_long),
                                  u32p, C.POINTER(C.c_long)]
        L.vgo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def hash_world_name(name: str) -> int:
    return int(lib().vgo_hash_world_name(name.encode("utf-8")))


VAR_FBM_RUNNING_FREQ, VAR_TIE2_OTHER, VAR_TIE3_OTHER, VAR_SKEW_EXACT = 1, 2, 4, 8


def set_noise_variant(flags: int = 0, order2=None, order3=None):
    """Flip the unanchored libnoise choices (vg_oracle.h VGO_VAR_*; gradient-table permutations).
    Global to the library: callers reset it with set_noise_variant() when they are done."""
    o2 = None if order2 is None else np.ascontiguousarray(order2, np.uint8)
    o3 = None if order3 is None else np.ascontiguousarray(order3, np.uint8)
    assert (o2 is None or o2.size == 24) and (o3 is None or o3.size == 12)
    lib().vgo_set_noise_variant(int(flags), None if o2 is None else _p(o2), None if o3 is None else _p(o3))


def noise_variant() -> int:
    return int(lib().vgo_get_noise_variant())


def use_counting_build() -> bool:
    """Switch this module to a -DVGO_COUNT_OPS build of the oracle (the "counting scalar" of
    SURVEY.md 8d: every f64 add / multiply of the noise arithmetic goes through a counter)."""
    global _LIB_PATH, _lib
    counting = os.path.join(_HERE, "build", "liboracle_count.so")
    r = subprocess.run(["make", "-C", _HERE, "-B", "EXTRA=-DVGO_COUNT_OPS", "OUT=build/liboracle_count.so"],
                       capture_output=True)
    if r.returncode != 0 or not os.path.exists(counting):
        return False
    _LIB_PATH = counting
    _lib = None
    return True


def op_counts(reset: bool = True):
    """(f64 adds, f64 multiplies) of the noise arithmetic on this thread since the last reset."""
    a, m = C.c_uint64(0), C.c_uint64(0)
    lib().vgo_op_counts(C.byref(a), C.byref(m), int(reset))
    return int(a.value), int(m.value)


def max_threads() -> int:
    return int(lib().vgo_max_threads())


def region_xy(ax0: int, ay0: int, nx: int, ny: int) -> np.ndarray:
    """Area coordinates of a dense grid, grid index = ix + nx*iy (the vgo_world layout)."""
    ix, iy = np.meshgrid(np.arange(nx, dtype=np.uint32), np.arange(ny, dtype=np.uint32))
    xy = np.stack([ix.ravel() + np.uint32(ax0), iy.ravel() + np.uint32(ay0)], axis=1)
    return np.ascontiguousarray(xy, dtype=np.uint32)


class Simplex:
    """libnoise Simplex::new(seed) (permutation table) with .sample2/.sample3."""

    def __init__(self, seed: int):
        self._buf = (C.c_uint8 * 512)()
        lib().vgo_simplex_new(C.addressof(self._buf), seed & 0xFFFFFFFFFFFFFFFF)

    @property
    def perm(self) -> np.ndarray:
        return np.frombuffer(self._buf, dtype=np.uint8).copy()

    def sample2(self, x, y) -> float:
        return float(lib().vgo_simplex2(C.addressof(self._buf), x, y))

    def sample3(self, x, y, z) -> float:
        return float(lib().vgo_simplex3(C.addressof(self._buf), x, y, z))


class _Fbm(C.Structure):
    _fields_ = [("src", C.c_void_p), ("octaves", C.c_int), ("frequency", C.c_double),
                ("lacunarity", C.c_double), ("persistence", C.c_double),
                ("normalization_factor", C.c_double)]


class Fbm:
    """Simplex::new(seed).fbm(octaves, frequency, lacunarity, persistence)."""

    def __init__(self, simplex: Simplex, octaves, frequency, lacunarity, persistence):
        self._s = simplex
        self._f = _Fbm()
        lib().vgo_fbm_new(C.byref(self._f), C.addressof(simplex._buf), octaves, frequency, lacunarity, persistence)

    @property
    def normalization_factor(self):
        return float(self._f.normalization_factor)

    def sample2(self, x, y):
        return float(lib().vgo_fbm2(C.byref(self._f), x, y))

    def sample3(self, x, y, z):
        return float(lib().vgo_fbm3(C.byref(self._f), x, y, z))


def grad2_table() -> np.ndarray:
    out = np.zeros((24, 2))
    lib().vgo_grad2_table(_p(out, C.c_double))
    return out


def grad3_table() -> np.ndarray:
    out = np.zeros((12, 3))
    lib().vgo_grad3_table(_p(out, C.c_double))
    return out


def chacha12_words(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.uint32)
    lib().vgo_chacha12_words(seed & 0xFFFFFFFFFFFFFFFF, _p(out, C.c_uint32), n)
    return out


def chacha_block(key_words, counter: int, double_rounds: int) -> np.ndarray:
    key = np.ascontiguousarray(key_words, np.uint32)
    out = np.zeros(16, np.uint32)
    lib().vgo_chacha_block_raw(_p(key, C.c_uint32), counter, double_rounds, _p(out, C.c_uint32))
    return out


def siphash(c_rounds, d_rounds, k0, k1, data: bytes) -> int:
    return int(lib().vgo_siphash(c_rounds, d_rounds, k0, k1, data, len(data)))


class Generator:
    """AreaGenerator (generator.rs:37-146) for one seed."""

    def __init__(self, seed: int):
        self.seed = seed & 0xFFFFFFFFFFFFFFFF
        self._g = lib().vgo_generator_new(self.seed)

    def __del__(self):
        if getattr(self, "_g", None):
            lib().vgo_generator_free(self._g)
            self._g = None

    def generate_area(self, ax: int, ay: int):
        vox = np.empty(VOXELS_IN_AREA, np.uint8)
        mh = np.empty(COLUMNS_IN_AREA, np.uint8)
        st = GenStats()
        lib().vgo_generate_area(self._g, ax, ay, _p(vox), _p(mh), C.byref(st))
        return vox, mh, st.as_dict()

    def sample_column(self, ax, ay, x, y):
        out = np.zeros(4, np.uint32)
        lib().vgo_sample_column(self._g, ax, ay, x, y, _p(out, C.c_uint32))
        return tuple(int(v) for v in out)


def generate_areas(seed: int, area_xy: np.ndarray, rebuild_tables_per_area=False, threads=0):
    """AreaLoader::load_all_blocking with every area generated. Returns voxels[n,32768],
    max_height[n,256], stats dict, threads used."""
    area_xy = np.ascontiguousarray(area_xy, dtype=np.uint32).reshape(-1, 2)
    n = area_xy.shape[0]
    vox = np.empty((n, VOXELS_IN_AREA), np.uint8)
    mh = np.empty((n, COLUMNS_IN_AREA), np.uint8)
    st = GenStats()
    used = lib().vgo_generate_areas(seed & 0xFFFFFFFFFFFFFFFF, _p(area_xy, C.c_uint32), n, _p(vox),
                                    _p(mh), int(bool(rebuild_tables_per_area)), int(threads),
                                    C.byref(st))
    return vox, mh, st.as_dict(), int(used)


def update_all_column_heights(vox: np.ndarray) -> np.ndarray:
    vox = np.ascontiguousarray(vox, np.uint8)
    mh = np.empty(COLUMNS_IN_AREA, np.uint8)
    lib().vgo_update_all_column_heights(_p(vox), _p(mh))
    return mh


class Area:
    """model::area::Area (area.rs:11-201), voxels + max_height only."""

    def __init__(self):
        self.voxels = np.empty(VOXELS_IN_AREA, np.uint8)
        self.max_height = np.empty(COLUMNS_IN_AREA, np.uint8)
        lib().vgo_area_new(_p(self.voxels), _p(self.max_height))

    def set(self, x, y, z, voxel):
        lib().vgo_area_set(_p(self.voxels), _p(self.max_height), x, y, z, voxel)

    def set_without_updating_max_height(self, x, y, z, voxel):
        self.voxels[x + 16 * y + 256 * z] = voxel

    def get(self, x, y, z):
        return int(self.voxels[x + 16 * y + 256 * z])

    def sample_height(self, x, y):
        return int(self.max_height[x + 16 * y])

    def update_all_column_heights(self):
        lib().vgo_update_all_column_heights(_p(self.voxels), _p(self.max_height))


class World:
    """A dense loaded grid of areas + Renderer.meshes as per-voxel face masks."""

    def __init__(self, ax0, ay0, nx, ny, voxels, max_height):
        self.ax0, self.ay0, self.nx, self.ny = ax0, ay0, nx, ny
        self.voxels = np.ascontiguousarray(voxels, np.uint8).reshape(nx * ny, VOXELS_IN_AREA)
        self.max_height = np.ascontiguousarray(max_height, np.uint8).reshape(nx * ny, COLUMNS_IN_AREA)
        self.face_mask = np.zeros_like(self.voxels)
        self._w = _World(ax0, ay0, nx, ny, self.voxels.ctypes.data, self.max_height.ctypes.data,
                         self.face_mask.ctypes.data)

    @classmethod
    def generate(cls, seed, ax0, ay0, nx, ny, threads=0):
        vox, mh, _, _ = generate_areas(seed, region_xy(ax0, ay0, nx, ny), threads=threads)
        return cls(ax0, ay0, nx, ny, vox, mh)

    def slot(self, ax, ay):
        return (ax - self.ax0) + self.nx * (ay - self.ay0)

    def mesh_area(self, ax, ay) -> int:
        """Renderer::load_full_area. Returns the face count."""
        r = lib().vgo_mesh_area(C.byref(self._w), ax, ay)
        if r < 0:
            raise ValueError("area or one of its neighbours is not loaded")
        return int(r)

    def mesh_areas(self, area_xy, threads=1):
        """Renderer::load_all_blocking over a list (masks + emission). Returns (faces, recs)."""
        area_xy = np.ascontiguousarray(area_xy, np.uint32).reshape(-1, 2)
        nr = C.c_long(0)
        nf = lib().vgo_mesh_areas(C.byref(self._w), _p(area_xy, C.c_uint32), area_xy.shape[0],
                                  int(threads), C.byref(nr))
        if nf < 0:
            raise ValueError("an area or one of its neighbours is not loaded")
        return int(nf), int(nr.value)

    def emit_area(self, ax, ay, first_face_base=0):
        """Records, vertices (structured, 4/face) and u16 indices (6/face) of a meshed area."""
        fm = self.face_mask[self.slot(ax, ay)]
        n_recs = int(np.count_nonzero(fm))
        n_faces = int(np.unpackbits(fm).sum())
        recs = np.zeros(n_recs, REC_DTYPE)
        verts = np.zeros(n_faces * 4, VERTEX_DTYPE)
        idx = np.zeros(n_faces * 6, np.uint16)
        nr = C.c_long(0)
        nf = lib().vgo_emit_area(C.byref(self._w), ax, ay, first_face_base, recs.ctypes.data,
                                 C.byref(nr), verts.ctypes.data, idx.ctypes.data)
        assert nf == n_faces and nr.value == n_recs
        return recs, verts, idx

    def apply_edit(self, gx, gy, gz, voxel):
        """world.set + renderer.update_location."""
        if lib().vgo_apply_edit(C.byref(self._w), gx, gy, gz, voxel) != 0:
            raise ValueError("edit outside the loaded grid")

    def update_location(self, gx, gy, gz):
        """renderer.update_location alone."""
        lib().vgo_update_location(C.byref(self._w), int(gx), int(gy), int(gz))

    def area_xy(self) -> np.ndarray:
        return region_xy(self.ax0, self.ay0, self.nx, self.ny)

    def records(self, area_xy):
        """The current meshes (face masks) of the listed areas as a record stream in index order:
        (rec_first u32[n+1], recs) -- what Renderer.meshes holds, flattened."""
        area_xy = np.ascontiguousarray(area_xy, np.uint32).reshape(-1, 2)
        rec_first = np.zeros(len(area_xy) + 1, np.uint32)
        out, base = [], 0
        for i, (ax, ay) in enumerate(area_xy):
            recs, _, _ = self.emit_area(int(ax), int(ay), first_face_base=base)
            out.append(recs)
            base += int(np.unpackbits(self.face_mask[self.slot(int(ax), int(ay))]).sum())
            rec_first[i + 1] = rec_first[i] + len(recs)
        return rec_first, (np.concatenate(out) if out else np.zeros(0, REC_DTYPE))

    def explode(self, positions):
        """BombSimulator::handle_explosions for one tick -> (removed [k,3] u32, bombs [m,3] u32), internal
        coordinates in the reference's visiting order."""
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        n = len(pos)
        removed = np.zeros((max(1331 * n, 1), 3), np.uint32)
        bombs = np.zeros((max(1331 * n, 1), 3), np.uint32)
        nr, nb = C.c_long(0), C.c_long(0)
        if lib().vgo_explode(C.byref(self._w), pos.ctypes.data_as(C.POINTER(C.c_float)), n, _p(removed, C.c_uint32),
                             C.byref(nr), _p(bombs, C.c_uint32), C.byref(nb)) != 0:
            raise ValueError("explosion reaches outside the loaded grid")
        return removed[: nr.value].copy(), bombs[: nb.value].copy()

    def water_tick(self, check_xyz):
        """WaterSimulator::simulate_voxels over an ORDERED list -> (changed [k,4], next [m,3])."""
        loc = np.ascontiguousarray(check_xyz, np.uint32).reshape(-1, 3)
        n = loc.shape[0]
        changed = np.zeros((max(4 * n, 1), 4), np.uint32)
        nxt = np.zeros((max(7 * n, 1), 3), np.uint32)
        nc, nn = C.c_long(0), C.c_long(0)
        if lib().vgo_water_tick(C.byref(self._w), _p(loc, C.c_uint32), n, _p(changed, C.c_uint32), C.byref(nc),
                                _p(nxt, C.c_uint32), C.byref(nn)) != 0:
            raise ValueError("water tick reaches outside the loaded grid")
        return changed[: nc.value].copy(), nxt[: nn.value].copy()

    def falling_check(self, check_xyz, cap_spawned=1 << 16, cap_next=7 << 16):
        loc = np.ascontiguousarray(check_xyz, np.uint32).reshape(-1, 3)
        spawned = np.zeros((cap_spawned, 4), np.uint32)
        nxt = np.zeros((cap_next, 3), np.uint32)
        ns, nn = C.c_long(0), C.c_long(0)
        rc = lib().vgo_falling_check(C.byref(self._w), _p(loc, C.c_uint32), loc.shape[0], _p(spawned, C.c_uint32), cap_spawned,
                                     C.byref(ns), _p(nxt, C.c_uint32), cap_next, C.byref(nn))
        if rc != 0:
            raise ValueError(f"falling check failed ({rc})")
        return spawned[: ns.value].copy(), nxt[: nn.value].copy()

    def falling_land(self, positions, voxel_types):
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        types = np.ascontiguousarray(voxel_types, np.uint8).reshape(-1)
        n = pos.shape[0]
        keep = np.zeros(max(n, 1), np.uint8)
        placed = np.zeros((max(n, 1), 4), np.uint32)
        nxt = np.zeros((max(7 * n, 1), 3), np.uint32)
        npl, nn = C.c_long(0), C.c_long(0)
        if lib().vgo_falling_land(C.byref(self._w), pos.ctypes.data_as(C.POINTER(C.c_float)), _p(types), n, _p(keep),
                                  _p(placed, C.c_uint32), C.byref(npl), _p(nxt, C.c_uint32), C.byref(nn)) != 0:
            raise ValueError("falling voxel outside the loaded grid")
        return keep[:n].copy(), placed[: npl.value].copy(), nxt[: nn.value].copy()

    def light_flood(self) -> np.ndarray:
        light = np.zeros_like(self.voxels)
        lib().vgo_light_flood_world(C.byref(self._w), _p(light))
        return light


def height_map(area_xy, max_height, cam_ax, cam_ay, pixels=None):
    area_xy = np.ascontiguousarray(area_xy, np.uint32).reshape(-1, 2)
    max_height = np.ascontiguousarray(max_height, np.uint8)
    if pixels is None:
        pixels = np.full((512, 512, 4), 255, np.uint8)  # Image::gen_image_color(WHITE)
    lib().vgo_height_map(_p(area_xy, C.c_uint32), area_xy.shape[0], _p(max_height), cam_ax, cam_ay,
                         _p(pixels))
    return pixels


def should_generate_face(cur, nbr) -> bool:
    return bool(lib().vgo_should_generate_face(cur, nbr))


def should_generate_top_face(cur, nbr) -> bool:
    return bool(lib().vgo_should_generate_top_face(cur, nbr))


# ------------------------------------------------------------------------------ per-frame visibility
def make_view(cam_pos, cam_target, render_size=8, head_in_water=False, show_map=False) -> View:
    """Camera3D position / target -> View, with look = (target - position).normalize_or_zero()
    in f32 like glam (renderer.rs:381)."""
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


def area_render_distance(look, render_size) -> float:
    a = (C.c_float * 3)(*[float(x) for x in look])
    return float(lib().vgo_area_render_distance(a, int(render_size)))


def is_area_visible(ax, ay, view: View, render_distance) -> bool:
    return bool(lib().vgo_is_area_visible(int(ax), int(ay), C.byref(view), float(render_distance)))


def is_voxel_visible(gx, gy, gz, look, cam_pos, max_distance) -> bool:
    lk = (C.c_float * 3)(*[float(x) for x in look])
    cp = (C.c_float * 3)(*[float(x) for x in cam_pos])
    return bool(lib().vgo_is_voxel_visible(int(gx), int(gy), int(gz), lk, cp, float(max_distance)))


def filter_visible(recs, rec_first, area_xy, view: View):
    """-> dict(area_visible u8[n], visible u32[], type_first u32[28], n_faces, lamps u32[])."""
    recs = np.ascontiguousarray(recs, dtype=REC_DTYPE)
    rec_first = np.ascontiguousarray(rec_first, dtype=np.uint32)
    area_xy = np.ascontiguousarray(area_xy, dtype=np.uint32).reshape(-1, 2)
    n = len(area_xy)
    area_visible = np.zeros(n, np.uint8)
    visible = np.zeros(max(len(recs), 1), np.uint32)
    lamps = np.zeros(max(len(recs), 1), np.uint32)
    type_first = np.zeros(28, np.uint32)
    n_faces, n_lamps = C.c_long(0), C.c_long(0)
    nv = lib().vgo_filter_visible(recs.ctypes.data_as(C.c_void_p), _p(rec_first, C.c_uint32), _p(area_xy, C.c_uint32),
                                  n, C.byref(view), _p(area_visible), _p(visible, C.c_uint32),
                                  _p(type_first, C.c_uint32), C.byref(n_faces), _p(lamps, C.c_uint32),
                                  C.byref(n_lamps))
    return dict(area_visible=area_visible, visible=visible[:nv].copy(), type_first=type_first,
                n_faces=int(n_faces.value), lamps=lamps[:n_lamps.value].copy())


# ------------------------------------------------------------------------------ area save files
AREA_PAYLOAD_BYTES = 3 + 32768
AREA_FILE_MAX_BYTES = 33300


def lz4_decompress_block(src: bytes, cap: int):
    """-> bytes, or None for a malformed stream."""
    a = np.frombuffer(bytes(src), np.uint8) if len(src) else np.zeros(1, np.uint8)
    out = np.empty(max(cap, 1), np.uint8)
    n = lib().vgo_lz4_decompress_block(_p(a), len(src), _p(out), cap)
    return None if n < 0 else out[:n].tobytes()


def lz4_compress_block(src: bytes) -> bytes:
    a = np.frombuffer(bytes(src), np.uint8)
    out = np.empty(len(src) + len(src) // 255 + 64, np.uint8)
    n = lib().vgo_lz4_compress_block(_p(a), len(src), _p(out))
    return out[:n].tobytes()


def area_file_encode(voxels: np.ndarray) -> bytes:
    v = np.ascontiguousarray(voxels, np.uint8).reshape(32768)
    out = np.empty(AREA_FILE_MAX_BYTES, np.uint8)
    n = lib().vgo_area_file_encode(_p(v), _p(out))
    return out[:n].tobytes()


def area_files_encode(voxels: np.ndarray, threads: int = 0):
    """store_all_blocking over a batch on `threads` threads (0 = all). -> (total bytes, sizes[n])."""
    v = np.ascontiguousarray(voxels, np.uint8).reshape(-1, 32768)
    out = np.empty((v.shape[0], AREA_FILE_MAX_BYTES), np.uint8)
    sizes = np.zeros(v.shape[0], np.int64)
    total = lib().vgo_area_files_encode(_p(v), v.shape[0], int(threads), _p(out), _p(sizes, C.c_long))
    return int(total), sizes


def area_file_decode(data: bytes):
    """-> voxels u8[32768], or None where the reference falls back to generation."""
    a = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    v = np.empty(32768, np.uint8)
    return v if lib().vgo_area_file_decode(_p(a), len(data), _p(v)) == 0 else None
