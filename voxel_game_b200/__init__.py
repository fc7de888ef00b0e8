"""voxel_game_b200 -- B200-native (sm_100a) chunk pipeline of IvanDimovSIT/voxel_game:
terrain generation -> column heights -> face-culled meshing, plus batched edits.

The product is libvoxel_cuda.so (hand-written CUDA behind the C ABI of include/voxel_cuda.h);
this package is the thin Python plumbing used by tests and bench.py.  No CPU fallback.
"""
from .binding import (  # noqa: F401
    ABI_SYMBOLS, EDIT_DTYPE, LIB_PATH, REC_DTYPE, VERTEX_DTYPE, Context, PinnedArray, View, VxError,
    build, hash_world_name, load, make_view, unpack_records,
)
from .host import HOST_LIB_PATH, HOST_SYMBOLS, HostRenderer, expand_records, load_host  # noqa: F401
from .world import balanced_rows, region_xy, render_zone, row_costs, shard_rows, spawn_region, strip_cost  # noqa: F401

__all__ = ["Context", "PinnedArray", "VxError", "build", "load", "hash_world_name", "region_xy",
           "shard_rows", "spawn_region", "ABI_SYMBOLS", "REC_DTYPE", "VERTEX_DTYPE", "EDIT_DTYPE",
           "LIB_PATH", "View", "make_view", "unpack_records", "render_zone"]
