"""Same-box A/B of the save-file encoder: kernel time of vx_encode_areas (size + emit passes, CUDA events
inside the library) over the 130x130 region, and a checksum of the files (all variants must agree).
  VX_LIB_PATH=voxel_game_b200/build/var_x.so python profiles/scripts/r02_ab_codec.py"""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voxel_game_b200 as vg

seed = vg.hash_world_name("test")
xy = vg.region_xy(62500 - 65, 62500 - 65, 130, 130)
with vg.Context(0, seed=seed) as c:
    c.set_timing(True)
    c.generate(xy, read=False)
    for _ in range(3):
        c.encode_areas(xy, read=False)
    c.timings()
    k = 20
    for _ in range(k):
        total = c.encode_areas(xy, read=False)
    ms = c.timings()["codec_ms"] / k
    blob, offs = c.encode_areas(xy)
    print(os.path.basename(os.environ.get("VX_LIB_PATH", "default")), f"encode kernels {ms:.4f} ms  files {total} B  crc {zlib.crc32(blob.tobytes()):08x}")
