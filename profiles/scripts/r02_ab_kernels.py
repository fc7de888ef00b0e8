"""Same-box A/B of library variants: kernel times of the config-3 step (CUDA events inside the library)
and a checksum of the generated block IDs (all variants must agree).
  VX_LIB_PATH=voxel_game_b200/build/var_x.so python profiles/scripts/r02_ab_kernels.py [steps]"""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import voxel_game_b200 as vg

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = vg.hash_world_name("test")
ax0, ay0, nx, ny = vg.spawn_region(128)
ny = int(os.environ.get("VX_AB_ROWS", ny))  # e.g. 16: the strip of one rank of eight
g, m = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2), vg.region_xy(ax0, ay0, nx, ny)
with vg.Context(0, seed=seed) as ctx:
    ctx.set_timing(True)
    for _ in range(5):
        ctx.generate(g, read=False); info = ctx.mesh(m)
    ctx.timings()
    import time
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(steps):
        ctx.generate(g, read=False); info = ctx.mesh(m)
    ctx.sync(); wall = (time.perf_counter() - t0) * 1e3 / steps
    t = ctx.timings()
    vox, mh = ctx.read_areas(g[::7])
    crc = zlib.crc32(vox.tobytes()) ^ zlib.crc32(mh.tobytes())
    k = {n: t[n + "_ms"] / steps for n in ("gen_columns", "gen_caves", "gen_finish", "mesh_count", "mesh_emit")}
    print(os.path.basename(os.environ.get("VX_LIB_PATH", "default")), f"rows {ny} step {wall:.4f} ms", " ".join(f"{a}={b:.4f}" for a, b in k.items()),
          f"cave_evals={t['cave_evals']} faces={info['n_faces']} crc={crc:08x}", flush=True)
