"""Wall time of vx_generate / vx_mesh against the device time of their kernels (host overhead)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import voxel_game_b200 as vg

seed = vg.hash_world_name("test")
ax0, ay0, nx, ny = vg.spawn_region(128)
gen_xy = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2)
mesh_xy = vg.region_xy(ax0, ay0, nx, ny)
with vg.Context(0, seed=seed) as c:
    c.set_timing(True)
    for _ in range(5):
        c.generate(gen_xy, read=False)
        c.mesh(mesh_xy)
    g_wall = m_wall = g_dev = m_dev = 0.0
    n = 30
    t_all = time.perf_counter()
    for _ in range(n):
        t0 = time.perf_counter()
        c.generate(gen_xy, read=False)
        t1 = time.perf_counter()
        t = c.timings()
        g_dev += t["gen_columns_ms"] + t["gen_caves_ms"] + t["gen_finish_ms"]
        t2 = time.perf_counter()
        c.mesh(mesh_xy)
        t3 = time.perf_counter()
        t = c.timings()
        m_dev += t["mesh_count_ms"] + t["mesh_emit_ms"]
        g_wall += t1 - t0
        m_wall += t3 - t2
    t_all = time.perf_counter() - t_all
    print(f"generate: wall {g_wall / n * 1e3:.3f} ms, kernels {g_dev / n:.3f} ms")
    print(f"mesh:     wall {m_wall / n * 1e3:.3f} ms, kernels {m_dev / n:.3f} ms")
    print(f"loop:     {t_all / n * 1e3:.3f} ms per step")
    c.set_timing(False)
    t_all = time.perf_counter()
    for _ in range(n):
        c.generate(gen_xy, read=False)
        c.mesh(mesh_xy)
    t_all = time.perf_counter() - t_all
    print(f"loop without event timing: {t_all / n * 1e3:.3f} ms per step")
