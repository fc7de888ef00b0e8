import time, numpy as np, voxel_game_b200 as vg, oracle
seed = oracle.hash_world_name("test")
ctx = vg.Context(0, seed=seed)
print("fp64 rate Gop/s, ms:", ctx.fp64_rate())
ctx.set_timing(True)
for n in (32, 128):
    ax0, ay0, nx, ny = vg.spawn_region(n)
    xy = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2)
    for rep in range(3):
        t = time.time(); ctx.generate(xy, read=False); dt = time.time() - t
        print(n, "gen", len(xy), "areas", f"{dt*1e3:.2f} ms wall", ctx.timings())
    inner = vg.region_xy(ax0, ay0, nx, ny)
    for rep in range(3):
        t = time.time(); info = ctx.mesh(inner); dt = time.time() - t
        print(n, "mesh", len(inner), f"{dt*1e3:.2f} ms wall", info, ctx.timings())
