"""Where does the step time go besides the kernels?  Times the bench step with and without the
library's CUDA-event timers, and the generate half alone (which never waits for the host)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import voxel_game_b200 as vg

seed = vg.hash_world_name("test")
ctx = vg.Context(0, seed=seed)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ax0, ay0, nx, ny = vg.spawn_region(128)
gen_xy = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2)
mesh_xy = vg.region_xy(ax0, ay0, nx, ny)


def timed(fn, steps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps


def step():
    ctx.generate(gen_xy, read=False)
    ctx.mesh(mesh_xy)


for timing in (True, False):
    ctx.set_timing(timing)
    print("timing", timing, "step            dev %.4f ms  wall %.4f ms" % timed(step))
    ctx.timings()
    print("timing", timing, "generate only   dev %.4f ms  wall %.4f ms" % timed(lambda: ctx.generate(gen_xy, read=False)))
    ctx.timings()
    print("timing", timing, "mesh only       dev %.4f ms  wall %.4f ms" % timed(lambda: ctx.mesh(mesh_xy)))
    t = ctx.timings()
    if timing:
        print("   mesh kernels per call: count %.4f emit %.4f" % (t["mesh_count_ms"] / 35, t["mesh_emit_ms"] / 35))
