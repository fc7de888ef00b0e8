"""Round-2 overlap exploration: the device-resident step from 1..3 host threads with one context each
(kernels of different contexts co-run where the SM has room).  Run once per library variant:
  VX_LIB=voxel_game_b200/build/var_x.so python profiles/scripts/r02_overlap.py [steps]"""
import sys, time, os, threading, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if os.environ.get("VX_LIB"):
    shutil.copy(os.environ["VX_LIB"], os.path.join(ROOT, "voxel_game_b200", "libvoxel_cuda.so"))
import numpy as np
import torch
import voxel_game_b200 as vg

REGION = 128
seed = vg.hash_world_name("test")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ax0, ay0, nx, ny = vg.spawn_region(REGION)
g, m = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2), vg.region_xy(ax0, ay0, nx, ny)

def run(nthreads):
    ctxs = [vg.Context(0, seed=seed) for _ in range(nthreads)]
    def one(c):
        c.generate(g, read=False)
        c.mesh(m)
    for c in ctxs:
        for _ in range(3): one(c)
    torch.cuda.synchronize()
    barrier = threading.Barrier(nthreads + 1)
    def worker(k):
        barrier.wait()
        for _ in range(steps): one(ctxs[k])
        ctxs[k].sync()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(nthreads)]
    for t in th: t.start()
    barrier.wait()
    t0 = time.perf_counter()
    for t in th: t.join()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / (steps * nthreads)
    print(f"{os.environ.get('VX_LIB','default')}: {nthreads} contexts: {ms:.3f} ms/step  {len(m)/ms/1e3:.2f} M chunks/s", flush=True)
    for c in ctxs: c.close()

for n in (1, 2, 3):
    run(n)
