"""Round-2 exploration 3: two host threads, one context each (GIL released inside the C calls)."""
import sys, time, os, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import voxel_game_b200 as vg

REGION = 128
seed = vg.hash_world_name("test")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ax0, ay0, nx, ny = vg.spawn_region(REGION)
g, m = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2), vg.region_xy(ax0, ay0, nx, ny)

def run(mode):
    ctxs = [vg.Context(0, seed=seed) for _ in range(nthreads)]
    bufs = []
    for c in ctxs:
        c.generate(g, read=False); inf = c.mesh(m)
        total = c.encode_areas(g, read=False)
        bufs.append(dict(files=vg.PinnedArray((int(total * 1.1) + 4096,), np.uint8), rf=vg.PinnedArray((len(m) + 1,), np.uint32),
                         packed=vg.PinnedArray((int(inf["n_recs"] * 1.05),), np.uint32)))
    def one(c, b):
        c.generate(g, read=False)
        if mode == "e2e":
            c.encode_areas(g, out=b["files"].array)
        inf = c.mesh(m)
        if mode == "e2e":
            c.mesh_read_compact(inf, out=(b["rf"].array, b["packed"].array))
    for c, b in zip(ctxs, bufs):
        for _ in range(3): one(c, b)
    torch.cuda.synchronize()
    barrier = threading.Barrier(nthreads + 1)
    def worker(k):
        barrier.wait()
        for _ in range(steps): one(ctxs[k], bufs[k])
        ctxs[k].sync()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(nthreads)]
    for t in th: t.start()
    barrier.wait()
    t0 = time.perf_counter()
    for t in th: t.join()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / (steps * nthreads)
    print(f"{mode} with {nthreads} host threads/contexts: {ms:.3f} ms/step  {len(m)/ms/1e3:.2f} M chunks/s")
    for c in ctxs: c.close()

run("device")
run("e2e")
