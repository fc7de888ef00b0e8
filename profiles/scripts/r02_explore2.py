"""Round-2 exploration 2: where the host time of a cold-list step goes; two-context compact e2e."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import voxel_game_b200 as vg

REGION = 128
seed = vg.hash_world_name("test")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20

def region(shift=0):
    ax0, ay0, nx, ny = vg.spawn_region(REGION)
    ax0 += shift
    return vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2), vg.region_xy(ax0, ay0, nx, ny)

ctx = vg.Context(0, seed=seed)
lists = [region(130 * (i + 1)) for i in range(steps + 3)]
prev = None
acc = {"unload": 0.0, "generate": 0.0, "mesh": 0.0}
for i in range(steps + 3):
    gi, mi = lists[i]
    t0 = time.perf_counter()
    if prev is not None:
        ctx.unload(prev)
    t1 = time.perf_counter()
    ctx.generate(gi, read=False)
    t2 = time.perf_counter()
    ctx.mesh(mi)
    t3 = time.perf_counter()
    prev = gi
    if i >= 3:
        acc["unload"] += t1 - t0; acc["generate"] += t2 - t1; acc["mesh"] += t3 - t2
torch.cuda.synchronize()
print("cold-list host wall per call (ms):", {k: round(v * 1e3 / steps, 3) for k, v in acc.items()})
ctx.unload(prev)

# same-list for comparison
g, m = region()
acc = {"generate": 0.0, "mesh": 0.0}
for i in range(steps + 3):
    t1 = time.perf_counter()
    ctx.generate(g, read=False)
    t2 = time.perf_counter()
    info = ctx.mesh(m)
    t3 = time.perf_counter()
    if i >= 3:
        acc["generate"] += t2 - t1; acc["mesh"] += t3 - t2
torch.cuda.synchronize()
print("same-list host wall per call (ms):", {k: round(v * 1e3 / steps, 3) for k, v in acc.items()})

# ---- E: compact e2e, two contexts alternating
c2 = vg.Context(0, seed=seed)
ctxs = [ctx, c2]
bufs = []
for c in ctxs:
    c.generate(g, read=False); inf = c.mesh(m)
    total = c.encode_areas(g, read=False)
    bufs.append(dict(files=vg.PinnedArray((int(total * 1.1) + 4096,), np.uint8), rf=vg.PinnedArray((len(m) + 1,), np.uint32),
                     packed=vg.PinnedArray((int(inf["n_recs"] * 1.05),), np.uint32), mh=vg.PinnedArray((len(g), 256), np.uint8)))
def blocking_part(c, b):
    c.encode_areas(g, out=b["files"].array)
    inf = c.mesh(m)
    c.mesh_read_compact(inf, out=(b["rf"].array, b["packed"].array))
def step(i):
    cur, nxt = i & 1, (i + 1) & 1
    ctxs[nxt].generate(g, read=False)
    blocking_part(ctxs[cur], bufs[cur])
ctxs[0].generate(g, read=False)
for i in range(4):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    step(i)
torch.cuda.synchronize()
e = (time.perf_counter() - t0) * 1e3 / steps
print(f"E compact e2e, two contexts: {e:.3f} ms/step  {len(m)/e/1e3:.2f} M chunks/s")
# single context reference
def single(i):
    ctx.generate(g, read=False)
    blocking_part(ctx, bufs[0])
for i in range(3): single(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps): single(i)
torch.cuda.synchronize()
d = (time.perf_counter() - t0) * 1e3 / steps
print(f"D compact e2e, one context: {d:.3f} ms/step  {len(m)/d/1e3:.2f} M chunks/s")
