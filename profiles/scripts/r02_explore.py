"""Round-2 exploration (not a bench value): step variants on one GPU.
  A  one context, same list every step (the round-1 timed loop)
  B  one context, region shifted every step + previous region unloaded (cold lists)
  C  two contexts on their own streams, software-pipelined (mesh of one overlaps generate of the other)
  D  compact end to end: generate -> encode -> files to pinned; mesh -> 4-byte records to pinned
"""
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

def timed(fn, k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(k):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / k

# ---- A
ctx = vg.Context(0, seed=seed)
g, m = region()
for _ in range(5):
    ctx.generate(g, read=False); info = ctx.mesh(m)
a = timed(lambda i: (ctx.generate(g, read=False), ctx.mesh(m)), steps)
print(f"A same-list: {a:.3f} ms/step  {len(m)/a/1e3:.2f} M chunks/s")

# ---- B cold lists
lists = [region(130 * (i + 1)) for i in range(steps + 3)]
prev = [None]
def cold(i):
    gi, mi = lists[i]
    if prev[0] is not None:
        ctx.unload(prev[0])
    ctx.generate(gi, read=False); ctx.mesh(mi)
    prev[0] = gi
for i in range(3):
    cold(i)
b = timed(lambda i: cold(i + 3), steps)
print(f"B cold-list: {b:.3f} ms/step  {len(m)/b/1e3:.2f} M chunks/s")
ctx.unload(prev[0])

# ---- C two contexts pipelined
c2 = vg.Context(0, seed=seed)
ctxs = [ctx, c2]
for c in ctxs:
    c.generate(g, read=False); c.mesh(m)
def pipe(i):
    # generate for the NEXT step goes out on the other context before this step's mesh waits
    cur, nxt = ctxs[i & 1], ctxs[(i + 1) & 1]
    nxt.generate(g, read=False)
    cur.mesh(m)
ctxs[0].generate(g, read=False)
c = timed(pipe, steps)
print(f"C two-context pipeline: {c:.3f} ms/step  {len(m)/c/1e3:.2f} M chunks/s")
c2.close()

# ---- D compact e2e
info = ctx.mesh(m) if ctx.generate(g, read=False) is not None else None
total = ctx.encode_areas(g, read=False)
files = vg.PinnedArray((int(total * 1.1) + 4096,), np.uint8)
mh = vg.PinnedArray((len(g), 256), np.uint8)
rf = vg.PinnedArray((len(m) + 1,), np.uint32)
packed = vg.PinnedArray((int(info["n_recs"] * 1.05),), np.uint32)
def compact(i):
    ctx.generate(g, read=False)
    ctx.encode_areas(g, out=files.array)
    inf = ctx.mesh(m)
    ctx.mesh_read_compact(inf, out=(rf.array, packed.array))
for i in range(3):
    compact(i)
d = timed(compact, steps)
print(f"D compact e2e: {d:.3f} ms/step  {len(m)/d/1e3:.2f} M chunks/s  bytes/step {total + info['n_recs']*4}")
t = timed(lambda i: ctx.encode_areas(g, out=files.array), 5)
print(f"   encode+read alone: {t:.3f} ms")
t = timed(lambda i: ctx.mesh_read_compact(info, out=(rf.array, packed.array)), 5)
print(f"   compact read alone: {t:.3f} ms")
# host expansion of the compact stream
import voxel_game_b200.host as H
ff = np.empty(len(m) + 1, np.uint32)
recs = np.empty(info["n_recs"], vg.REC_DTYPE); verts = np.empty(info["n_faces"] * 4, vg.VERTEX_DTYPE); idx = np.empty(info["n_faces"] * 6, np.uint16)
for th in (1, 8, 16, 32, 0):
    t0 = time.perf_counter()
    vg.expand_records(m, rf.array, packed.array[:info["n_recs"]], threads=th, out=(ff, recs, verts, idx))
    print(f"   host expand threads={th}: {(time.perf_counter()-t0)*1e3:.1f} ms")
hr = vg.HostRenderer(ctx)
for tr in ("compact", "raw"):
    hr.clear()
    t0 = time.perf_counter()
    ms = hr.load_all_blocking(m, transport=tr, threads=0)
    print(f"   HostRenderer.load_all_blocking {tr}: {ms:.1f} ms ({hr.voxel_count()} voxel meshes)")
t0 = time.perf_counter(); hr.clear(); print(f"   clear: {(time.perf_counter()-t0)*1e3:.1f} ms")
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
