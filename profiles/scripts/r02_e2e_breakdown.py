"""Where the end-to-end step (deferred reads) spends its host time: wall clock per call, averaged.
  python profiles/scripts/r02_e2e_breakdown.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import voxel_game_b200 as vg

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = vg.hash_world_name("test")
ax0, ay0, nx, ny = vg.spawn_region(128)
g, m = vg.region_xy(ax0 - 1, ay0 - 1, nx + 2, ny + 2), vg.region_xy(ax0, ay0, nx, ny)
with vg.Context(0, seed=seed) as c:
    c.generate(g, read=False); info = c.mesh(m); total = c.encode_areas(g, read=False)
    sets = [{"files": vg.PinnedArray((int(total * 1.1),), np.uint8), "off": vg.PinnedArray((len(g) + 1,), np.uint64),
             "mh": vg.PinnedArray((len(g), 256), np.uint8), "rf": vg.PinnedArray((len(m) + 1,), np.uint32),
             "packed": vg.PinnedArray((int(info["n_recs"] * 1.05),), np.uint32)} for _ in range(2)]
    for s in sets:
        for a in s.values():
            a.array.view(np.uint8).reshape(-1)[::4096] = 0
    c.set_deferred_reads(True)
    names = ["generate+heights", "encode_areas", "mesh", "wait_reads", "encode_read", "mesh_read_compact"]
    acc = np.zeros(len(names))
    for i in range(steps + 3):
        p = sets[i & 1]
        t = [time.perf_counter()]
        c.generate(g, max_height_out=p["mh"].array, read=False); t.append(time.perf_counter())
        tot = c.encode_areas(g, read=False); t.append(time.perf_counter())
        inf = c.mesh(m); t.append(time.perf_counter())
        c.wait_reads(); t.append(time.perf_counter())
        c.encode_read(len(g), tot, out=p["files"].array, offsets_out=p["off"].array); t.append(time.perf_counter())
        c.mesh_read_compact(inf, out=(p["rf"].array, p["packed"].array)); t.append(time.perf_counter())
        if i >= 3:
            acc += np.diff(t)
    c.wait_reads()
    c.set_deferred_reads(False)
    acc *= 1e3 / steps
    print("deferred step, ms per call:", " ".join(f"{n}={v:.3f}" for n, v in zip(names, acc)), f"sum={acc.sum():.3f}")
    # generate without the heights copy, for the share of that copy and its sync
    t0 = time.perf_counter()
    for i in range(steps):
        c.generate(g, read=False)
    c.sync()
    print(f"generate alone (no outputs): {(time.perf_counter() - t0) * 1e3 / steps:.3f} ms")
