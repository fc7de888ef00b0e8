"""Where a batch of the edit storm (config 4) spends its time: wall clock per call, 1 000 edits per batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import voxel_game_b200 as vg

seed = vg.hash_world_name("test")
n_side = 64
ex0, ey0, _, _ = vg.spawn_region(n_side)
exy = vg.region_xy(ex0, ey0, n_side, n_side)
with vg.Context(0, seed=seed) as ec:
    ec.set_mesh_retention(True)
    ec.generate(vg.region_xy(ex0 - 1, ey0 - 1, n_side + 2, n_side + 2), read=False)
    ec.mesh(exy)
    rng = np.random.default_rng(1)
    n_edits, batch = 60_000, 1000
    edits = np.zeros(n_edits, vg.EDIT_DTYPE)
    edits["gx"] = ex0 * 16 + rng.integers(0, n_side * 16, n_edits)
    edits["gy"] = ey0 * 16 + rng.integers(0, n_side * 16, n_edits)
    edits["gz"] = rng.integers(1, 127, n_edits)
    edits["voxel"] = rng.choice([0, 0, 1, 6, 9, 11, 26], n_edits)
    locs = np.stack([edits["gx"], edits["gy"], edits["gz"]], axis=1).astype(np.uint32)
    t = {"apply_edits": 0.0, "update_locations(no read)": 0.0, "update_locations(read)": 0.0}
    for b0 in range(0, n_edits, batch):
        ec.sync(); t0 = time.perf_counter()
        ec.apply_edits(edits[b0:b0 + batch])
        t1 = time.perf_counter()
        if (b0 // batch) % 2:
            ec.update_locations(locs[b0:b0 + batch], read=False); ec.sync()
            t["update_locations(no read)"] += time.perf_counter() - t1
        else:
            ec.update_locations(locs[b0:b0 + batch])
            t["update_locations(read)"] += time.perf_counter() - t1
        t["apply_edits"] += t1 - t0
    nb = n_edits // batch
    print(f"per batch of {batch}: apply_edits {t['apply_edits']/nb*1e3:.3f} ms, update_locations without read "
          f"{t['update_locations(no read)']/(nb/2)*1e3:.3f} ms, with read {t['update_locations(read)']/(nb/2)*1e3:.3f} ms")
    ec.set_timing(True)
    ec.timings()
    ec.apply_edits(edits[:batch]); print("edit kernels (apply):", ec.timings()["edit_ms"], "ms")
    ec.update_locations(locs[:batch], read=False); print("edit kernels (update):", ec.timings()["edit_ms"], "ms")
