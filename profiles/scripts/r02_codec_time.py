"""Kernel times of the save-file codec (CUDA events inside the library) + file statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import voxel_game_b200 as vg

seed = vg.hash_world_name("test")
xy = vg.region_xy(62500 - 65, 62500 - 65, 130, 130)
with vg.Context(0, seed=seed) as c:
    c.set_timing(True)
    c.generate(xy, read=False)
    blob, offs = c.encode_areas(xy)
    c.timings()
    blob, offs = c.encode_areas(xy)
    enc = c.timings()["codec_ms"]
    vox, mh = c.read_areas(xy)
    with vg.Context(0, seed=seed) as d:
        d.set_timing(True)
        d.decode_areas(xy, blob, offs, read_heights=False)
        d.timings()
        ts = []
        for _ in range(5):
            ok, dmh = d.decode_areas(xy, blob, offs)
            ts.append(d.timings()["codec_ms"])
        dvox, _ = d.read_areas(xy)
    # sequences per file (LZ4 token walk on the host, first 200 files)
    seqs = []
    for i in range(200):
        f = bytes(blob[int(offs[i]):int(offs[i + 1])]); ip, n, k = 4, len(f), 0
        while ip < n:
            t = f[ip]; ip += 1; lit = t >> 4
            if lit == 15:
                while True:
                    b = f[ip]; ip += 1; lit += b
                    if b != 255: break
            ip += lit; k += 1
            if ip >= n: break
            ip += 2
            if (t & 15) == 15:
                while True:
                    b = f[ip]; ip += 1
                    if b != 255: break
        seqs.append(k)
    print(f"encode {enc:.4f} ms  decode {min(ts):.4f} ms (runs {['%.3f' % t for t in ts]})  bytes/area {len(blob)/len(xy):.0f}  "
          f"sequences/file mean {np.mean(seqs):.0f} max {max(seqs)}  round trip ok {bool(ok.all() and np.array_equal(dvox, vox) and np.array_equal(dmh, mh))}")
