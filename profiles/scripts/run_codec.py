"""Small driver for profiling the save-file codec kernels under ncu (not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import voxel_game_b200 as vg

seed = vg.hash_world_name("test")
xy = vg.region_xy(62500 - 64, 62500 - 64, 128, 128)
with vg.Context(0, seed=seed) as c:
    c.generate(xy, read=False)
    for _ in range(3):
        blob, offs = c.encode_areas(xy)
    with vg.Context(0, seed=seed) as d:
        for _ in range(2):
            ok, _ = d.decode_areas(xy, blob, offs, read_heights=False)
    print("bytes per area", len(blob) / len(xy), "all ok", bool(ok.all()))
