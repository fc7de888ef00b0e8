"""Generates tests/golden/*.npz from the CPU oracle (oracle/).

The Rust reference cannot be built or run in this image (no cargo; its noise crate is not
vendored), so these vectors are NOT outputs of the reference binary: they freeze the oracle's
outputs for fixed seeds/regions so that (a) the oracle cannot drift silently and (b) the GPU tests
can check the CUDA path on the GPU box without needing anything but this file.  When a Rust build
becomes available, regenerate the same arrays from it and diff (INTEGRATION.md "Pinning parity").

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint64)[0]


def region_case(world_name, ax0, ay0, n):
    seed = oracle.hash_world_name(world_name)
    ring = oracle.region_xy(ax0 - 1, ay0 - 1, n + 2, n + 2)
    vox, mh, stats, _ = oracle.generate_areas(seed, ring)
    world = oracle.World(ax0 - 1, ay0 - 1, n + 2, n + 2, vox, mh)
    inner = oracle.region_xy(ax0, ay0, n, n)
    rows = []
    for ax, ay in inner:
        nf = world.mesh_area(int(ax), int(ay))
        recs, verts, idx = world.emit_area(int(ax), int(ay))
        s = world.slot(int(ax), int(ay))
        rows.append((ax, ay, digest(vox[s]), digest(mh[s]), nf, len(recs), digest(recs), digest(verts), digest(idx)))
    rows = np.array(rows, dtype=np.uint64)
    return dict(seed=np.uint64(seed), world_name=world_name, origin=np.array([ax0, ay0], np.uint32), n=n, rows=rows,
                stats=np.array([stats[k] for k in ("simplex2_evals", "simplex3_evals", "cave_voxels", "cave_columns",
                                                   "trees_placed")], np.uint64))


def main():
    # 8x8 at the spawn area (seed "test"), plus a region with caves and lakes, plus another seed
    np.savez_compressed(os.path.join(HERE, "spawn_8x8_test.npz"), **region_case("test", 62496, 62496, 8))
    np.savez_compressed(os.path.join(HERE, "caves_6x6_test.npz"), **region_case("test", 62470, 62500, 6))
    np.savez_compressed(os.path.join(HERE, "far_4x4_seed2.npz"), **region_case("another world", 1234, 98765, 4))
    # a lake (WaterSource next to Ice/Snow/Sand shores): transparent-voxel face rules in generated terrain
    np.savez_compressed(os.path.join(HERE, "lake_4x4_test.npz"), **region_case("test", 62508, 62504, 4))
    np.savez_compressed(os.path.join(HERE, "cold_4x4_test.npz"), **region_case("test", 62486, 62492, 4))
    # one full area, raw, so that a single chunk can be inspected / diffed by hand
    seed = oracle.hash_world_name("test")
    vox, mh, _ = oracle.Generator(seed).generate_area(62500, 62500)
    np.savez_compressed(os.path.join(HERE, "area_62500_62500_test.npz"), seed=np.uint64(seed), voxels=vox, max_height=mh)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
