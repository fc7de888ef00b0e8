"""Host-side region helpers (pure index math; mirrors src/service/active_zone.rs and the
multi-GPU sharding of DESIGN.md).  No voxel data is touched here."""
from __future__ import annotations

import numpy as np

#: area containing the spawn point: world (0,0) = internal (1_000_000, 1_000_000)
#: (world_actions.rs:138, location.rs:6,76-84) -> area 62500
SPAWN_AREA = (62500, 62500)


def region_xy(ax0: int, ay0: int, nx: int, ny: int) -> np.ndarray:
    """Dense nx*ny grid of AreaLocations, index = ix + nx*iy."""
    ix, iy = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64))
    xy = np.stack([ix.ravel() + ax0, iy.ravel() + ay0], axis=1)
    return np.ascontiguousarray(xy, dtype=np.uint32)


def spawn_region(n: int) -> tuple[int, int, int, int]:
    """N x N region centred on the spawn area: [62500-N/2, 62500+N/2)^2."""
    return SPAWN_AREA[0] - n // 2, SPAWN_AREA[1] - n // 2, n, n


def render_zone(center: tuple[int, int], render_size: int) -> np.ndarray:
    """get_render_zone (active_zone.rs:10-32): square of side 2r+1, x outer, y inner."""
    r = int(render_size)
    # `(area_location.x as i32 + x) as u32`: wraps below zero like the reference
    out = [((center[0] + x) & 0xFFFFFFFF, (center[1] + y) & 0xFFFFFFFF)
           for x in range(-r, r + 1) for y in range(-r, r + 1)]
    return np.asarray(out, dtype=np.uint32)


def render_zone_on_world_load(center, render_size: int) -> np.ndarray:
    """get_render_zone_on_world_load (active_zone.rs:35-43): max(r - 3, 7)."""
    return render_zone(center, max(max(render_size - 3, 0), 7))


def load_zone(center, render_size: int) -> np.ndarray:
    """get_load_zone (active_zone.rs:46-48): r + 2."""
    return render_zone(center, render_size + 2)


def load_zone_on_world_load(center, render_size: int) -> np.ndarray:
    """get_load_zone_on_world_load (active_zone.rs:51-56): r + 1."""
    return render_zone(center, render_size + 1)


def shard_rows(ny: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous strip of rows [lo, hi) of an ny-row region owned by `rank` (equal counts, the
    remainder spread over the first ranks). No halo: vx_mesh regenerates the ring it needs."""
    base, rem = divmod(ny, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


#: Cost model of the config-3 step on B200 (kernel times of bench.py at N = 1), in units of "K1 of one
#: area" (0.491 ms / 16 900 = 29 ns): generating an area costs 1 + CAVE_VOXEL_WEIGHT per voxel that runs
#: the cave noise (gen_caves: 0.304 ms / 6.77 M = 0.045 ns each) + CAVE_AREA_WEIGHT if it has any
#: (gen_finish: 0.038 ms over ~2 700 areas), meshing it MESH_WEIGHT (count + emit: 28 ns).
CAVE_VOXEL_WEIGHT = 0.00155
CAVE_AREA_WEIGHT = 0.5
MESH_WEIGHT = 0.97


def strip_cost(gen_cost, mesh_cost, lo: int, hi: int) -> float:
    """Cost of the strip of rows [lo, hi): it generates rows lo - 1 .. hi (the ring is regenerated) and
    meshes rows lo .. hi - 1.  gen_cost has ny + 2 entries (entry i = row i - 1), mesh_cost ny."""
    return float(np.sum(gen_cost[lo:hi + 2]) + np.sum(mesh_cost[lo:hi]))


def balanced_rows(gen_cost, mesh_cost, world_size: int) -> list[tuple[int, int]]:
    """Cut rows 0..ny into world_size contiguous non-empty strips [lo, hi) whose largest cost
    (strip_cost: own rows generated and meshed, plus the two ring rows generated) is minimal:
    binary search on the bound with a greedy sweep.  Deterministic: every rank computes the same cut."""
    gen_cost = np.asarray(gen_cost, np.float64)
    mesh_cost = np.asarray(mesh_cost, np.float64)
    ny = len(mesh_cost)
    assert len(gen_cost) == ny + 2 and ny >= world_size >= 1

    def sweep(bound):
        strips, lo = [], 0
        while lo < ny:
            hi = lo + 1
            while hi < ny and strip_cost(gen_cost, mesh_cost, lo, hi + 1) <= bound:
                hi += 1
            strips.append((lo, hi))
            lo = hi
        return strips

    lo_b = max(strip_cost(gen_cost, mesh_cost, r, r + 1) for r in range(ny))  # a single row must fit
    hi_b = strip_cost(gen_cost, mesh_cost, 0, ny)
    for _ in range(60):
        mid = 0.5 * (lo_b + hi_b)
        if len(sweep(mid)) <= world_size:
            hi_b = mid
        else:
            lo_b = mid
    strips = sweep(hi_b)
    while len(strips) < world_size:  # fewer strips than ranks: halve the one with the most rows
        k = max(range(len(strips)), key=lambda i: strips[i][1] - strips[i][0])
        a, b = strips[k]
        assert b - a >= 2
        strips[k:k + 1] = [(a, (a + b) // 2), ((a + b) // 2, b)]
    return strips


def row_costs(cave_columns, cave_voxels, nx: int, ny: int):
    """(gen_cost[ny + 2], mesh_cost[ny]) of an nx x ny region from the cost hints (vx_cave_columns) of its
    (nx + 2) x (ny + 2) ring-padded area grid (index = ix + (nx + 2) * iy, as region_xy orders it)."""
    cc = np.asarray(cave_columns, np.float64).reshape(ny + 2, nx + 2)
    cv = np.asarray(cave_voxels, np.float64).reshape(ny + 2, nx + 2)
    gen = (nx + 2) + CAVE_VOXEL_WEIGHT * cv.sum(axis=1) + CAVE_AREA_WEIGHT * (cc > 0).sum(axis=1)
    mesh = np.full(ny, MESH_WEIGHT * nx)
    return gen, mesh
