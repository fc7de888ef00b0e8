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
