"""z-slab sharding of the dense SDF grid across ranks (host-side logic; transport is torch.distributed).

Every query point is independent given the image, so the (z,y,x) grid -- x fastest, z slowest, i.e. slabs
are contiguous in the output array -- is cut into contiguous z-ranges, one per rank.  No data-path
collective is needed until the slabs are gathered to rank 0 (``dist.gather`` of slabs padded to the largest
one, trimmed on unpack); for a host result every rank writes its slab straight into one shared pinned host grid.
"""
from __future__ import annotations


def z_bounds(R: int, world: int):
    """plane boundaries: rank r owns z-planes [b[r], b[r+1])  (floor(r*R/world))."""
    if world < 1 or R < 1:
        raise ValueError("world and R must be positive")
    return [(r * R) // world for r in range(world + 1)]


def slab(R: int, world: int, rank: int):
    b = z_bounds(R, world)
    return b[rank], b[rank + 1]


def max_planes(R: int, world: int) -> int:
    b = z_bounds(R, world)
    return max(b[i + 1] - b[i] for i in range(world))


def unpack_gathered(full, R: int, world: int, out):
    """full: [world*max_planes, R, R] (all_gather of padded slabs) -> out[R,R,R] (tensor or ndarray)."""
    b = z_bounds(R, world)
    mp = max_planes(R, world)
    for r in range(world):
        n = b[r + 1] - b[r]
        out[b[r]:b[r + 1]] = full[r * mp:r * mp + n]
    return out


def unpack_gather_list(gathered, R: int, world: int, out):
    """gathered: list of `world` padded slabs [max_planes, R, R] (dist.gather on rank 0) -> out[R,R,R]."""
    b = z_bounds(R, world)
    for r in range(world):
        out[b[r]:b[r + 1]] = gathered[r][:b[r + 1] - b[r]]
    return out
