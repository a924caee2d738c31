"""Multi-GPU sharding of the ray march (one process per GPU, torch.distributed).

Rays are independent and the tree is read-only (reference ``src/cuda/volrend.cu:78-173`` has no
inter-thread communication), so the path shards with no data-path collective except ONE gather
of finished RGBA8 pixels to rank 0:

* view sharding   pose i -> rank i mod world            (BASELINE configs 2, 3, 5)
* tile sharding   interleaved row bands of one view     (BASELINE config 4: 1920x1080)

The gather runs over NCCL (NVLink/NVSwitch) on GPUs and over gloo in the CPU tests; the
render callback is injected so the host logic is testable without a device.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

Rect = Tuple[int, int, int, int]  # x0, y0, w, h


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin view assignment: neighbouring poses cost about the same, so this balances."""
    return list(range(rank, n_views, world))


def band_rects(width: int, height: int, band_h: int = 8) -> List[Rect]:
    """Full-width row bands, band_h a multiple of the 4-row warp tile."""
    if band_h % 4:
        raise ValueError("band_h must be a multiple of 4 (warp tiles are 8x4 pixels)")
    return [(0, y, width, min(band_h, height - y)) for y in range(0, height, band_h)]


def shard_bands(width: int, height: int, rank: int, world: int, band_h: int = 8) -> List[Rect]:
    """Interleaved bands (band b -> rank b mod world): scene-dependent cost averages out."""
    return band_rects(width, height, band_h)[rank::world]


def merge_adjacent(rects: Sequence[Rect]) -> List[Rect]:
    """Coalesce vertically adjacent full-width bands (world == 1 => one full-frame rect)."""
    out: List[Rect] = []
    for r in rects:
        if out and out[-1][0] == r[0] and out[-1][2] == r[2] and out[-1][1] + out[-1][3] == r[1]:
            p = out.pop()
            out.append((p[0], p[1], p[2], p[3] + r[3]))
        else:
            out.append(r)
    return out


def render_tile_sharded(render_rect: Callable[[Rect], "torch.Tensor"], width: int, height: int, rank: int,
                        world: int, band_h: int = 8, dst: int = 0, group=None):
    """Render this rank's bands with ``render_rect(rect) -> uint8 [h,w,4]`` and gather the whole
    frame on ``dst``.  Returns the [H,W,4] frame on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    mine = shard_bands(width, height, rank, world, band_h)
    parts = [render_rect(r) for r in mine]
    n_bands = len(band_rects(width, height, band_h))
    per_rank = (n_bands + world - 1) // world
    # equal-sized messages: pad the band list of the ranks that own one band fewer
    dev = parts[0].device if parts else torch.device("cpu")
    local = torch.zeros((per_rank, band_h, width, 4), dtype=torch.uint8, device=dev)
    for i, (r, p) in enumerate(zip(mine, parts)):
        local[i, : r[3]] = p
    if world == 1:
        gathered = [local]
    else:
        gathered = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local, gathered, dst=dst, group=group)
        if rank != dst:
            return None
    frame = torch.empty((height, width, 4), dtype=torch.uint8, device=dev)
    for rk in range(world):
        for i, r in enumerate(shard_bands(width, height, rk, world, band_h)):
            frame[r[1]: r[1] + r[3]] = gathered[rk][i, : r[3]]
    return frame


def render_view_sharded(render_views: Callable[[List[int]], "torch.Tensor"], n_views: int, rank: int, world: int,
                        dst: int = 0, group=None):
    """``render_views(indices) -> uint8 [len(indices),H,W,4]``; gathers all views, in pose order,
    on ``dst`` ([n_views,H,W,4]); None elsewhere."""
    import torch
    import torch.distributed as dist
    mine = shard_views(n_views, rank, world)
    local = render_views(mine)
    per_rank = (n_views + world - 1) // world
    if local.shape[0] < per_rank:   # pad to equal message size
        pad = torch.zeros((per_rank - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    if world == 1:
        return local[:n_views]
    gathered = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty((n_views,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for rk in range(world):
        idx = shard_views(n_views, rk, world)
        out[idx] = gathered[rk][: len(idx)]
    return out
