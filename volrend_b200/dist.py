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


def band_rows(height: int, band_h: int, world: int, rank: int) -> int:
    """Rows owned by ``rank`` (== vr_band_rows in the C-ABI)."""
    return sum(r[3] for r in shard_bands(1, height, rank, world, band_h))


def render_tile_sharded(render_part: Callable[[int, int, int], "torch.Tensor"], width: int, height: int, rank: int,
                        world: int, band_h: int = 8, dst: int = 0, group=None):
    """``render_part(band_h, world, rank) -> uint8 [rows, W, 4]``: this rank's bands rendered
    compactly by ONE launch (``volrend_b200.render_bands`` / ``vr_render_bands``).  Gathers the
    compact buffers on ``dst`` and interleaves them into the [H,W,4] frame (None elsewhere)."""
    import torch
    import torch.distributed as dist
    part = render_part(band_h, world, rank)
    rows_max = band_rows(height, band_h, world, 0)          # rank 0 owns the most bands
    local = part
    if part.shape[0] < rows_max:                             # equal-sized messages
        local = torch.zeros((rows_max, width, 4), dtype=torch.uint8, device=part.device)
        local[: part.shape[0]] = part
    if world == 1:
        gathered = [local]
    else:
        gathered = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local.contiguous(), gathered, dst=dst, group=group)
        if rank != dst:
            return None
    frame = torch.empty((height, width, 4), dtype=torch.uint8, device=part.device)
    full_bands = height // band_h
    if full_bands % world == 0 and height % band_h == 0:
        # regular case: one strided copy per rank
        fv = frame.view(full_bands // world, world, band_h, width, 4)
        for rk in range(world):
            fv[:, rk] = gathered[rk][: (full_bands // world) * band_h].view(full_bands // world, band_h, width, 4)
        return frame
    for rk in range(world):   # ragged case: one index_copy per rank (row indices cached)
        idx = _band_row_index(height, band_h, world, rk, part.device)
        frame.index_copy_(0, idx, gathered[rk][: idx.numel()])
    return frame


_ROW_INDEX_CACHE: dict = {}


def _band_row_index(height: int, band_h: int, world: int, rank: int, device):
    import torch
    key = (height, band_h, world, rank, str(device))
    if key not in _ROW_INDEX_CACHE:
        rows = [y for r in shard_bands(1, height, rank, world, band_h) for y in range(r[1], r[1] + r[3])]
        _ROW_INDEX_CACHE[key] = torch.tensor(rows, dtype=torch.long, device=device)
    return _ROW_INDEX_CACHE[key]


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


def band_scatter_plan(width: int, height: int, band_h: int, world: int, rank: int, bytes_per_pixel: int = 4):
    """2-D copies that move ``rank``'s compact bands (band after band, ``band_rows`` rows) to their interleaved
    rows of a full [H, W] frame: a list of (dst_offset, dst_pitch, src_offset, src_pitch, width_bytes, rows), all in
    bytes relative to the start of the frame / of the compact buffer.  The full-height bands are ONE strided copy
    (row = one band); a ragged last band, when this rank owns it, is a second, contiguous one.  This is what the
    copy engines execute (vr_copy2d_async / cudaMemcpy3DPeerAsync in vr_mg.cu), no kernel involved."""
    row = width * bytes_per_pixel
    band = row * band_h
    n_bands = (height + band_h - 1) // band_h
    owned = len(range(rank, n_bands, world))
    ragged = height % band_h != 0 and owned > 0 and (n_bands - 1) % world == rank
    full = owned - 1 if ragged else owned
    plan = []
    if full > 0:
        plan.append((rank * band, band * world, 0, band, band, full))
    if ragged:
        tail = (height % band_h) * row
        plan.append((rank * band + full * band * world, tail, full * band, tail, tail, 1))
    return plan
