"""`-m "not gpu"`: the N>1 host logic (view / ray-tile sharding + the single gather) under gloo,
world_size 2 and 3, with the CPU oracle standing in for the device render."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from volrend_b200 import dist as vd
from volrend_b200 import synth


def test_sharding_partitions():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 200):
            got = sorted(i for r in range(world) for i in vd.shard_views(n, r, world))
            assert got == list(range(n))
        for (w, h, bh) in ((64, 48, 8), (1920, 1080, 8), (33, 17, 4)):
            rows = np.zeros(h, int)
            for r in range(world):
                for (x0, y0, ww, hh) in vd.shard_bands(w, h, r, world, bh):
                    assert x0 == 0 and ww == w and y0 % 4 == 0
                    rows[y0:y0 + hh] += 1
            assert np.all(rows == 1)
    assert vd.merge_adjacent(vd.shard_bands(64, 48, 0, 1)) == [(0, 0, 64, 48)]
    with pytest.raises(ValueError):
        vd.band_rects(8, 8, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, n_views, q):
    from oracle import binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = synth.make_tree("lego", depth=5, basis_dim=4, seed=3)
    ot = ob.OracleTree.from_synth(st)
    poses = synth.nerf_synthetic_test_poses(n_views)
    opt = ob.make_options()

    def cam(i):
        return ob.make_camera(W, H, synth.focal_for(W), synth.focal_for(W), synth.c2w_to_colmajor12(poses[i]))

    def render_part(band_h, n_parts, part):
        # stand-in for vr_render_bands: this part's bands, compactly, band after band
        rows = [ob.render(ot, cam(1), opt, tile=r, want_float=False, nthreads=1)[1]
                for r in vd.shard_bands(W, H, part, n_parts, band_h)]
        return torch.from_numpy(np.concatenate(rows, 0) if rows else np.zeros((0, W, 4), np.uint8))

    def render_views(idx):
        return torch.from_numpy(np.stack([ob.render(ot, cam(i), opt, want_float=False, nthreads=1)[1] for i in idx])
                                if idx else np.zeros((0, H, W, 4), np.uint8))

    frame = vd.render_tile_sharded(render_part, W, H, rank, world, band_h=8)
    assert vd.band_rows(H, 8, world, rank) == sum(r[3] for r in vd.shard_bands(W, H, rank, world, 8))
    views = vd.render_view_sharded(render_views, n_views, rank, world)
    if rank == 0:
        full = ob.render(ot, cam(1), opt, want_float=False, nthreads=1)[1]
        allv = np.stack([ob.render(ot, cam(i), opt, want_float=False, nthreads=1)[1] for i in range(n_views)])
        q.put((bool(np.array_equal(frame.numpy(), full)), bool(np.array_equal(views.numpy(), allv))))
    else:
        assert frame is None and views is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_and_view_sharding_reassemble_bit_exact(built, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    H = 36 if world == 3 else 48      # 36 rows: ragged bands; 48 rows / world 2: the regular strided path
    procs = [ctx.Process(target=_worker, args=(r, world, port, 40, H, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok_tiles, ok_views = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_tiles and ok_views


def test_band_scatter_plan_reassembles_frames():
    """The 2-D copy plan of the peer-copy gather (no GPU needed: the copies are emulated with numpy strides)."""
    rng = np.random.default_rng(0)
    for (W, H, bh) in ((16, 48, 8), (20, 36, 8), (8, 1080, 8), (12, 17, 4), (4, 8, 16)):
        full = rng.integers(0, 256, (H, W, 4)).astype(np.uint8)
        for world in (1, 2, 3, 8):
            out = np.zeros(H * W * 4, np.uint8)
            for rank in range(world):
                rows = [full[y0:y0 + h] for (_, y0, _, h) in vd.shard_bands(W, H, rank, world, bh)]
                compact = (np.concatenate(rows, 0) if rows else np.zeros((0, W, 4), np.uint8)).reshape(-1)
                assert compact.size == vd.band_rows(H, bh, world, rank) * W * 4
                for (do, dp, so, sp, wb, n) in vd.band_scatter_plan(W, H, bh, world, rank):
                    for r in range(n):
                        out[do + r * dp: do + r * dp + wb] = compact[so + r * sp: so + r * sp + wb]
            assert np.array_equal(out.reshape(H, W, 4), full), (W, H, bh, world)
