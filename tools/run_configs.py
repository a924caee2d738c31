#!/usr/bin/env python
"""Runs the five BASELINE.json configs on the seeded stand-in scenes and records Mrays/s, FPS,
algorithmic GB/s (fraction of the measured HBM peak) and an oracle spot check for each.
1 GPU:   python tools/run_configs.py [1,2,3,4,5]
N GPUs:  python -m torch.distributed.run --nproc-per-node N tools/run_configs.py 4,5
         (config 4: ray-tile sharding of every frame + NCCL gather; config 5: one scene per rank)
Writes gpurun_out/configs_N<world>.json."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402  (checker only)
from volrend_b200 import Camera, N3Tree, RenderOptions, dist as vd, lib, render_bands, render_batch, launch_renderer, synth  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
which = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,3,4,5").split(",")]
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    PEAK = 6650.0
out = {"world": world, "hbm_peak_gbs": PEAK}


def cams_for(poses, W, H, fx):
    cs = []
    for p in poses:
        c = Camera(W, H, fx, fx)
        c.set_c2w(p)
        cs.append(c)
    return cs


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_views(st, tree, cams, opt, bd, label, spot=None):
    W, H = cams[0].width, cams[0].height
    imgs = torch.zeros((len(cams), H, W, 4), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(5, dtype=torch.int64, device=dev)
    render_batch(tree, cams, opt, imgs, counters=cnt)
    torch.cuda.synchronize()
    S, D, SH, HIT, F = cnt.cpu().tolist()
    per_shaded = 6 * bd if bd > 0 else 6
    A = 4 * D + 2 * S + per_shaded * SH + 4 * W * H * len(cams)
    ms = timed(lambda: render_batch(tree, cams, opt, imgs))
    r = dict(label=label, views=len(cams), W=W, H=H, ms_per_frame=ms / len(cams), fps=len(cams) / ms * 1e3,
             mrays_s=W * H * len(cams) / ms / 1e3, alg_gbs=A / ms / 1e6, frac_of_hbm_peak=A / ms / 1e6 / PEAK,
             samples_per_ray=S / (W * H * len(cams)), shaded_per_ray=SH / (W * H * len(cams)),
             ref_child_loads_per_sample=D / max(S, 1), node_fetches_per_sample=F / max(S, 1))
    if spot is not None:   # oracle check of a window of view 0
        x0, y0, w, h = spot
        fo = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        launch_renderer(tree, cams[0], opt, None, None, None, True, float_out=fo, tile=spot)
        torch.cuda.synchronize()
        ot = ob.OracleTree.from_synth(st)
        oc = ob.make_camera(W, H, cams[0].fx, cams[0].fy, np.ascontiguousarray(cams[0].transform, np.float32).reshape(12))
        oo = ob.make_options(step_size=opt.step_size, stop_thresh=opt.stop_thresh, sigma_thresh=opt.sigma_thresh)
        f, _, _ = ob.render(ot, oc, oo, tile=spot, want_u8=False)
        r["oracle_window_max_abs_err"] = float(np.abs(fo.cpu().numpy() - f).max())
    return r


poses200 = synth.nerf_synthetic_test_poses(200)
t00 = time.time()

if 1 in which and rank == 0:
    st = synth.make_config1_tree()
    tree = N3Tree.from_synth(st)
    cams = cams_for([synth.config1_pose()], 64, 64, synth.focal_for(64))
    out["config1"] = bench_views(st, tree, cams, RenderOptions(), 1, "64x64 full depth-4 SH1 (CPU leg = oracle)", spot=(0, 0, 64, 64))
    print("config1", json.dumps(out["config1"]), flush=True)

if 2 in which and rank == 0:
    st = synth.make_tree("lego", depth=10, basis_dim=16, seed=0)
    tree = N3Tree.from_synth(st)
    cams = cams_for(poses200, 800, 800, synth.focal_for(800))
    out["config2"] = bench_views(st, tree, cams, RenderOptions(), 16, "lego stand-in SH16 depth-10, 800x800, 200 poses",
                                 spot=(360, 380, 96, 64))
    print("config2", json.dumps(out["config2"]), flush=True)
    del tree

if 3 in which and rank == 0:
    st = synth.make_tree("drums", depth=10, basis_dim=16, seed=1)
    tree = N3Tree.from_synth(st)
    cams = cams_for(poses200[::5], 800, 800, synth.focal_for(800))
    sweep = []
    for step in (1e-5, 1e-4, 1e-3, 1e-2):
        for stop in (0.0, 1e-3, 1e-2, 1e-1):
            opt = RenderOptions(step_size=step, stop_thresh=stop)
            r = bench_views(st, tree, cams, opt, 16, f"step={step:g} stop={stop:g}",
                            spot=(380, 400, 48, 32) if (step, stop) in ((1e-4, 1e-2), (1e-2, 0.0), (1e-5, 1e-1)) else None)
            sweep.append(r)
            print("config3", json.dumps(r), flush=True)
    for sig in (0.0, 1.0):
        r = bench_views(st, tree, cams, RenderOptions(sigma_thresh=sig), 16, f"sigma_thresh={sig:g}")
        sweep.append(r)
    out["config3"] = dict(scene="drums stand-in SH16 depth-10, 800x800, 40 poses", nodes=st.capacity, sweep=sweep)
    del tree

if 4 in which:
    st = synth.make_tree("gyroid_small", depth=11, basis_dim=25, seed=0, band_cells=1.0)
    tree = N3Tree.from_synth(st)
    W, H, fx = 1920, 1080, 1500.0
    poses = synth.nerf_synthetic_test_poses(40, radius=1.6, elev_deg=25.0)
    cams = cams_for(poses, W, H, fx)
    opt = RenderOptions()
    if world == 1:
        out["config4"] = bench_views(st, tree, cams, opt, 25, "gyroid depth-11 SH25, 1920x1080, 40 poses, 1 GPU",
                                     spot=(900, 500, 64, 48))
        out["config4"]["nodes"] = st.capacity
    else:
        # ray-tile sharding: every frame is split into interleaved 8-row bands; each rank renders its
        # bands with ONE launch into a compact buffer; one NCCL gather per frame to rank 0
        full = None
        rows = vd.band_rows(H, 8, world, rank)
        part_buf = torch.empty((rows, W, 4), dtype=torch.uint8, device=dev)

        def frame(i):
            global full

            def rp(band_h, n_parts, part):
                render_bands(tree, cams[i], opt, band_h, n_parts, part, part_buf)
                return part_buf
            full = vd.render_tile_sharded(rp, W, H, rank, world, 8)

        for i in range(3):
            frame(i)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(len(cams)):
            frame(i)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            one = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
            launch_renderer(tree, cams[-1], opt, one, None, None, True)
            torch.cuda.synchronize()
            out["config4"] = dict(label=f"gyroid depth-11 SH25, 1920x1080, 40 poses, {world} GPUs ray-tile sharded (interleaved 8-row bands, one launch + one gather per frame)",
                                  ms_per_frame=ms / len(cams), mrays_s=W * H * len(cams) / ms / 1e3, nodes=st.capacity,
                                  sharded_equals_single_gpu=bool(torch.equal(full, one)))
    if rank == 0:
        print("config4", json.dumps(out["config4"]), flush=True)
    del tree

if 5 in which:
    scenes = list(range(8))
    mine = scenes[rank::world]
    tot_ms, n_frames = 0.0, 0
    per_scene = []
    for sd in mine:
        st = synth.make_tree("lego" if sd % 2 == 0 else "drums", depth=9, basis_dim=16, seed=sd)
        tree = N3Tree.from_synth(st)
        cams = cams_for(poses200, 800, 800, synth.focal_for(800))
        imgs = torch.zeros((len(cams), 800, 800, 4), dtype=torch.uint8, device=dev)
        ms = timed(lambda: render_batch(tree, cams, RenderOptions(), imgs), reps=2, warm=1)
        per_scene.append(dict(seed=sd, nodes=st.capacity, ms_per_frame=ms / len(cams)))
        tot_ms += ms
        n_frames += len(cams)
        del tree
    t = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
    n = torch.tensor([float(n_frames)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    if rank == 0:
        out["config5"] = dict(label=f"8 scenes x 200 poses, 800x800, {world} GPU(s), scenes round-robin over ranks",
                              total_ms=float(t.item()), mrays_s=800 * 800 * float(n.item()) / float(t.item()) / 1e3,
                              rank0_scenes=per_scene)
        print("config5", json.dumps(out["config5"]), flush=True)

if rank == 0:
    out["wall_s"] = time.time() - t00
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"configs_N{world}.json"), "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
