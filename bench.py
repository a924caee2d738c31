#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 PlenOctree ray-marcher.

Metric (BASELINE.json): Mrays/s at 800x800 (FPS = Mrays/s / 0.64), plus the achieved
algorithmic GB/s against the measured HBM peak.  Workload = BASELINE config 2 ("lego tree.npz,
800x800, 200 test poses"): the real scene is an external download that does not exist on the
box, so the seeded lego-like stand-in of volrend_b200/synth.py is used and named in `config`.

  step      one sweep of the 200-pose NeRF-synthetic test orbit (main_headless.cpp:208-223),
            rendered by ONE batched launch of the fused march kernel (vr_render_batch)
  value     W*H*views / device time, frames stay in HBM (whole job, all ranks)
  e2e       same sweep through the host-buffer entry point (vr_render_frames_host): one launch
            per pose as the reference CLI does, camera in via kernel parameters, every RGBA8
            frame copied to pinned host memory inside the timed region
  N > 1     weak scaling: each rank renders its own 200 views of a 200*N-view orbit (tree
            replicated), then ONE NCCL gather of the finished RGBA8 frames to rank 0
  --impl reference   the UNMODIFIED reference CUDA renderer (oracle/_ref/libvolrend_ref.so,
            built from /root/reference by oracle/Makefile.ref) on the same tree and poses, timed
            exactly like main_headless.cpp:203-228; falls back to the CPU oracle port when that
            library is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 800
N_POSES = 200
WORKLOAD = ("synthetic lego-like SH16 N3Tree stand-in (depth 10, seed 0; real lego tree.npz is an external "
            "download), 800x800, 200 NeRF-synthetic test-orbit poses, default RenderOptions")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "of fallback (6.65 TB/s, B200_PROFILING.md)"


def build_scene(rank: int, world: int):
    from volrend_b200 import synth
    depth = env_int("VR_BENCH_DEPTH", 10)
    st = synth.make_tree("lego", depth=depth, basis_dim=16, seed=0)
    poses = synth.nerf_synthetic_test_poses(N_POSES * world)[rank::world][:N_POSES]
    return st, poses, depth


def cpu_baseline(st, poses, budget_s: float = 12.0):
    """Oracle port on the host cores, bounded sample of the same workload (full 800x800 frames)."""
    from oracle import binding as ob
    from volrend_b200 import synth
    cores = os.cpu_count() or 1
    ot = ob.OracleTree.from_synth(st)
    opt = ob.make_options()
    t0 = time.perf_counter()
    n, counters = 0, []
    while True:
        pose = poses[(n * 37) % len(poses)]
        cam = ob.make_camera(W, H, synth.focal_for(W), synth.focal_for(W), synth.c2w_to_colmajor12(pose))
        _, _, c = ob.render(ot, cam, opt, want_float=False, want_u8=True, nthreads=cores)
        counters.append(((n * 37) % len(poses), c))
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 400:
            break
    return {"value": W * H * n / el / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{n} full 800x800 frames of the workload, oracle/march_oracle.c with {cores} threads, {el:.1f} s"}, counters


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CUDA renderer (or the CPU port when it is absent)."""
    if rank != 0:
        return
    from volrend_b200 import synth
    st, poses, depth = build_scene(0, 1)
    line = {"impl": "reference", "metric": "Mrays/s @ 800x800", "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tree_depth": depth, "nodes": st.capacity,
                       "l2": "inputs larger than L2 (1.1 GB tree, a different pose every frame)"}}
    from oracle import ref_binding as rb
    import torch
    if rb.available() and torch.cuda.is_available():
        path = "/tmp/vr_bench_ref_tree.npz"
        st.save_npz(path)
        rt = rb.RefTree(path)
        c12 = np.stack([synth.c2w_to_colmajor12(p) for p in poses])
        fx = synth.focal_for(W)
        opt = rb.make_options()
        host = torch.empty((N_POSES, H, W, 4), dtype=torch.uint8).pin_memory()
        for _ in range(args.warmup):
            rt.time_frames(W, H, fx, fx, c12, opt)
        cs = ClockSampler(0)
        cs.start()
        ms = [rt.time_frames(W, H, fx, fx, c12, opt) for _ in range(args.steps)]
        clocks = cs.stop()
        for _ in range(min(args.warmup, 2)):
            rt.time_frames(W, H, fx, fx, c12, opt, with_d2h=True, host_out=host)
        ms_e = [rt.time_frames(W, H, fx, fx, c12, opt, with_d2h=True, host_out=host) for _ in range(args.steps)]
        rt.close()
        t, te = sum(ms) / len(ms), sum(ms_e) / len(ms_e)
        val = W * H * N_POSES / t / 1e3
        line.update({"value": val, "ms_per_step": t, "clocks": clocks, "gpu_launches": N_POSES * args.steps,
                     "reference": "volrend::launch_renderer from /root/reference/src/cuda/volrend.cu, built -arch=sm_100 "
                                  "by oracle/Makefile.ref, timed as main_headless.cpp:203-228 on the same B200",
                     "e2e": {"value": W * H * N_POSES / te / 1e3, "unit": "Mrays/s",
                             "h2d_bytes_per_step": 48 * N_POSES, "d2h_bytes_per_step": 4 * W * H * N_POSES},
                     "cpu_baseline": None})
        cb, _ = cpu_baseline(st, poses, budget_s=8.0)
        line["cpu_baseline"] = cb
    else:
        # no reference binary (or no GPU): the CPU oracle port with all host threads
        per_step = []
        for _ in range(max(1, min(args.steps, 3))):
            cb, _ = cpu_baseline(st, poses, budget_s=8.0)
            per_step.append(cb["value"])
        val = float(np.mean(per_step))
        line.update({"value": val, "ms_per_step": W * H * N_POSES / val / 1e3, "gpu_launches": 0,
                     "cpu_baseline": {**cb, "value": val},
                     "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: volrend_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from volrend_b200 import Camera, N3Tree, RenderOptions, lib, render_batch, render_frames_host, synth
    lib().vr_set_variant(args.variant)
    st, poses, depth = build_scene(rank, world)
    tree = N3Tree.from_synth(st)
    info = tree.info()
    cams = []
    for p in poses:
        c = Camera(W, H, synth.focal_for(W), synth.focal_for(W))
        c.set_c2w(p)
        cams.append(c)
    opt = RenderOptions()
    imgs = [torch.zeros((N_POSES, H, W, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
    host = torch.empty((N_POSES, H, W, 4), dtype=torch.uint8).pin_memory()
    gathered = None
    if world > 1 and rank == 0:
        gathered = [torch.empty_like(imgs[0]) for _ in range(world)]

    # ---- algorithmic bytes of this rank's sweep from the instrumented kernel (not timed)
    cnt = torch.zeros(5, dtype=torch.int64, device=dev)
    render_batch(tree, cams, opt, imgs[0], counters=cnt)
    torch.cuda.synchronize()
    S, D, SH, HIT, FETCH = [int(v) for v in cnt.cpu().tolist()]
    a_step = 4 * D + 2 * S + 6 * 16 * SH + 4 * W * H * N_POSES

    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    gathered_ev = [None, None]

    n_chunks = 1 if world == 1 else 4     # N > 1: gather chunk k while chunk k+1 renders

    def step(i):
        buf = imgs[i & 1]
        cur = torch.cuda.current_stream()
        if world > 1 and gathered_ev[i & 1] is not None:
            cur.wait_event(gathered_ev[i & 1])      # this buffer's previous gather has finished
        if world == 1:
            render_batch(tree, cams, opt, buf)
            return
        # the one collective of the path: gather finished RGBA8 frames on rank 0.  The persistent
        # render kernel owns every SM while it runs, so the sweep is cut into a few launches and each
        # chunk's gather (own stream) overlaps the following chunk's rendering.
        per = (N_POSES + n_chunks - 1) // n_chunks
        ev = None
        for c0 in range(0, N_POSES, per):
            c1 = min(N_POSES, c0 + per)
            render_batch(tree, cams[c0:c1], opt, buf[c0:c1])
            done = torch.cuda.Event()
            done.record(cur)
            comm.wait_event(done)
            with torch.cuda.stream(comm):
                dist.gather(buf[c0:c1], [g[c0:c1] for g in gathered] if rank == 0 else None, dst=0)
                ev = torch.cuda.Event()
                ev.record(comm)
        gathered_ev[i & 1] = ev

    def sync_all():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for i in range(args.warmup):
        step(i)
    sync_all()
    launches0 = lib().vr_launch_count()
    cs = ClockSampler(local_rank)
    if rank == 0:
        cs.start()
    kern_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        step(i)
        k1.record()
        kern_ms.append((k0, k1))
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm)
    e1.record()
    sync_all()
    clocks = cs.stop() if rank == 0 else None
    launches = lib().vr_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    kms = float(np.mean([a.elapsed_time(b) for a, b in kern_ms]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps

    # ---- e2e: host-buffer entry point, one launch per pose + D2H of every frame
    for _ in range(2):
        render_frames_host(tree, cams, opt, host)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        render_frames_host(tree, cams, opt, host)
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())

    a_all = torch.tensor([float(a_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(a_all, op=dist.ReduceOp.SUM)

    if rank == 0:
        peak, peak_note = measured_peak()
        rays = W * H * N_POSES * world
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))["dram_bytes_per_frame"] * N_POSES
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": "Mrays/s @ 800x800", "value": rays / ms_step / 1e3, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "fps": N_POSES * world / ms_step * 1e3,
            "config": {"workload": WORKLOAD, "tree_depth": depth, "nodes": st.capacity,
                       "tree_bytes_device": info["node_bytes"] + info["rec_total_bytes"],
                       "views_per_step_per_gpu": N_POSES, "parallelism": f"views x{world}",
                       "kernel_variant": lib().vr_get_variant(),
                       "l2": "inputs larger than L2 (1.1 GB tree, a different pose every frame); no flush needed"},
            "roofline": {"bound": "hbm", "achieved": a_step / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": a_step / (kms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_note,
                         "algorithmic_bytes_per_launch": a_step, "kernel_ms_per_launch": kms,
                         "counters": {"samples": S, "child_loads": D, "shaded": SH, "rays_hit": HIT,
                                      "node_fetches": FETCH},
                         "note": "algorithmic bytes count the reference algorithm's touches (SURVEY 8d); "
                                 "this kernel skips most of the root-restart chain, see DESIGN.md"},
            "e2e": {"value": rays / e2e_ms / 1e3, "unit": "Mrays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 64 * N_POSES * world, "d2h_bytes_per_step": 4 * W * H * N_POSES * world},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            cb, samples = cpu_baseline(st, poses)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
