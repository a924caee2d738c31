#!/usr/bin/env python
"""Per-tile timeline of ONE single-frame launch (vr_debug_trace): utilisation over time, tile
duration distribution, who finishes last.  Writes gpurun_out/trace_frame.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import synth, N3Tree, Camera, RenderOptions, lib  # noqa: E402
from volrend_b200._capi import check  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
st = synth.make_tree("lego", depth=depth, basis_dim=16)
tree = N3Tree.from_synth(st)
W = H = 800
poses = synth.nerf_synthetic_test_poses(200)
n_items = (W // 8) * (H // 4)
out = {}
os.makedirs("gpurun_out", exist_ok=True)
for pi in (17, 120):
    cam = Camera(W, H, synth.focal_for(W), synth.focal_for(W))
    cam.set_c2w(poses[pi])
    img = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(5, dtype=torch.int64, device="cuda")
    tr = torch.zeros((n_items, 4), dtype=torch.int64, device="cuda")
    c, o = cam._as_c(), RenderOptions()._as_c()
    for rep in range(3):
        tr.zero_()
        check(lib().vr_debug_trace(tree._handle, C.byref(c), C.byref(o), img.data_ptr(), cnt.data_ptr(), tr.data_ptr(), None))
        torch.cuda.synchronize()
    t = tr.cpu().numpy()
    t0 = t[:, 0].min()
    beg, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3       # us
    dur = end - beg
    span = end.max()
    warps = np.unique(t[:, 3])
    per_warp_busy = np.array([dur[t[:, 3] == w].sum() for w in warps])
    per_warp_n = np.array([(t[:, 3] == w).sum() for w in warps])
    grid = np.arange(0, span, 5.0)
    active = [int(((beg <= g) & (end > g)).sum()) for g in grid]
    order = np.argsort(-dur)
    res = dict(makespan_us=float(span), n_warps=int(len(warps)), sum_tile_us=float(dur.sum()),
               ideal_us=float(dur.sum() / len(warps)), tile_us_pcts={p: float(np.percentile(dur, p)) for p in (10, 50, 90, 99, 100)},
               per_warp_busy_us=dict(min=float(per_warp_busy.min()), mean=float(per_warp_busy.mean()), max=float(per_warp_busy.max())),
               tiles_per_warp=dict(min=int(per_warp_n.min()), max=int(per_warp_n.max())),
               active_warps_every_5us=active,
               last_finishers=[dict(item=int(i), ty=int(i // (W // 8)), beg=float(beg[i]), dur=float(dur[i])) for i in np.argsort(-end)[:8]],
               longest=[dict(item=int(i), beg=float(beg[i]), dur=float(dur[i])) for i in order[:8]])
    out[f"pose{pi}"] = res
    np.savez_compressed(f"gpurun_out/trace_frame_pose{pi}.npz", beg=beg.astype(np.float32), dur=dur.astype(np.float32), warp=t[:, 3].astype(np.int32),
                        sm=t[:, 2].astype(np.int32))
    print(json.dumps(res)[:1500])
json.dump(out, open("gpurun_out/trace_frame.json", "w"), indent=1)
