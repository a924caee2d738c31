#!/usr/bin/env python
"""Run-to-run stability of the batched launch: N repetitions, per-launch CUDA-event times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import synth, N3Tree, Camera, RenderOptions, render_batch, lib  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 9
st = synth.make_tree("lego", depth=depth, basis_dim=16)
tree = N3Tree.from_synth(st)
poses = synth.nerf_synthetic_test_poses(200)
for nv in (200, 40, 8):
    cams = []
    for p in poses[:: 200 // nv][:nv]:
        c = Camera(800, 800, synth.focal_for(800), synth.focal_for(800))
        c.set_c2w(p)
        cams.append(c)
    imgs = torch.zeros((nv, 800, 800, 4), dtype=torch.uint8, device="cuda")
    opt = RenderOptions()
    for _ in range(3):
        render_batch(tree, cams, opt, imgs)
    torch.cuda.synchronize()
    ts = []
    for rep in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        render_batch(tree, cams, opt, imgs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / nv)
    ts = np.array(ts)
    print(f"views {nv:4d}: ms/frame min {ts.min():.4f} median {np.median(ts):.4f} p90 {np.percentile(ts, 90):.4f} max {ts.max():.4f}  "
          f"outliers(>1.3x median) {(ts > 1.3 * np.median(ts)).sum()}/60  worst idx {int(ts.argmax())}", flush=True)
