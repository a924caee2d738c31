#!/usr/bin/env python
"""Tiny driver for ncu: builds the bench tree once and launches a few batches of views with the
chosen kernel variant.  usage: profile_target.py [variant] [n_views] [reps] [depth] [basis]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import synth, N3Tree, Camera, RenderOptions, render_batch, launch_renderer, lib  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 10
basis = int(sys.argv[5]) if len(sys.argv) > 5 else 16
lib().vr_set_variant(variant)
st = synth.make_tree("lego", depth=depth, basis_dim=basis)
tree = N3Tree.from_synth(st)
W = H = 800
poses = synth.nerf_synthetic_test_poses(200)[:: max(1, 200 // n_views)][:n_views]
cams = []
for p in poses:
    c = Camera(W, H, synth.focal_for(W), synth.focal_for(W))
    c.set_c2w(p)
    cams.append(c)
imgs = torch.zeros((len(cams), H, W, 4), dtype=torch.uint8, device="cuda")
opt = RenderOptions()
for _ in range(reps):
    if n_views == 1:
        launch_renderer(tree, cams[0], opt, imgs[0], None, None, True)
    else:
        render_batch(tree, cams, opt, imgs)
torch.cuda.synchronize()
print("ok variant", lib().vr_get_variant(), "views", len(cams), "checksum", int(imgs.long().sum()))
