#!/usr/bin/env python
"""Small workload for compute-sanitizer: both product kernels (batch + single view, SH16 / SH25 / RGBA), bands, composite,
the multi-GPU C-ABI on a repeated device list.  usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import (Camera, MultiGpuRenderer, N3Tree, RenderOptions, VR_MG_TILES, VR_MG_VIEWS, launch_renderer, lib, render_bands,  # noqa: E402
                          render_batch, synth)

W, H = 200, 152
poses = synth.nerf_synthetic_test_poses(6)
tot = 0
for kw in (dict(basis_dim=16), dict(basis_dim=25), dict(fmt="RGBA"), dict(basis_dim=9, fmt="SG")):
    st = synth.make_tree("lego", depth=7, seed=3, **kw)
    tree = N3Tree.from_synth(st)
    cams = []
    for p in poses:
        c = Camera(W, H, synth.focal_for(W), synth.focal_for(W))
        c.set_c2w(p)
        cams.append(c)
    for v in (0, 7, 3 + 16 * 193):
        if not lib().vr_variant_supported(tree.info()["kernel_basis"], v):
            continue
        lib().vr_set_variant(v)
        imgs = torch.zeros((len(cams), H, W, 4), dtype=torch.uint8, device="cuda")
        render_batch(tree, cams, RenderOptions(), imgs)
        one = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        launch_renderer(tree, cams[2], RenderOptions(), one, None, None, True)
        band = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        render_bands(tree, cams[1], RenderOptions(), 8, 3, 2, band)
        depth = torch.full((H, W), 3.0, dtype=torch.float32, device="cuda")
        launch_renderer(tree, cams[3], RenderOptions(), one, depth, None, False)
        torch.cuda.synchronize()
        tot += int(imgs.long().sum()) + int(one.long().sum())
    lib().vr_set_variant(0)
    mg = MultiGpuRenderer(tree, [0, 0])
    host = np.zeros((len(cams), H, W, 4), np.uint8)
    for mode in (VR_MG_VIEWS, VR_MG_TILES):
        mg.render(cams, RenderOptions(), mode=mode, band_h=8, batch=4, out_host=host)
        tot += int(host.sum())
    mg.close()
print("sanitize target done, checksum", tot)
