#!/usr/bin/env python
"""Variant sweep for one gpurun call: bit-exactness of every kernel variant against the inline-shading kernel on a
small tree, then batch (200 views / launch) and per-frame (1 view / launch, C-ABI called with
prebuilt structs) timings on the bench tree.  usage: gpu_variants.py v1,v2,... [depth]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import synth, N3Tree, Camera, RenderOptions, launch_renderer, render_batch, render_frames_host, lib  # noqa: E402
from volrend_b200 import _capi  # noqa: E402

variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "7,3091").split(",")]
INLINE = 3 + 16 * 193
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
res = {}

# ---- parity of variants (bit-exact vs variant 1)
st = synth.make_tree("lego", depth=7, basis_dim=16, seed=3)
tree = N3Tree.from_synth(st)
cam = Camera(256, 200, synth.focal_for(256), synth.focal_for(256))
cam.set_c2w(synth.nerf_synthetic_test_poses(8)[3])
ref = None
for v in [INLINE] + variants:
    assert lib().vr_set_variant(v) == 0, v
    fo = torch.zeros((200, 256, 4), dtype=torch.float32, device=dev)
    img = torch.zeros((200, 256, 4), dtype=torch.uint8, device=dev)
    launch_renderer(tree, cam, RenderOptions(), img, None, None, True, float_out=fo)
    torch.cuda.synchronize()
    f = fo.cpu().numpy()
    if ref is None:
        ref = f
    else:
        ok = bool(np.array_equal(f, ref))
        res.setdefault("bit_exact_vs_inline", {})[v] = ok
        print("variant", v, "bit-exact vs inline-shading kernel:", ok, flush=True)
del tree

# ---- timing
st = synth.make_tree("lego", depth=depth, basis_dim=16)
tree = N3Tree.from_synth(st)
W = H = 800
poses = synth.nerf_synthetic_test_poses(200)
cams = []
for p in poses:
    c = Camera(W, H, synth.focal_for(W), synth.focal_for(W))
    c.set_c2w(p)
    cams.append(c)
opt = RenderOptions()
imgs = torch.zeros((len(cams), H, W, 4), dtype=torch.uint8, device=dev)
host = torch.empty((len(cams), H, W, 4), dtype=torch.uint8).pin_memory()
ccams = [c._as_c() for c in cams]
copt = opt._as_c()
stream = torch.cuda.current_stream().cuda_stream
for v in variants:
    lib().vr_set_variant(v)
    r = {}
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        render_batch(tree, cams, opt, imgs)
        e1.record()
        torch.cuda.synchronize()
        r["batch_ms_per_frame"] = e0.elapsed_time(e1) / len(cams)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, cc in enumerate(ccams):
            lib().vr_render(tree._handle, C.byref(cc), C.byref(copt), None, imgs[0].data_ptr(), None, None, stream)
        e1.record()
        torch.cuda.synchronize()
        r["per_frame_ms"] = e0.elapsed_time(e1) / len(cams)
    for rep in range(2):
        t0 = time.perf_counter()
        render_frames_host(tree, cams, opt, host)
        r["frames_host_ms_per_frame"] = (time.perf_counter() - t0) * 1e3 / len(cams)
    res[f"v{v}"] = r
    print(f"variant {v:3d}: batch {r['batch_ms_per_frame']:.4f} ms/frame ({W*H/r['batch_ms_per_frame']/1e3:.0f} Mrays/s)  "
          f"per-frame launches {r['per_frame_ms']:.4f}  frames_host {r['frames_host_ms_per_frame']:.4f}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/variants.json", "w"), indent=1)
