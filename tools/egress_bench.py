#!/usr/bin/env python
"""Image egress: 200 poses at 800x800 -> pinned host frames / PNG files."""
import json
import os
import shutil
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import synth, N3Tree, Camera, RenderOptions, render_frames_host, render_frames_png  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
st = synth.make_tree("lego", depth=depth, basis_dim=16)
tree = N3Tree.from_synth(st)
cams = []
for p in synth.nerf_synthetic_test_poses(200):
    c = Camera(800, 800, synth.focal_for(800), synth.focal_for(800))
    c.set_c2w(p)
    cams.append(c)
opt = RenderOptions()
host = torch.empty((200, 800, 800, 4), dtype=torch.uint8).pin_memory()
out = {"cores": os.cpu_count()}
for _ in range(2):
    render_frames_host(tree, cams, opt, host)
t0 = time.perf_counter()
for _ in range(3):
    render_frames_host(tree, cams, opt, host)
out["frames_host_fps"] = 600 / (time.perf_counter() - t0)
d = "/tmp/vr_png_out"
for nt in (1, 8, 32):
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    paths = [f"{d}/{i:04d}.png" for i in range(200)]
    render_frames_png(tree, cams, opt, paths, n_threads=nt)
    t0 = time.perf_counter()
    render_frames_png(tree, cams, opt, paths, n_threads=nt)
    out[f"png_fps_{nt}_threads"] = 200 / (time.perf_counter() - t0)
out["png_bytes"] = os.path.getsize(paths[0])
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/egress_bench.json", "w"), indent=1)
