#!/usr/bin/env python
"""Tree load -> device timings: plain upload + re-layout, and a quantised
file decoded on the GPU (ours) vs on the CPU (reference loader src/n3tree.cpp:279-340 through
oracle/_ref, and our numpy decode).  Writes gpurun_out/load_bench.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volrend_b200 import N3Tree, synth  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.cuda.init()
out = {}
st = synth.make_tree("lego", depth=depth, basis_dim=16)
out["nodes"] = st.capacity
out["raw_bytes"] = st.nbytes()


def best(fn, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del r
    return min(ts)


t = N3Tree()
t.load_npz(dict(child=st.child, data=st.data, offset=st.offset, invradius3=st.invradius3,
                data_dim=np.int64(st.data_dim), data_format=np.array(st.data_format)))


def plain():
    t.load_cuda()
    return None


out["plain_upload_relayout_s"] = best(plain)
out["plain_upload_GBps"] = st.nbytes() / out["plain_upload_relayout_s"] / 1e9
print("plain upload + re-layout: %.3f s (%.1f GB/s of raw tree)" % (out["plain_upload_relayout_s"], out["plain_upload_GBps"]), flush=True)
t.free_cuda()

npz = synth.quantise_tree(st, n_retain=1, seed=0)
path = "/tmp/quant_bench.npz"
np.savez(path, **npz)
out["quant_file_bytes"] = os.path.getsize(path)
tq = N3Tree(gpu_decode=True)
tq.load_npz(dict(npz))
out["quant_gpu_decode_upload_s"] = best(lambda: tq.load_cuda())
tq.free_cuda()
tc = N3Tree(gpu_decode=False)
t0 = time.perf_counter()
tc.load_npz(dict(npz))
out["quant_numpy_decode_s"] = time.perf_counter() - t0
out["quant_cpu_decoded_upload_s"] = best(lambda: tc.load_cuda())
tc.free_cuda()
print("quantised: GPU decode+upload %.3f s | numpy decode %.2f s + upload %.3f s" %
      (out["quant_gpu_decode_upload_s"], out["quant_numpy_decode_s"], out["quant_cpu_decoded_upload_s"]), flush=True)
t0 = time.perf_counter()
ours = N3Tree(path, gpu_decode=True)
torch.cuda.synchronize()
out["open_file_ours_gpu_decode_s"] = time.perf_counter() - t0
if rb.available():
    t0 = time.perf_counter()
    rt = rb.RefTree(path)
    torch.cuda.synchronize()
    out["open_file_reference_loader_s"] = time.perf_counter() - t0
    rt.close()
print("open(quantised file): ours %.2f s, reference loader %s s" %
      (out["open_file_ours_gpu_decode_s"], out.get("open_file_reference_loader_s")), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/load_bench.json", "w"), indent=1)
