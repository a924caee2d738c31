#!/usr/bin/env python
"""Generate tests/golden/*.npz: outputs of the UNMODIFIED reference CUDA renderer (oracle/_ref,
built from /root/reference by oracle/Makefile.ref) on the deterministic cases of
tests/golden_cases.py.  Needs a GPU: run under gpurun, it writes gpurun_out/golden/ which is then
copied to tests/golden/ and committed.

  ref_f32  float RGBA of the reference's per-pixel code (tap of out[4], see oracle/ref_harness.cu)
  ref_u8   bytes written by volrend::launch_renderer (src/cuda/volrend.cu:166-172)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import CASES, build_case, composite_inputs  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402
from volrend_b200 import synth  # noqa: E402


def write_case_files(name, st, ndc, tmpdir="/tmp/golden_in"):
    os.makedirs(tmpdir, exist_ok=True)
    path = os.path.join(tmpdir, name + ".npz")
    st.save_npz(path)
    pb = path[:-4] + "_poses_bounds.npy"
    if ndc is not None:
        p = np.zeros((1, 17), np.float32)
        p[0, 0] = p[0, 6] = p[0, 12] = 1.0          # identity rotation
        p[0, 9], p[0, 4], p[0, 14] = ndc              # width, height, focal (n3tree.cpp:27-29)
        p[0, 15], p[0, 16] = 1.0, 10.0
        np.save(pb, p)
    elif os.path.exists(pb):
        os.remove(pb)
    return path


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        st, W, H, pose, optkw, ndc = build_case(name)
        path = write_case_files(name, st, ndc)
        rt = rb.RefTree(path)
        info = rt.info()
        assert bool(info["use_ndc"]) == (ndc is not None)
        fx = synth.focal_for(W)
        c12 = synth.c2w_to_colmajor12(pose)
        opt = rb.make_options(**optkw)
        comp = composite_inputs(name, W, H)
        rin, din = comp if comp is not None else (None, None)
        f = rt.render_f32(W, H, fx, fx, c12, opt, rgba_in=rin, depth_in=din)
        u = rt.render_u8(W, H, fx, fx, c12, opt, rgba_in=rin, depth_in=din)
        rt.close()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), case=np.array(name), ref_f32=f, ref_u8=u)
        print(name, info, "alpha mean %.3f" % f[..., 3].mean(), "rgb mean %.3f" % f[..., :3].mean(), flush=True)


if __name__ == "__main__":
    main()
