"""Deterministic inputs of the golden-vector cases (shared by tools/make_golden.py, which runs the
reference CUDA kernel on them on the B200, and by the CPU/GPU parity tests)."""
import numpy as np

from volrend_b200 import synth

CASES = ["cfg1_sh1", "lego_sh16", "drums_sh9", "lego_sh25_bbox", "lego_sh4_stop0", "lego_rgba",
         "lego_sg9_rot", "lego_sh16_ndc", "lego_sh9_depth", "lego_asg4", "lego_sh9_composite"]


def composite_inputs(name: str, W: int, H: int):
    """Existing colour + per-pixel depth limit of the launch_renderer(offscreen=false) cases
    (volrend.cu:92-96,143-163); None for offscreen cases."""
    if name != "lego_sh9_composite":
        return None
    rng = np.random.default_rng(21)
    rgba = rng.integers(0, 256, (H, W, 4)).astype(np.uint8)
    depth = rng.uniform(2.0, 5.5, (H, W)).astype(np.float32)     # world units; cuts some rays inside the object
    depth[: H // 4] = 1e9                                        # a band without a depth limit
    return rgba, depth


def build_case(name: str):
    """-> (SynthTree, W, H, pose 4x4, options dict, ndc tuple | None)"""
    poses = synth.nerf_synthetic_test_poses(8)
    if name == "cfg1_sh1":
        return synth.make_config1_tree(), 64, 64, synth.config1_pose(), {}, None
    if name == "lego_sh16":
        return synth.make_tree("lego", depth=6, basis_dim=16, seed=11), 96, 96, poses[3], {}, None
    if name == "drums_sh9":
        return synth.make_tree("drums", depth=6, basis_dim=9, seed=12), 80, 64, poses[5], {}, None
    if name == "lego_sh25_bbox":
        return (synth.make_tree("lego", depth=5, basis_dim=25, seed=13), 64, 64, poses[1],
                dict(render_bbox=[0.1, 0.05, 0.0, 0.8, 1.0, 0.9], background_brightness=0.5), None)
    if name == "lego_sh4_stop0":
        return (synth.make_tree("lego", depth=5, basis_dim=4, seed=14), 64, 48, poses[6],
                dict(stop_thresh=0.0, sigma_thresh=0.0, step_size=1e-3), None)
    if name == "lego_rgba":
        return synth.make_tree("lego", depth=5, fmt="RGBA", seed=15), 64, 48, poses[2], {}, None
    if name == "lego_sg9_rot":
        return (synth.make_tree("lego", depth=5, basis_dim=9, fmt="SG", seed=16), 64, 48, poses[7],
                dict(rot_dirs=[0.3, -0.2, 0.5], basis_minmax=[1, 7]), None)
    if name == "lego_sh16_ndc":
        # forward-facing camera in front of the NDC frustum (maybe_world2ndc, volrend.cu:34-54)
        pose = np.eye(4, dtype=np.float32)
        pose[:3, 3] = [0.05, -0.03, 0.2]
        st = synth.make_tree("lego", depth=5, basis_dim=16, seed=17, world_radius=1.0)
        return st, 64, 48, pose, {}, (64.0, 48.0, 60.0)
    if name == "lego_sh9_depth":
        return synth.make_tree("lego", depth=5, basis_dim=9, seed=18), 64, 48, poses[4], dict(render_depth=1), None
    if name == "lego_asg4":
        # anisotropic spherical gaussians (lumisphere.hpp:14-28), lobes in extra_data
        return synth.make_tree("lego", depth=5, basis_dim=4, fmt="ASG", seed=19), 64, 48, poses[0], {}, None
    if name == "lego_sh9_composite":
        return synth.make_tree("lego", depth=5, basis_dim=9, seed=20), 64, 48, poses[3], {}, None
    raise KeyError(name)
