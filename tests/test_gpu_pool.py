"""`-m gpu`: the ray-pool kernel (variant 8, vr_march_q.cuh): parked rays are re-marched by other warps, so every
entry point that can reach it is compared bit for bit with the plain queue kernel (variant 7) at sizes where parking
really happens (many tiles per CTA), including surface output, tiles, bands, composite (parking disabled) and
back-to-back launches that share the pool memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cam(W, H, pose):
    from volrend_b200 import Camera, synth
    c = Camera(W, H, synth.focal_for(W), synth.focal_for(W))
    c.set_c2w(pose)
    return c


@pytest.mark.parametrize("basis,fmt", [(16, "SH"), (25, "SH"), (9, "SH"), (4, "SH"), (9, "SG")])
def test_pool_kernel_equals_queue_kernel(built, basis, fmt):
    import torch
    from volrend_b200 import N3Tree, RenderOptions, launch_renderer, lib, render_bands, render_batch, synth
    if not lib().vr_variant_supported(basis, 8):
        pytest.skip("the ray-pool kernel is an experiment: build with make lib EXTRA=-DVR_EXPERIMENTS")
    st = synth.make_tree("lego", depth=8, basis_dim=basis, seed=basis, fmt=fmt)
    tree = N3Tree.from_synth(st)
    W, H = 640, 480
    poses = synth.nerf_synthetic_test_poses(24)
    cams = [_cam(W, H, p) for p in poses]
    opts = [RenderOptions(), RenderOptions(stop_thresh=0.0, sigma_thresh=0.0), RenderOptions(render_depth=True),
            RenderOptions(step_size=1e-3, background_brightness=0.4)]
    res = {}
    for v in (7, 8):
        assert lib().vr_set_variant(v) == 0
        out = []
        for opt in opts:
            imgs = torch.zeros((len(cams), H, W, 4), dtype=torch.uint8, device="cuda")
            fo = torch.zeros((len(cams), H, W, 4), dtype=torch.float32, device="cuda")
            cnt = torch.zeros(5, dtype=torch.int64, device="cuda")
            render_batch(tree, cams, opt, imgs, float_out=fo)
            render_batch(tree, cams, opt, None, counters=cnt)
            torch.cuda.synchronize()
            out.append((fo.cpu().numpy(), imgs.cpu().numpy(), cnt.cpu().tolist()))
        # single-frame launches back to back on one stream (they share the pool memory and overlap through PDL)
        single = torch.zeros((6, H, W, 4), dtype=torch.uint8, device="cuda")
        for rep in range(2):
            for i in range(6):
                launch_renderer(tree, cams[i], opts[0], single[i], None, None, True)
        # a tile and a band launch
        tile = torch.zeros((200, 300, 4), dtype=torch.float32, device="cuda")
        launch_renderer(tree, cams[3], opts[0], None, None, None, True, float_out=tile, tile=(100, 150, 300, 200))
        band = torch.zeros((max(1, H // 3 + 8), W, 4), dtype=torch.uint8, device="cuda")
        rows = render_bands(tree, cams[5], opts[0], 8, 3, 1, band)
        # composite mode (parking is disabled there, the kernel must still be right)
        comp = torch.full((H, W, 4), 77, dtype=torch.uint8, device="cuda")
        depth = torch.full((H, W), 3.5, dtype=torch.float32, device="cuda")
        launch_renderer(tree, cams[7], opts[0], comp, depth, None, False)
        torch.cuda.synchronize()
        res[v] = (out, single.cpu().numpy(), tile.cpu().numpy(), band[:rows].cpu().numpy(), comp.cpu().numpy())
    lib().vr_set_variant(0)
    a, b = res[7], res[8]
    for (fa, ua, ca), (fb, ub, cb) in zip(a[0], b[0]):
        assert np.array_equal(fa, fb) and np.array_equal(ua, ub)
        assert ca == cb
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y)
    assert a[0][0][2][0] > 0 and (a[0][0][1][..., :3] != 255).any()      # something was rendered
