"""`-m "not gpu"`: host-side mirror of the reference loader / camera / options."""
import os

import numpy as np
import pytest

from volrend_b200.host import Camera, DataFormat, N3Tree, RenderOptions


@pytest.mark.parametrize("s,fmt,bd", [("SH16", DataFormat.SH, 16), ("SG25", DataFormat.SG, 25),
                                      ("ASG4", DataFormat.ASG, 4), ("RGBA", DataFormat.RGBA, -1),
                                      ("SH1", DataFormat.SH, 1), ("XY9", DataFormat.RGBA, 9)])
def test_data_format_parse(s, fmt, bd):
    # src/n3tree.cpp:55-78
    d = DataFormat()
    d.parse(s)
    assert (d.format, d.basis_dim) == (fmt, bd)
    if fmt != DataFormat.RGBA:
        assert d.to_string() == s


def test_render_options_defaults():
    o = RenderOptions()._as_c()
    assert abs(o.step_size - 1e-4) < 1e-10 and list(o.basis_minmax) == [0, 24] and o.render_depth == 0


def test_camera_transform_layout():
    # camera.cpp:47-55: columns right, up, back, centre; column-major 12 floats
    cam = Camera(64, 48, 100.0)
    assert cam.fy == 100.0
    t = cam.transform
    assert t.shape == (4, 3)
    np.testing.assert_allclose(np.linalg.norm(t[:3], axis=1), 1.0, atol=1e-6)
    np.testing.assert_allclose(t[3], [-3.55, 0.0, 3.55])
    np.testing.assert_allclose(np.cross(t[2], t[0]), t[1], atol=1e-6)   # up = back x right
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = [1, 2, 3]
    m[0, 1] = 0.5
    cam.set_c2w(m)
    c = cam._as_c()
    assert list(c.c2w)[9:12] == [1, 2, 3] and c.c2w[3] == 0.5       # element (row 0, col 1)


def _loader_only(npz) -> N3Tree:
    t = N3Tree()
    t.load_npz(npz)
    return t


def test_load_npz_plain(synth_mod, tmp_path):
    st = synth_mod.make_tree("lego", depth=4, basis_dim=9, seed=3)
    p = str(tmp_path / "t.npz")
    st.save_npz(p)
    with np.load(p) as z:
        t = _loader_only({k: z[k] for k in z.files})
    assert (t.N, t.data_dim, t.capacity) == (2, 28, st.capacity)
    assert t.data_format.to_string() == "SH9"
    np.testing.assert_array_equal(t.child_, st.child)
    np.testing.assert_array_equal(t.data_.view(np.uint16), st.data.view(np.uint16))
    np.testing.assert_allclose(t.scale, st.invradius3)


def test_load_npz_legacy_inference(synth_mod):
    # n3tree.cpp:240-254: no data_format key
    st = synth_mod.make_tree("lego", depth=3, basis_dim=4, seed=1)
    npz = dict(data_dim=np.int64(13), invradius=np.float64(0.25), offset=st.offset, child=st.child, data=st.data)
    t = _loader_only(npz)
    assert t.data_format.format == DataFormat.SH and t.data_format.basis_dim == 4
    np.testing.assert_allclose(t.scale, [0.25] * 3)
    st4 = synth_mod.make_tree("lego", depth=3, fmt="RGBA", seed=1)
    t = _loader_only(dict(data_dim=np.int64(4), invradius=np.float64(0.25), offset=st4.offset, child=st4.child, data=st4.data))
    assert t.data_format.format == DataFormat.RGBA


def test_load_npz_quantised_decode_matches_reference_loop(synth_mod):
    """Vectorised decode == the scalar loops of src/n3tree.cpp:309-340."""
    rng = np.random.default_rng(0)
    cap, N, n_basis, n_retain = 5, 2, 3, 1
    n_total = n_basis + n_retain
    data_dim = 3 * n_total + 1
    n_child = cap * N ** 3
    qc = rng.standard_normal((n_basis, 65536, 3)).astype(np.float16)
    qm = rng.integers(0, 65536, (n_basis, cap, N, N, N)).astype(np.uint16)
    sigma = rng.random((cap, N, N, N)).astype(np.float16)
    retained = rng.standard_normal((n_retain, cap, N, N, N, 3)).astype(np.float16)
    child = np.zeros((cap, N, N, N), np.int32)
    t = _loader_only(dict(data_dim=np.int64(data_dim), data_format=np.array(f"SH{n_total}"),
                          invradius3=np.ones(3, np.float32), offset=np.zeros(3, np.float32), child=child,
                          quant_colors=qc, quant_map=qm, sigma=sigma, data_retained=retained))
    want = np.zeros((n_child, data_dim), np.float16)
    qmf, sf, rf = qm.reshape(n_basis, n_child), sigma.reshape(n_child), retained.reshape(n_retain, n_child, 3)
    for i in range(n_child):
        for j in range(n_basis):
            boff = j + n_retain
            col = qc[j, qmf[j, i]]
            for k in range(3):
                want[i, boff] = col[k]
                boff += n_total
        want[i, data_dim - 1] = sf[i]
        for j in range(n_retain):
            boff = j
            for k in range(3):
                want[i, boff] = rf[j, i, k]
                boff += n_total
    np.testing.assert_array_equal(t.data_.reshape(n_child, data_dim).view(np.uint16), want.view(np.uint16))
    assert t.capacity == cap


def test_load_npz_rejects_non_half(synth_mod):
    st = synth_mod.make_tree("lego", depth=3, basis_dim=1, seed=1)
    with pytest.raises(RuntimeError, match="half precision"):
        _loader_only(dict(data_dim=np.int64(4), data_format=np.array("SH1"), invradius3=st.invradius3,
                          offset=st.offset, child=st.child, data=st.data.astype(np.float32)))


def test_missing_file_leaves_tree_unloaded(tmp_path, capsys):
    t = N3Tree()
    t.open(str(tmp_path / "nope.npz"))
    assert not t.is_data_loaded() and not t.is_cuda_loaded()
    assert "does not exist" in capsys.readouterr().out


def test_pose_files_roundtrip(synth_mod, tmp_path):
    poses = synth_mod.nerf_synthetic_test_poses(4)
    paths = synth_mod.write_pose_files(poses, str(tmp_path), 1111.11)
    back = np.loadtxt(paths[2]).reshape(4, 4)
    np.testing.assert_allclose(back, poses[2], rtol=1e-6, atol=1e-7)
    K = np.loadtxt(os.path.join(str(tmp_path), "intrinsics.txt"))
    assert abs(K[0, 0] - 1111.11) < 1e-3 and abs(K[1, 1] - 1111.11) < 1e-3
    np.testing.assert_allclose(np.linalg.norm(poses[:, :3, 3], axis=1), 4.031128874, rtol=1e-5)


def test_quantised_file_lazy_decode_equals_eager(synth_mod):
    st = synth_mod.make_tree("lego", depth=4, basis_dim=9, seed=5)
    npz = synth_mod.quantise_tree(st, n_retain=2, seed=1)
    lazy, eager = N3Tree(gpu_decode=True), N3Tree(gpu_decode=False)
    lazy.load_npz(dict(npz))
    eager.load_npz(dict(npz))
    assert lazy._data is None and lazy.quant_ is not None and eager._data is not None
    np.testing.assert_array_equal(lazy.data_.view(np.uint16), eager.data_.view(np.uint16))
    # sigma survives, and the retained functions land in the leading slots of each channel
    np.testing.assert_array_equal(eager.data_[..., -1].view(np.uint16), st.data[..., -1].view(np.uint16))
    np.testing.assert_array_equal(eager.data_[..., 9 + 1].view(np.uint16),
                                  npz["data_retained"][1][..., 1].view(np.uint16))


def test_png_writer_roundtrip(built, tmp_path):
    """vr_write_png (level-0 PNG without libpng, src/imwrite.cpp semantics): any PNG reader must decode
    exactly the RGBA8 bytes; sizes that exercise several 64 KB stored blocks and odd widths."""
    from PIL import Image
    from volrend_b200 import write_png_file
    rng = np.random.default_rng(0)
    for (h, w) in ((1, 1), (7, 13), (64, 64), (200, 333)):
        img = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
        p = str(tmp_path / f"t_{h}x{w}.png")
        assert write_png_file(p, img)
        back = np.asarray(Image.open(p))
        assert back.shape == (h, w, 4) and np.array_equal(back, img)
    assert not write_png_file(str(tmp_path / "nodir" / "x.png"), np.zeros((2, 2, 4), np.uint8))


def test_tile_decode_magic_division(tmp_path):
    """vr_types.h:set_div (tile index -> view / row by multiply-high + shift) against n // d,
    compiled as a host program with g++ and the CUDA headers (no GPU needed)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    src = tmp_path / "div.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdint>
#include "vr_types.h"
static uint32_t umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
int main() {
    uint64_t bad = 0, checked = 0;
    uint32_t ds[] = {1, 2, 3, 5, 7, 25, 100, 200, 255, 256, 257, 4999, 5000, 20000, 65535, 65536, 65537,
                     1048575, 1048576, 1048577, 16777215, 0x7fffffffu};
    for (uint32_t d : ds) {
        uint32_t mul; int32_t sh;
        vrb::set_div(d, mul, sh);
        auto q = [&](uint32_t n) { return sh < 0 ? n : (umulhi(n, mul) >> sh); };
        for (uint64_t k = 0; k < 4000; ++k) {   // multiples of d and their neighbours, and a stride sweep
            uint64_t c[] = {k * d, k * d + d - 1, k * d + 1, k * 536870u + 17u, 0x7fffffffull - k};
            for (uint64_t n : c) {
                if (n > 0x7fffffffull) continue;
                ++checked;
                if (q((uint32_t)n) != (uint32_t)(n / d)) ++bad;
            }
        }
    }
    std::printf("%llu %llu\n", (unsigned long long)checked, (unsigned long long)bad);
    return bad != 0;
}
''')
    exe = tmp_path / "div"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", cuda_inc, "-I", os.path.join(root, "include"),
                    "-I", os.path.join(root, "volrend_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    checked, bad = (int(v) for v in out.stdout.split())
    assert bad == 0 and checked > 100000
