// volrend_headless_mg -- multi-GPU sibling of the reference's volrend_headless (main_headless.cpp).
//
// Same command line (npz file, c2w pose files, -w/-h/--fx/--fy/-i/-o/--scale/--max_imgs/-r and the common
// render options of src/opts.cpp) plus
//     --gpus N        devices 0..N-1 of this box (default: all visible)
//     --mode views    pose i -> GPU i % N                                (throughput of a pose sweep)
//            tiles    every frame cut into bands of --band rows, band b -> GPU b % N   (one frame on N GPUs)
//     --batch K       poses per batch, i.e. per kernel launch on every GPU (default: a quarter of the poses in views mode so
//                     that copies overlap rendering, all poses in tiles mode)
//     --reps R        timed repetitions of the whole pose list (default 1; the first, untimed pass warms up)
//     --check         also render everything on GPU 0 alone and compare the bytes
// The reference's loader (src/n3tree.cpp + cnpy), option parser (src/opts.cpp) and headers are used
// unchanged; the tree's host arrays go through the C-ABI (vr_mg_create) to every GPU.  Timing is printed in
// the reference's format (main_headless.cpp:230-231) from CUDA events, max over GPUs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <glm/mat4x3.hpp>
#include <glm/mat4x4.hpp>

#include "volrend/internal/auto_filesystem.hpp"

#include "volrend/common.hpp"
#include "volrend/n3tree.hpp"

#include "volrend/internal/opts.hpp"

#include "volrend/cuda/common.cuh"

#include "volrend_b200.h"

namespace volrend {
// The loader calls these at the end of open() (src/n3tree.cpp:107,150-152): this binary keeps the tree on
// the host and hands it to vr_mg_create, which uploads it to every GPU.
void N3Tree::load_cuda() { cuda_loaded_ = true; }
void N3Tree::free_cuda() {}
cudaError_t cuda_assert(const cudaError_t code, const char* const file, const int line, const bool abort) {
    if (code != cudaSuccess) {
        fprintf(stderr, "cuda_assert: %s %s %d\n", cudaGetErrorString(code), file, line);
        if (abort) {
            cudaDeviceReset();
            exit(code);
        }
    }
    return code;
}
}  // namespace volrend

namespace {

std::string base_name(const std::string& s) {
    const size_t p = s.find_last_of("/\\");
    std::string b = p == std::string::npos ? s : s.substr(p + 1);
    const size_t d = b.find_last_of('.');
    return d == std::string::npos ? b : b.substr(0, d);
}

// one or more row-major 4x4 c2w matrices per file (main_headless.cpp:41-63)
int read_poses(const std::string& path, std::vector<glm::mat4x3>& out) {
    std::ifstream ifs(path);
    if (!ifs) {
        fprintf(stderr, "ERROR: '%s' does not exist\n", path.c_str());
        std::exit(1);
    }
    int cnt = 0;
    for (;;) {
        float m[16];
        int got = 0;
        while (got < 12 && (ifs >> m[got])) ++got;
        if (got < 12) break;
        for (int k = 12; k < 16; ++k) ifs >> m[k];
        glm::mat4x3 t;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) t[c][r] = m[r * 4 + c];
        out.push_back(t);
        ++cnt;
    }
    return cnt;
}

void read_intrins(const std::string& path, float& fx, float& fy) {
    std::ifstream ifs(path);
    if (!ifs) {
        fprintf(stderr, "ERROR: intrin '%s' does not exist\n", path.c_str());
        std::exit(1);
    }
    float g;
    ifs >> fx >> g >> g >> g;
    ifs >> g >> fy;
}

[[noreturn]] void die(const char* what, const vr_mg* mg) {
    fprintf(stderr, "volrend_headless_mg: %s: %s\n", what, vr_mg_last_error(mg));
    std::exit(1);
}

}  // namespace

int main(int argc, char* argv[]) {
    using namespace volrend;
    cxxopts::Options cxxoptions("volrend_headless_mg", "Headless PlenOctree rendering on several GPUs (volrend_b200 backend)");
    internal::add_common_opts(cxxoptions);
    // clang-format off
    cxxoptions.add_options()
        ("o,write_images", "output directory of images; if empty, DOES NOT save (for timing only)",
                cxxopts::value<std::string>()->default_value(""))
        ("i,intrin", "intrinsics matrix 4x4; if set, overrides the fx/fy", cxxopts::value<std::string>()->default_value(""))
        ("r,reverse_yz", "use OpenCV camera space convention instead of NeRF", cxxopts::value<bool>())
        ("scale", "scaling to apply to image", cxxopts::value<float>()->default_value("1.0"))
        ("max_imgs", "max images to render, default no limit", cxxopts::value<int>()->default_value("0"))
        ("gpus", "number of GPUs (devices 0..N-1); 0 = all visible", cxxopts::value<int>()->default_value("0"))
        ("mode", "views | tiles", cxxopts::value<std::string>()->default_value("views"))
        ("band", "tiles mode: rows per band (multiple of 4)", cxxopts::value<int>()->default_value("8"))
        ("batch", "poses per launch per GPU; 0 = all", cxxopts::value<int>()->default_value("0"))
        ("reps", "timed repetitions of the pose list", cxxopts::value<int>()->default_value("1"))
        ("check", "compare with a single-GPU render of the same poses", cxxopts::value<bool>())
        ;
    // clang-format on
    cxxoptions.allow_unrecognised_options();
    cxxoptions.positional_help("npz_file [c2w_txt_4x4...]");
    cxxopts::ParseResult args = internal::parse_options(cxxoptions, argc, argv);

    std::vector<glm::mat4x3> trans;
    std::vector<std::string> basenames;
    for (auto path : args.unmatched()) {
        const int cnt = read_poses(path, trans);
        const std::string fname = base_name(path);
        if (cnt == 1) {
            basenames.push_back(fname);
        } else {
            for (int i = 0; i < cnt; ++i) {
                std::string tmp = std::to_string(i);
                while (tmp.size() < 6) tmp = "0" + tmp;
                basenames.push_back(fname + "_" + tmp);
            }
        }
    }
    if (args["reverse_yz"].as<bool>()) {
        glm::mat4x4 cam_trans(1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1);
        for (auto& t : trans) t = t * cam_trans;
    }
    if (trans.empty()) {
        fputs("WARNING: No camera poses specified, quitting\n", stderr);
        return 1;
    }
    const std::string out_dir = args["write_images"].as<std::string>();

    N3Tree tree(args["file"].as<std::string>());
    puts("");   // the loader's last line has no newline (src/n3tree.cpp:264)
    if (!tree.is_data_loaded()) {
        fputs("ERROR: tree could not be loaded\n", stderr);
        return 1;
    }
    int width = args["width"].as<int>(), height = args["height"].as<int>();
    float fx = args["fx"].as<float>();
    if (fx < 0) fx = 1111.11f;
    float fy = args["fy"].as<float>();
    if (fy < 0) fy = fx;
    {
        const std::string intrin_path = args["intrin"].as<std::string>();
        if (intrin_path.size()) read_intrins(intrin_path, fx, fy);
    }
    {
        const float scale = args["scale"].as<float>();
        if (scale != 1.f) {
            const int ow = width, oh = height;
            width *= scale;
            height *= scale;
            fx *= (float)width / ow;
            fy *= (float)height / oh;
        }
    }
    {
        const int max_imgs = args["max_imgs"].as<int>();
        if (max_imgs > 0 && trans.size() > (size_t)max_imgs) {
            trans.resize(max_imgs);
            basenames.resize(max_imgs);
        }
    }

    int n_visible = 0;
    cuda(GetDeviceCount(&n_visible));
    int n_gpus = args["gpus"].as<int>();
    if (n_gpus <= 0 || n_gpus > n_visible) n_gpus = n_visible;
    const std::string mode_s = args["mode"].as<std::string>();
    const int mode = mode_s == "tiles" ? VR_MG_TILES : VR_MG_VIEWS;
    const int band = args["band"].as<int>();
    int batch = args["batch"].as<int>();
    // default: pose sweeps go out in four batches, so that the copy of one batch travels while the next one renders;
    // ray tiles are cheap to move (each GPU sends 1/N of a frame) and go out in one launch
    if (batch <= 0 && mode == VR_MG_VIEWS && (int)trans.size() >= 8 * n_gpus) batch = ((int)trans.size() + 3) / 4;
    const int reps = args["reps"].as<int>() > 0 ? args["reps"].as<int>() : 1;

    vr_tree_desc d;
    memset(&d, 0, sizeof(d));
    d.child = tree.child_.data<int32_t>();
    d.data = reinterpret_cast<const uint16_t*>(tree.data_.data<__half>());
    d.extra = tree.extra_.data_holder.size() ? tree.extra_.data<float>() : nullptr;
    d.capacity = tree.capacity; d.N = tree.N; d.data_dim = tree.data_dim;
    d.format = (int)tree.data_format.format; d.basis_dim = tree.data_format.basis_dim;
    for (int i = 0; i < 3; ++i) { d.offset[i] = tree.offset[i]; d.scale[i] = tree.scale[i]; }
    d.use_ndc = tree.use_ndc ? 1 : 0;
    d.ndc_width = tree.ndc_width; d.ndc_height = tree.ndc_height; d.ndc_focal = tree.ndc_focal;

    std::vector<int> devices(n_gpus);
    for (int i = 0; i < n_gpus; ++i) devices[i] = i;
    vr_mg* mg = nullptr;
    if (vr_mg_create(&d, devices.data(), n_gpus, &mg) != VR_OK) die("vr_mg_create", nullptr);

    const RenderOptions ro = internal::render_options_from_args(args);
    vr_options o;
    o.step_size = ro.step_size; o.sigma_thresh = ro.sigma_thresh; o.stop_thresh = ro.stop_thresh;
    o.background_brightness = ro.background_brightness;
    for (int i = 0; i < 6; ++i) o.render_bbox[i] = ro.render_bbox[i];
    o.basis_minmax[0] = ro.basis_minmax[0]; o.basis_minmax[1] = ro.basis_minmax[1];
    for (int i = 0; i < 3; ++i) o.rot_dirs[i] = ro.rot_dirs[i];
    o.render_depth = ro.render_depth ? 1 : 0;

    std::vector<vr_camera> cams(trans.size());
    for (size_t i = 0; i < trans.size(); ++i) {
        cams[i].width = width; cams[i].height = height; cams[i].fx = fx; cams[i].fy = fy;
        memcpy(cams[i].c2w, &trans[i][0][0], 12 * sizeof(float));
    }
    const int n_views = (int)cams.size();
    const size_t frame = (size_t)4 * width * height;

    uint8_t* host = nullptr;
    const bool want_host = out_dir.size() || args["check"].as<bool>();
    if (want_host) cuda(MallocHost((void**)&host, frame * n_views));

    float ms = 0.f, ms_sum = 0.f;
    if (vr_mg_render(mg, cams.data(), n_views, &o, mode, band, batch, nullptr, nullptr, &ms) != VR_OK) die("vr_mg_render (warm-up)", mg);
    for (int r = 0; r < reps; ++r) {
        // like the reference's timed loop without -o (main_headless.cpp:208-223): frames stay on the GPU
        if (vr_mg_render(mg, cams.data(), n_views, &o, mode, band, batch, nullptr, nullptr, &ms) != VR_OK) die("vr_mg_render", mg);
        ms_sum += ms;
    }
    if (host && vr_mg_render(mg, cams.data(), n_views, &o, mode, band, batch, nullptr, host, &ms) != VR_OK)   // untimed: -o / --check
        die("vr_mg_render (read-back)", mg);
    const float ms_frame = ms_sum / reps / n_views;
    printf("%.10f ms per frame\n", ms_frame);
    printf("%.10f fps\n", 1000.f / ms_frame);
    printf("%.3f Mrays/s on %d GPU(s), mode %s, %d poses of %dx%d, batch %d\n", (double)width * height / ms_frame / 1e3, n_gpus,
           mode == VR_MG_TILES ? "tiles" : "views", n_views, width, height, batch > 0 ? batch : n_views);

    int rc = 0;
    if (args["check"].as<bool>()) {
        // the same poses on GPU 0 alone, straight through vr_render_batch
        cuda(SetDevice(0));
        uint8_t* solo = nullptr;
        cuda(Malloc((void**)&solo, frame * n_views));
        if (vr_render_batch(vr_mg_tree(mg, 0), cams.data(), n_views, &o, nullptr, solo, nullptr, nullptr, nullptr) != VR_OK) {
            fprintf(stderr, "vr_render_batch: %s\n", vr_last_error());
            return 1;
        }
        std::vector<uint8_t> ref(frame * n_views);
        cuda(Memcpy(ref.data(), solo, ref.size(), cudaMemcpyDeviceToHost));
        cuda(Free(solo));
        size_t bad = 0;
        for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != host[i];
        printf("check: %zu of %zu bytes differ from the single-GPU render%s\n", bad, ref.size(), bad ? "  <-- MISMATCH" : " (identical)");
        rc = bad ? 2 : 0;
    }
    if (out_dir.size()) {
        std::filesystem::create_directories(out_dir);
        for (int i = 0; i < n_views; ++i) {
            const std::string fpath = out_dir + "/" + basenames[i] + ".png";
            if (vr_write_png(fpath.c_str(), host + (size_t)i * frame, width, height) != VR_OK) {
                fprintf(stderr, "vr_write_png(%s): %s\n", fpath.c_str(), vr_last_error());
                rc = 1;
            }
        }
    }
    if (host) cudaFreeHost(host);
    vr_mg_destroy(mg);
    return rc;
}
