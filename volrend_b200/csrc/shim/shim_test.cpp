// shim_test.cpp -- exercises the drop-in C++ surface exactly as reference callers do:
//   shim_test <tree.npz> <pose.txt> <w> <h> <fx> <out_launch.rgba> <out_vr.rgba>
// 1. main_headless.cpp-style: N3Tree(path), Camera, launch_renderer into a cudaArray.
// 2. main.cpp-style: VolumeRenderer rend; rend.set(tree); rend.resize(w,h); rend.render().
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "volrend/common.hpp"
#include "volrend/cuda/common.cuh"
#include "volrend/cuda/renderer_kernel.hpp"
#include "volrend/n3tree.hpp"
#include "volrend/renderer.hpp"
#include "volrend_b200_shim.hpp"

using namespace volrend;

static bool read_pose(const char* path, glm::mat4x3& m) {
    std::ifstream ifs(path);
    if (!ifs) return false;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) ifs >> m[c][r];
    return bool(ifs);
}
static void dump(const char* path, const std::vector<uint8_t>& buf) {
    FILE* f = fopen(path, "wb");
    fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: shim_test tree.npz pose.txt w h fx out1 out2\n"); return 2; }
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    const float fx = (float)atof(argv[5]);
    glm::mat4x3 pose;
    if (!read_pose(argv[2], pose)) { fprintf(stderr, "bad pose file\n"); return 2; }
    N3Tree tree(argv[1]);
    if (!tree.is_data_loaded() || !tree.is_cuda_loaded()) { fprintf(stderr, "tree not loaded\n"); return 3; }
    std::vector<uint8_t> buf((size_t)4 * w * h);
    {
        Camera camera(w, h, fx, fx);
        camera.transform = pose;
        camera._update(false);
        cudaArray_t array;
        cudaStream_t stream;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
        cuda(MallocArray(&array, &cd, w, h, cudaArraySurfaceLoadStore));
        cuda(StreamCreateWithFlags(&stream, cudaStreamDefault));
        cudaArray_t depth_arr = nullptr;
        RenderOptions options;
        launch_renderer(tree, camera, options, array, depth_arr, stream, true);
        cuda(Memcpy2DFromArrayAsync(buf.data(), 4 * w, array, 0, 0, 4 * w, h, cudaMemcpyDeviceToHost, stream));
        cuda(StreamSynchronize(stream));
        dump(argv[6], buf);
        cuda(FreeArray(array));
        cuda(StreamDestroy(stream));
    }
    {
        VolumeRenderer rend;
        rend.set(tree);
        rend.resize(w, h);
        rend.camera.fx = rend.camera.fy = fx;
        // express the pose through the camera's own pose model (camera.hpp:31-37)
        rend.camera.v_back = pose[2];
        rend.camera.v_world_up = pose[1];
        rend.camera.center = pose[3];
        rend.render();
        if (!volrend_b200_read_pixels(rend, buf.data())) { fprintf(stderr, "read_pixels failed\n"); return 4; }
        dump(argv[7], buf);
        printf("backend %s\n", rend.get_backend());
    }
    return 0;
}
