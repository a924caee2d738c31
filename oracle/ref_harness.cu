// ref_harness.cu -- C-ABI around the UNMODIFIED reference CUDA renderer (test infrastructure).
//
// Built only where /root/reference exists (oracle/Makefile.ref), into oracle/_ref/.  No
// reference source is copied: this TU #includes the reference's own src/cuda/volrend.cu so
// that (a) `volrend::launch_renderer` is the reference's stock code path and (b) its
// anonymous-namespace helpers and `device::trace_ray` are visible to the float tap below.
//
//   ref_render_u8   -> volrend::launch_renderer (src/cuda/volrend.cu:195-245), bytes exactly
//                      as main_headless.cpp:214-219 would read them back
//   ref_render_f32  -> the un-quantised float out[4] of the same per-pixel code: a tap kernel
//                      that calls the reference's screen2worlddir / maybe_world2ndc /
//                      rodrigues / device::trace_ray in the order of volrend.cu:136-158
//   ref_time_frames -> the timed pose loop of main_headless.cpp:203-228
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../src/cuda/volrend.cu"  // resolved through -I$(VOLREND_REF)/include/.. (see Makefile.ref)

#include "volrend/camera.hpp"
#include "volrend/n3tree.hpp"

namespace volrend {
namespace device {
// Float tap: same statements as render_kernel (volrend.cu:136-158, offscreen branch), storing
// out[4] instead of truncating to bytes (volrend.cu:166).
__global__ static void tap_kernel(CameraSpec cam, TreeSpec tree, RenderOptions opt, float4* out4,
                                  const uint8_t* rgba_in = nullptr, const float* depth_in = nullptr) {
    CUDA_GET_THREAD_ID(idx, cam.width * cam.height);
    const int x = idx % cam.width, y = idx / cam.width;
    float dir[3], cen[3], out[4];
    const bool offscreen = rgba_in == nullptr;   // volrend.cu:90-96: existing colour for compositing
    uint8_t rgbx_init[4] = {0, 0, 0, 0};
    if (!offscreen) {
        for (int i = 0; i < 4; ++i) rgbx_init[i] = rgba_in[4 * idx + i];
    }
    out[0] = out[1] = out[2] = out[3] = 0.f;
    if (tree.N > 0) {
        screen2worlddir(x, y, cam, dir, cen);
        float vdir[3] = {dir[0], dir[1], dir[2]};
        maybe_world2ndc(tree, dir, cen);
        for (int i = 0; i < 3; ++i) {
            cen[i] = tree.offset[i] + tree.scale[i] * cen[i];
        }
        float t_max = 1e9f;
        if (!offscreen) t_max = depth_in[idx];   // volrend.cu:143-146
        rodrigues(opt.rot_dirs, vdir);
        trace_ray(tree, dir, vdir, cen, opt, t_max, out);
    }
    const float nalpha = 1.f - out[3];
    if (offscreen) {
        const float remain = opt.background_brightness * nalpha;
        out[0] += remain;
        out[1] += remain;
        out[2] += remain;
    } else {   // volrend.cu:159-163
        out[0] += rgbx_init[0] / 255.f * nalpha;
        out[1] += rgbx_init[1] / 255.f * nalpha;
        out[2] += rgbx_init[2] / 255.f * nalpha;
    }
    out4[idx] = make_float4(out[0], out[1], out[2], out[3]);
}
}  // namespace device
}  // namespace volrend

using namespace volrend;

extern "C" {

struct ref_options {
    float step_size, sigma_thresh, stop_thresh, background_brightness;
    float render_bbox[6];
    int32_t basis_minmax[2];
    float rot_dirs[3];
    int32_t render_depth;
};

static RenderOptions to_opts(const ref_options* o) {
    RenderOptions r;
    r.step_size = o->step_size; r.sigma_thresh = o->sigma_thresh; r.stop_thresh = o->stop_thresh;
    r.background_brightness = o->background_brightness;
    for (int i = 0; i < 6; ++i) r.render_bbox[i] = o->render_bbox[i];
    r.basis_minmax[0] = o->basis_minmax[0]; r.basis_minmax[1] = o->basis_minmax[1];
    for (int i = 0; i < 3; ++i) r.rot_dirs[i] = o->rot_dirs[i];
    r.render_depth = o->render_depth != 0;
    return r;
}

static void set_cam(Camera& cam, const float* c2w12) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 3; ++r) cam.transform[c][r] = c2w12[c * 3 + r];
    cam._update(false);
}

void* ref_tree_open(const char* npz_path) {
    N3Tree* t = new N3Tree(std::string(npz_path));
    if (!t->is_data_loaded()) { delete t; return nullptr; }
    return t;
}
void ref_tree_close(void* t) { delete static_cast<N3Tree*>(t); }
int ref_tree_info(void* tp, int* N, int* data_dim, int* basis_dim, int* format, long long* capacity, int* use_ndc) {
    N3Tree* t = static_cast<N3Tree*>(tp);
    *N = t->N; *data_dim = t->data_dim; *basis_dim = t->data_format.basis_dim;
    *format = (int)t->data_format.format; *capacity = t->capacity; *use_ndc = t->use_ndc;
    return 0;
}

// Reference launch_renderer into an RGBA8 cudaArray, copied back like main_headless.cpp:216-219.
// rgba_in/depth_in non-null selects offscreen=false (colour + depth compositing inputs).
int ref_render_u8(void* tp, int w, int h, float fx, float fy, const float* c2w12, const ref_options* o,
                  const uint8_t* rgba_in, const float* depth_in, uint8_t* out_host) {
    N3Tree& tree = *static_cast<N3Tree*>(tp);
    Camera cam(w, h, fx, fy);
    set_cam(cam, c2w12);
    RenderOptions opt = to_opts(o);
    cudaArray_t arr = nullptr, darr = nullptr;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    cuda(MallocArray(&arr, &cd, w, h, cudaArraySurfaceLoadStore));
    const bool offscreen = rgba_in == nullptr;
    if (!offscreen) {
        cudaChannelFormatDesc dd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        cuda(MallocArray(&darr, &dd, w, h, cudaArraySurfaceLoadStore));
        cuda(Memcpy2DToArray(arr, 0, 0, rgba_in, 4 * w, 4 * w, h, cudaMemcpyHostToDevice));
        cuda(Memcpy2DToArray(darr, 0, 0, depth_in, 4 * w, 4 * w, h, cudaMemcpyHostToDevice));
    }
    cudaStream_t stream;
    cuda(StreamCreateWithFlags(&stream, cudaStreamDefault));
    launch_renderer(tree, cam, opt, arr, darr, stream, offscreen);
    cuda(Memcpy2DFromArrayAsync(out_host, 4 * w, arr, 0, 0, 4 * w, h, cudaMemcpyDeviceToHost, stream));
    cuda(StreamSynchronize(stream));
    cudaError_t e = cudaGetLastError();
    cuda(StreamDestroy(stream));
    cuda(FreeArray(arr));
    if (darr) cuda(FreeArray(darr));
    return e == cudaSuccess ? 0 : -(int)e;
}

int ref_render_f32(void* tp, int w, int h, float fx, float fy, const float* c2w12, const ref_options* o,
                   float* out_host) {
    N3Tree& tree = *static_cast<N3Tree*>(tp);
    Camera cam(w, h, fx, fy);
    set_cam(cam, c2w12);
    RenderOptions opt = to_opts(o);
    float4* d_out = nullptr;
    cuda(Malloc((void**)&d_out, sizeof(float4) * w * h));
    const int N_CUDA_THREADS = 320;
    const int blocks = N_BLOCKS_NEEDED(w * h, N_CUDA_THREADS);
    device::tap_kernel<<<blocks, N_CUDA_THREADS>>>(cam, tree, opt, d_out);
    cuda(Memcpy(out_host, d_out, sizeof(float4) * w * h, cudaMemcpyDeviceToHost));
    cudaError_t e = cudaGetLastError();
    cuda(Free(d_out));
    return e == cudaSuccess ? 0 : -(int)e;
}

// Float tap of the non-offscreen branch (existing colour + per-pixel depth limit, volrend.cu:92-96,143-163).
int ref_render_f32_composite(void* tp, int w, int h, float fx, float fy, const float* c2w12, const ref_options* o,
                             const uint8_t* rgba_in, const float* depth_in, float* out_host) {
    N3Tree& tree = *static_cast<N3Tree*>(tp);
    Camera cam(w, h, fx, fy);
    set_cam(cam, c2w12);
    RenderOptions opt = to_opts(o);
    float4* d_out = nullptr;
    uint8_t* d_rgba = nullptr;
    float* d_depth = nullptr;
    cuda(Malloc((void**)&d_out, sizeof(float4) * w * h));
    cuda(Malloc((void**)&d_rgba, (size_t)4 * w * h));
    cuda(Malloc((void**)&d_depth, sizeof(float) * w * h));
    cuda(Memcpy(d_rgba, rgba_in, (size_t)4 * w * h, cudaMemcpyHostToDevice));
    cuda(Memcpy(d_depth, depth_in, sizeof(float) * w * h, cudaMemcpyHostToDevice));
    const int N_CUDA_THREADS = 320;
    const int blocks = N_BLOCKS_NEEDED(w * h, N_CUDA_THREADS);
    device::tap_kernel<<<blocks, N_CUDA_THREADS>>>(cam, tree, opt, d_out, d_rgba, d_depth);
    cuda(Memcpy(out_host, d_out, sizeof(float4) * w * h, cudaMemcpyDeviceToHost));
    cudaError_t e = cudaGetLastError();
    cuda(Free(d_out)); cuda(Free(d_rgba)); cuda(Free(d_depth));
    return e == cudaSuccess ? 0 : -(int)e;
}

// main_headless.cpp:203-228: events on the legacy stream around the whole pose loop, one
// launch_renderer per pose on `stream`; with_d2h adds the -o read-back (:216-219) into host_out
// (n frames).  Returns total milliseconds for the n poses (negative on error).
float ref_time_frames(void* tp, int w, int h, float fx, float fy, const float* c2w12s, int n, const ref_options* o,
                      int with_d2h, uint8_t* host_out) {
    N3Tree& tree = *static_cast<N3Tree*>(tp);
    Camera camera(w, h, fx, fy);
    cudaArray_t array;
    cudaStream_t stream;
    cudaChannelFormatDesc channelDesc = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    cuda(MallocArray(&array, &channelDesc, w, h, cudaArraySurfaceLoadStore));
    cuda(StreamCreateWithFlags(&stream, cudaStreamDefault));
    cudaArray_t depth_arr = nullptr;
    RenderOptions options = to_opts(o);
    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    cudaEventRecord(start);
    for (int i = 0; i < n; ++i) {
        set_cam(camera, c2w12s + 12 * i);
        launch_renderer(tree, camera, options, array, depth_arr, stream, true);
        if (with_d2h) {
            cuda(Memcpy2DFromArrayAsync(host_out + (size_t)i * 4 * w * h, 4 * w, array, 0, 0, 4 * w, h,
                                        cudaMemcpyDeviceToHost, stream));
        }
    }
    cudaEventRecord(stop);
    cudaEventSynchronize(stop);
    float ms = 0;
    cudaEventElapsedTime(&ms, start, stop);
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    cuda(FreeArray(array));
    cuda(StreamDestroy(stream));
    return e == cudaSuccess ? ms : -1.f;
}

}  // extern "C"
