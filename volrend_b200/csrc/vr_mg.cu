// vr_mg.cu -- one process driving N GPUs of a box (declared in include/volrend_b200.h, "multi-GPU").
//
// The path shards embarrassingly: the tree is replicated, every device renders its share
// with ONE persistent-kernel launch per batch of views, and the only exchange is the gather of finished
// RGBA8 pixels on the first device.  That gather is done by the COPY ENGINES over NVLink (peer copies
// issued on a second stream of the producing device, 2-D copies scatter interleaved bands straight into
// their rows of the frame): no SM-resident collective kernel competes with the persistent march kernel.
//   VR_MG_VIEWS  view i is rendered by device i % n           (main_headless.cpp:208-223 sharded by pose)
//   VR_MG_TILES  every frame is cut into bands of band_h rows, band b goes to device b % n (ray-tile
//                sharding of one frame: strong scaling, BASELINE config 4)
// One host thread per device issues that device's work, so launches on different devices do not
// serialise behind each other.
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "volrend_b200.h"

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    std::string err;
};

struct Dev {
    int device = 0;
    bool is_first = false;      // same physical device as the gathering one (plain device-to-device copies)
    vr_tree* tree = nullptr;
    cudaStream_t s_render = nullptr, s_copy = nullptr;
    cudaEvent_t e_start = nullptr, e_end = nullptr, e_rendered[2] = {nullptr, nullptr}, e_copied[2] = {nullptr, nullptr};
    uint8_t* local[2] = {nullptr, nullptr};   // compact outputs of this device, double-buffered per batch
    size_t local_bytes = 0;
    Worker* w = nullptr;
};

}  // namespace

struct vr_mg {
    std::vector<Dev> devs;
    uint8_t* gather = nullptr;   // on devs[0].device, grown on demand
    size_t gather_bytes = 0;
    cudaEvent_t e_all = nullptr; // on devs[0]: every peer's copies have landed
    std::string err;
};

namespace {

thread_local std::string g_mg_err;

int mg_fail(vr_mg* mg, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_mg_err = buf;
    if (mg) mg->err = buf;
    return code;
}

#define MG_CUDA(mg, expr)                                                                                     \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return mg_fail(mg, e__ == cudaErrorMemoryAllocation ? VR_ENOMEM : VR_ECUDA, "%s: %s (%s:%d)", #expr, \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                                      \
    } while (0)

void worker_main(Worker* w) {
    for (;;) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job = std::move(w->job);
            w->has_job = false;
        }
        const int rc = job();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->rc = rc;
            if (rc) w->err = vr_last_error();
            w->done = true;
        }
        w->cv.notify_all();
    }
}

void submit(Worker* w, std::function<int()> job) {
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = std::move(job);
        w->has_job = true;
        w->done = false;
    }
    w->cv.notify_all();
}

int wait_done(Worker* w) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    return w->rc;
}

// Peer (or same-device) copy of `rows` rows of `width` bytes; pitches in bytes.
cudaError_t copy2d(uint8_t* dst, int dst_dev, size_t dpitch, const uint8_t* src, int src_dev, size_t spitch, size_t width,
                   size_t rows, cudaStream_t st) {
    if (rows == 0 || width == 0) return cudaSuccess;
    if (dst_dev == src_dev) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToDevice, st);
    cudaMemcpy3DPeerParms p;
    memset(&p, 0, sizeof(p));
    p.srcDevice = src_dev;
    p.dstDevice = dst_dev;
    p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(src), spitch, width, rows);
    p.dstPtr = make_cudaPitchedPtr(dst, dpitch, width, rows);
    p.extent = make_cudaExtent(width, rows, 1);
    return cudaMemcpy3DPeerAsync(&p, st);
}

}  // namespace

extern "C" {

const char* vr_mg_last_error(const vr_mg* mg) { return mg ? mg->err.c_str() : g_mg_err.c_str(); }

void vr_mg_destroy(vr_mg* mg) {
    if (!mg) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (Dev& d : mg->devs) {
        if (d.w) {
            {
                std::lock_guard<std::mutex> lk(d.w->mu);
                d.w->quit = true;
            }
            d.w->cv.notify_all();
            if (d.w->th.joinable()) d.w->th.join();
            delete d.w;
        }
        cudaSetDevice(d.device);
        cudaDeviceSynchronize();
        if (d.tree) vr_tree_destroy(d.tree);
        for (auto* p : d.local) cudaFree(p);
        if (d.s_render) cudaStreamDestroy(d.s_render);
        if (d.s_copy) cudaStreamDestroy(d.s_copy);
        for (cudaEvent_t e : {d.e_start, d.e_end, d.e_rendered[0], d.e_rendered[1], d.e_copied[0], d.e_copied[1]})
            if (e) cudaEventDestroy(e);
    }
    if (!mg->devs.empty()) {
        cudaSetDevice(mg->devs[0].device);
        cudaFree(mg->gather);
        if (mg->e_all) cudaEventDestroy(mg->e_all);
    }
    cudaSetDevice(prev);
    delete mg;
}

int vr_mg_create(const vr_tree_desc* desc, const int* devices, int n_devices, vr_mg** out) {
    if (!desc || !out || n_devices < 1 || n_devices > 64) return mg_fail(nullptr, VR_EINVAL, "vr_mg_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return mg_fail(nullptr, VR_ENODEVICE, "no CUDA device: volrend_b200 has no CPU fallback");
    }
    int prev = 0;
    cudaGetDevice(&prev);
    vr_mg* mg = new vr_mg();
    struct Guard { vr_mg*& mg; int prev; bool ok = false; ~Guard() { if (!ok) { vr_mg_destroy(mg); mg = nullptr; } cudaSetDevice(prev); } } guard{mg, prev};
    mg->devs.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        Dev& d = mg->devs[i];
        d.device = devices ? devices[i] : i;
        if (d.device < 0 || d.device >= ndev) return mg_fail(mg, VR_EINVAL, "device %d does not exist (%d visible)", d.device, ndev);
        d.is_first = d.device == (devices ? devices[0] : 0);
    }
    // trees are built one after the other (the pinned-staging uploader is process-wide)
    for (Dev& d : mg->devs) {
        MG_CUDA(mg, cudaSetDevice(d.device));
        const int rc = vr_tree_create(desc, &d.tree);
        if (rc) return mg_fail(mg, rc, "device %d: %s", d.device, vr_last_error());
        MG_CUDA(mg, cudaStreamCreateWithFlags(&d.s_render, cudaStreamNonBlocking));
        MG_CUDA(mg, cudaStreamCreateWithFlags(&d.s_copy, cudaStreamNonBlocking));
        MG_CUDA(mg, cudaEventCreate(&d.e_start));
        MG_CUDA(mg, cudaEventCreate(&d.e_end));
        for (int k = 0; k < 2; ++k) {
            MG_CUDA(mg, cudaEventCreateWithFlags(&d.e_rendered[k], cudaEventDisableTiming));
            MG_CUDA(mg, cudaEventCreateWithFlags(&d.e_copied[k], cudaEventDisableTiming));
        }
        if (!d.is_first) {   // direct NVLink stores into the gather buffer; without peer access the driver stages the copy
            int can = 0;
            cudaDeviceCanAccessPeer(&can, d.device, mg->devs[0].device);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(mg->devs[0].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                cudaGetLastError();
            }
        }
        d.w = new Worker();
        d.w->th = std::thread(worker_main, d.w);
    }
    MG_CUDA(mg, cudaSetDevice(mg->devs[0].device));
    MG_CUDA(mg, cudaEventCreateWithFlags(&mg->e_all, cudaEventDisableTiming));
    guard.ok = true;
    *out = mg;
    return VR_OK;
}

int vr_mg_device_count(const vr_mg* mg) { return mg ? (int)mg->devs.size() : 0; }

int vr_mg_render(vr_mg* mg, const vr_camera* cams, int n_views, const vr_options* opt, int mode, int band_h, int batch,
                 uint8_t* rgba8_dev0, uint8_t* rgba8_host, float* ms_out) {
    if (!mg || !cams || !opt) return mg_fail(mg, VR_EINVAL, "null argument");
    if (n_views < 0) return mg_fail(mg, VR_EINVAL, "n_views < 0");
    if (ms_out) *ms_out = 0.f;
    if (n_views == 0) return VR_OK;
    if (mode != VR_MG_VIEWS && mode != VR_MG_TILES) return mg_fail(mg, VR_EINVAL, "mode must be VR_MG_VIEWS or VR_MG_TILES");
    const int n = (int)mg->devs.size();
    const int W = cams[0].width, H = cams[0].height;
    if (W <= 0 || H <= 0) return mg_fail(mg, VR_EINVAL, "bad camera size %dx%d", W, H);
    for (int i = 1; i < n_views; ++i)
        if (cams[i].width != W || cams[i].height != H) return mg_fail(mg, VR_EINVAL, "all views must share one image size");
    if (mode == VR_MG_TILES && (band_h < 4 || band_h % 4)) return mg_fail(mg, VR_EINVAL, "band_h must be a positive multiple of 4");
    if (batch <= 0 || batch > n_views) batch = n_views;
    int prev = 0;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    const size_t row_bytes = (size_t)4 * W, frame = row_bytes * H;
    const int dev0 = mg->devs[0].device;

    // destination frames on the first device
    uint8_t* dst = rgba8_dev0;
    MG_CUDA(mg, cudaSetDevice(dev0));
    if (!dst) {
        if (mg->gather_bytes < frame * n_views) {
            cudaFree(mg->gather);
            mg->gather = nullptr; mg->gather_bytes = 0;
            MG_CUDA(mg, cudaMalloc(&mg->gather, frame * n_views));
            mg->gather_bytes = frame * n_views;
        }
        dst = mg->gather;
    }
    // per-device compact buffers: VIEWS ceil(batch/n) frames, TILES batch x (rows of its bands)
    for (int p = 0; p < n; ++p) {
        Dev& d = mg->devs[p];
        const size_t need = mode == VR_MG_VIEWS ? frame * (size_t)((batch + n - 1) / n)
                                                : row_bytes * (size_t)vr_band_rows(H, band_h, n, p) * batch;
        if (d.local_bytes < need) {
            MG_CUDA(mg, cudaSetDevice(d.device));
            for (auto*& b : d.local) { cudaFree(b); b = nullptr; }
            d.local_bytes = 0;
            for (auto*& b : d.local) MG_CUDA(mg, cudaMalloc(&b, need ? need : 1));
            d.local_bytes = need;
        }
    }
    for (Dev& d : mg->devs) {   // everything idle: the per-device clocks start together
        MG_CUDA(mg, cudaSetDevice(d.device));
        MG_CUDA(mg, cudaDeviceSynchronize());
    }

    for (int p = 0; p < n; ++p) {
        Dev* d = &mg->devs[p];
        submit(d->w, [=]() -> int {
            if (cudaSetDevice(d->device) != cudaSuccess) return VR_ECUDA;
            if (cudaEventRecord(d->e_start, d->s_render) != cudaSuccess) return VR_ECUDA;
            int b = 0;
            for (int v0 = 0; v0 < n_views; v0 += batch, ++b) {
                const int nv = n_views - v0 < batch ? n_views - v0 : batch;
                const int k = b & 1;
                if (b >= 2 && cudaStreamWaitEvent(d->s_render, d->e_copied[k], 0) != cudaSuccess) return VR_ECUDA;   // buffer reuse
                uint8_t* loc = d->local[k];
                if (mode == VR_MG_VIEWS) {
                    // views v0+p, v0+p+n, ... of this batch
                    std::vector<vr_camera> mine;
                    for (int i = p; i < nv; i += n) mine.push_back(cams[v0 + i]);
                    if (!mine.empty()) {
                        int rc = vr_render_batch(d->tree, mine.data(), (int)mine.size(), opt, nullptr, loc, nullptr, nullptr, d->s_render);
                        if (rc) return rc;
                    }
                    if (cudaEventRecord(d->e_rendered[k], d->s_render) != cudaSuccess) return VR_ECUDA;
                    if (cudaStreamWaitEvent(d->s_copy, d->e_rendered[k], 0) != cudaSuccess) return VR_ECUDA;
                    int j = 0;
                    for (int i = p; i < nv; i += n, ++j) {
                        uint8_t* to = dst + (size_t)(v0 + i) * frame;
                        cudaError_t e = d->device == dev0 ? cudaMemcpyAsync(to, loc + (size_t)j * frame, frame, cudaMemcpyDeviceToDevice, d->s_copy)
                                                          : cudaMemcpyPeerAsync(to, dev0, loc + (size_t)j * frame, d->device, frame, d->s_copy);
                        if (e != cudaSuccess) return VR_ECUDA;
                    }
                } else {
                    const int rows = vr_band_rows(H, band_h, n, p);
                    if (rows > 0) {
                        int rc = vr_render_bands_batch(d->tree, cams + v0, nv, opt, band_h, n, p, loc, nullptr, d->s_render);
                        if (rc) return rc;
                    }
                    if (cudaEventRecord(d->e_rendered[k], d->s_render) != cudaSuccess) return VR_ECUDA;
                    if (cudaStreamWaitEvent(d->s_copy, d->e_rendered[k], 0) != cudaSuccess) return VR_ECUDA;
                    if (rows > 0) {
                        const size_t band_bytes = row_bytes * band_h;
                        const int n_bands = (H + band_h - 1) / band_h;
                        const int owned = (n_bands - p + n - 1) / n;                      // bands p, p+n, ...
                        const bool ragged = (H % band_h) != 0 && ((n_bands - 1) % n) == p;  // the short last band is ours
                        const int full = ragged ? owned - 1 : owned;
                        for (int i = 0; i < nv; ++i) {
                            const uint8_t* from = loc + (size_t)i * rows * row_bytes;
                            uint8_t* to = dst + (size_t)(v0 + i) * frame + (size_t)p * band_bytes;
                            if (copy2d(to, dev0, band_bytes * n, from, d->device, band_bytes, band_bytes, (size_t)full, d->s_copy) != cudaSuccess)
                                return VR_ECUDA;
                            if (ragged) {
                                const size_t tail = (size_t)(H % band_h) * row_bytes;
                                if (copy2d(to + (size_t)full * band_bytes * n, dev0, tail, from + (size_t)full * band_bytes, d->device, tail,
                                           tail, 1, d->s_copy) != cudaSuccess)
                                    return VR_ECUDA;
                            }
                        }
                    }
                }
                if (cudaEventRecord(d->e_copied[k], d->s_copy) != cudaSuccess) return VR_ECUDA;
            }
            if (cudaStreamWaitEvent(d->s_render, d->e_copied[(b - 1) & 1], 0) != cudaSuccess) return VR_ECUDA;
            if (cudaEventRecord(d->e_end, d->s_render) != cudaSuccess) return VR_ECUDA;
            return VR_OK;
        });
    }
    int rc_all = VR_OK;
    std::string first_err;
    for (Dev& d : mg->devs) {
        const int rc = wait_done(d.w);
        if (rc && !rc_all) { rc_all = rc; first_err = d.w->err; }
    }
    if (rc_all) {
        for (Dev& d : mg->devs) { cudaSetDevice(d.device); cudaDeviceSynchronize(); }
        cudaGetLastError();
        return mg_fail(mg, rc_all, "vr_mg_render: %s", first_err.empty() ? "CUDA call failed on a worker" : first_err.c_str());
    }
    // the first device's stream sees every peer's frames; optional read-back to the host from there
    Dev& d0 = mg->devs[0];
    MG_CUDA(mg, cudaSetDevice(dev0));
    for (Dev& d : mg->devs) MG_CUDA(mg, cudaStreamWaitEvent(d0.s_render, d.e_end, 0));
    if (rgba8_host) MG_CUDA(mg, cudaMemcpyAsync(rgba8_host, dst, frame * n_views, cudaMemcpyDeviceToHost, d0.s_render));
    MG_CUDA(mg, cudaEventRecord(mg->e_all, d0.s_render));
    // e_end of device 0 is re-recorded behind the joins so that its interval covers the whole job
    MG_CUDA(mg, cudaEventRecord(d0.e_end, d0.s_render));
    float ms_max = 0.f;
    for (Dev& d : mg->devs) {
        MG_CUDA(mg, cudaSetDevice(d.device));
        MG_CUDA(mg, cudaEventSynchronize(d.e_end));
        float ms = 0.f;
        MG_CUDA(mg, cudaEventElapsedTime(&ms, d.e_start, d.e_end));
        if (ms > ms_max) ms_max = ms;
    }
    if (ms_out) *ms_out = ms_max;
    return VR_OK;
}

const uint8_t* vr_mg_frames_dev0(const vr_mg* mg) { return mg ? mg->gather : nullptr; }

vr_tree* vr_mg_tree(const vr_mg* mg, int index) {
    if (!mg || index < 0 || index >= (int)mg->devs.size()) return nullptr;
    return mg->devs[index].tree;
}

int vr_mg_device(const vr_mg* mg, int index) {
    if (!mg || index < 0 || index >= (int)mg->devs.size()) return -1;
    return mg->devs[index].device;
}

}  // extern "C"
