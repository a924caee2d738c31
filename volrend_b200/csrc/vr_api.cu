// vr_api.cu -- C-ABI of the B200 PlenOctree ray-marcher (declared in include/volrend_b200.h).
//
// Owns the device layout of a tree (the job of N3Tree::load_cuda, reference
// src/cuda/n3tree.cu:9-41) and the launch logic (launch_renderer, src/cuda/volrend.cu:195-245).
// There is no CPU path in this library: every entry point needs a CUDA device.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "volrend_b200.h"
#include "vr_kernels.h"
#include "vr_types.h"

using namespace vrb;

// ------------------------------------------------------------------------------------ errors
namespace {
thread_local std::string g_err;
std::atomic<int> g_variant{0};
std::atomic<unsigned long long> g_launches{0};
constexpr int kDefaultVariant = 3 + 16 * 193;  // persistent warps (kind 3) + cache hints (1) + wide tables (64) + wide-indexed records (128)
constexpr int kQueueSlots = 256;
constexpr int kCamRing = 8192;  // device ring of per-view cameras for batched launches

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define VR_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? VR_ENOMEM : VR_ECUDA, "%s: %s (%s:%d)", \
                        #expr, cudaGetErrorString(e__), __FILE__, __LINE__);                       \
    } while (0)

int kernel_basis(int format, int basis_dim) {
    if (format == VR_FMT_RGBA || basis_dim < 0) return -1;
    switch (basis_dim) {
        case 1: case 4: case 9: case 16: case 25: return basis_dim;
        default: return 1;  // rt_core.cuh:134-160: no switch case matches -> only coefficient 0
    }
}
int rec_bytes_for(int kbd) { return kbd <= 1 ? 8 : ((3 * kbd * 2 + 15) / 16) * 16; }
}  // namespace

struct vr_tree {
    int device = 0;
    int num_sms = 0;
    TreeDev dev{};
    uint32_t* nodes = nullptr;
    unsigned char* recs = nullptr;
    uint32_t* top = nullptr;
    uint32_t* wide = nullptr;
    uint32_t* wslot = nullptr;
    unsigned char* wrecs = nullptr;
    long long n_tables = 0;
    float* extra = nullptr;
    unsigned int* queues = nullptr;  // kQueueSlots x {head, done}
    CamDev* cam_ring = nullptr;      // kCamRing entries; batches take consecutive slots
    std::mutex cam_mu;
    unsigned int cam_pos = 0;
    // vr_render_frames_host: chunk ring, streams and events are created once and kept
    struct HostPath {
        static constexpr int kRing = 4;
        uint8_t* buf[kRing] = {};
        size_t buf_bytes = 0;
        cudaEvent_t rendered[kRing] = {}, copied[kRing] = {};
        cudaStream_t sr[2] = {nullptr, nullptr}, sc = nullptr;
        bool ready = false;
    } host;
    std::mutex host_mu;
    std::atomic<unsigned int> next_queue{0};
    vr_tree_info info{};
    int data_dim = 0;
    size_t l2_window_bytes = 0;  // node-table window kept in persisting L2 (VR_L2_PERSIST=1)
};

// ------------------------------------------------------------------------------------ re-layout kernels
namespace {

// nodes[i]: absolute child id, or leaf bit | sigma.  Also validates child links.
__global__ void relayout_nodes_kernel(const int32_t* __restrict__ child, const unsigned short* __restrict__ data,
                                      uint32_t* __restrict__ nodes, long long n_slots, long long capacity,
                                      int data_dim, int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const int32_t rel = child[i];
    if (rel != 0) {
        const long long tgt = (i >> 3) + rel;
        if (tgt <= 0 || tgt >= capacity) {
            atomicExch(bad, 1);
            nodes[i] = kLeafBit;
        } else {
            nodes[i] = (uint32_t)tgt;
        }
    } else {
        nodes[i] = kLeafBit | (uint32_t)data[(size_t)i * data_dim + (data_dim - 1)];
    }
}

// recs[i]: colour coefficients of slot i, padded.  One thread per (slot, 16-byte chunk).
__global__ void relayout_recs_kernel(const unsigned short* __restrict__ data, unsigned char* __restrict__ recs,
                                     long long n_slots, int data_dim, int basis_dim, int kbd, int rec_bytes) {
    const int chunks = rec_bytes >= 16 ? rec_bytes / 16 : 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long slot = gid / chunks;
    const int chunk = (int)(gid % chunks);
    if (slot >= n_slots) return;
    const unsigned short* src = data + (size_t)slot * data_dim;
    if (rec_bytes == 8) {
        // one coefficient per channel: RGBA -> halfs 0,1,2 ; basis -> k[0], k[bd], k[2bd]
        const int stride = kbd < 0 ? 1 : basis_dim;
        ushort4 v;
        v.x = src[0]; v.y = src[stride]; v.z = src[2 * stride]; v.w = 0;
        reinterpret_cast<ushort4*>(recs)[slot] = v;
        return;
    }
    const int n_coef = 3 * kbd;
    unsigned short h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = chunk * 8 + j;
        h[j] = c < n_coef ? src[c] : (unsigned short)0;
    }
    uint4 v;
    v.x = h[0] | ((uint32_t)h[1] << 16); v.y = h[2] | ((uint32_t)h[3] << 16);
    v.z = h[4] | ((uint32_t)h[5] << 16); v.w = h[6] | ((uint32_t)h[7] << 16);
    reinterpret_cast<uint4*>(recs + (size_t)slot * rec_bytes)[chunk] = v;
}

// Quantised trees (scripts/compress_octree.py; CPU decode in src/n3tree.cpp:309-340):
//   data[slot][j + n_retain + k*n_total] = quant_colors[j][quant_map[j][slot]][k]   j < n_quant
//   data[slot][j + k*n_total]            = data_retained[j][slot][k]                j < n_retain
//   data[slot][data_dim-1]               = sigma[slot]
// decoded here straight into the padded records / node words.
__global__ void quant_nodes_kernel(const int32_t* __restrict__ child, const unsigned short* __restrict__ sigma,
                                   uint32_t* __restrict__ nodes, long long n_slots, long long capacity,
                                   int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const int32_t rel = child[i];
    if (rel != 0) {
        const long long tgt = (i >> 3) + rel;
        if (tgt <= 0 || tgt >= capacity) {
            atomicExch(bad, 1);
            nodes[i] = kLeafBit;
        } else {
            nodes[i] = (uint32_t)tgt;
        }
    } else {
        nodes[i] = kLeafBit | (uint32_t)sigma[i];
    }
}

__device__ __forceinline__ unsigned short quant_coeff(const unsigned short* __restrict__ colors,
                                                      const unsigned short* __restrict__ qmap,
                                                      const unsigned short* __restrict__ retained, long long n_slots,
                                                      long long slot, int n_retain, int j, int k) {
    if (j < n_retain) return retained[((size_t)j * n_slots + slot) * 3 + k];
    const int q = j - n_retain;
    const unsigned int id = qmap[(size_t)q * n_slots + slot];
    return colors[((size_t)q * 65536 + id) * 3 + k];
}

__global__ void quant_recs_kernel(const unsigned short* __restrict__ colors, const unsigned short* __restrict__ qmap,
                                  const unsigned short* __restrict__ retained, unsigned char* __restrict__ recs,
                                  long long n_slots, int n_total, int n_retain, int kbd, int rec_bytes) {
    const int chunks = rec_bytes >= 16 ? rec_bytes / 16 : 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long slot = gid / chunks;
    const int chunk = (int)(gid % chunks);
    if (slot >= n_slots) return;
    if (rec_bytes == 8) {  // one coefficient per channel (basis sizes the reference's switch ignores)
        ushort4 v;
        v.x = quant_coeff(colors, qmap, retained, n_slots, slot, n_retain, 0, 0);
        v.y = quant_coeff(colors, qmap, retained, n_slots, slot, n_retain, 0, 1);
        v.z = quant_coeff(colors, qmap, retained, n_slots, slot, n_retain, 0, 2);
        v.w = 0;
        reinterpret_cast<ushort4*>(recs)[slot] = v;
        return;
    }
    unsigned short h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = chunk * 8 + i;                 // record half index = k*kbd + j (kbd == n_total here)
        h[i] = c < 3 * kbd ? quant_coeff(colors, qmap, retained, n_slots, slot, n_retain, c % n_total, c / n_total)
                           : (unsigned short)0;
    }
    uint4 v;
    v.x = h[0] | ((uint32_t)h[1] << 16); v.y = h[2] | ((uint32_t)h[3] << 16);
    v.z = h[4] | ((uint32_t)h[5] << 16); v.w = h[6] | ((uint32_t)h[7] << 16);
    reinterpret_cast<uint4*>(recs + (size_t)slot * rec_bytes)[chunk] = v;
}

// depth[n] of every node by level-synchronous relaxation from the root.
__global__ void node_depth_kernel(const uint32_t* __restrict__ nodes, int* __restrict__ depth, long long capacity,
                                  int level, int* __restrict__ changed) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= capacity || depth[n] != level) return;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const uint32_t w = nodes[n * 8 + s];
        if (!(w & kLeafBit)) {
            depth[w] = level + 1;
            *changed = 1;
        }
    }
}

// top[cell]: leaf word (bit31 | depth<<28 | sigma) or the depth-4 node id.
__global__ void build_top_kernel(const uint32_t* __restrict__ nodes, uint32_t* __restrict__ top) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= kTopCells) return;
    const uint32_t cx = (cell >> 8) & 15, cy = (cell >> 4) & 15, cz = cell & 15;
    uint32_t node = 0;
    for (int l = 1; l <= kTopLevel; ++l) {
        const int sh = kTopLevel - l;
        const uint32_t oct = (((cx >> sh) & 1u) << 2) | (((cy >> sh) & 1u) << 1) | ((cz >> sh) & 1u);
        const uint32_t w = nodes[node * 8u + oct];
        if (w & kLeafBit) {
            top[cell] = kLeafBit | ((uint32_t)l << 28) | (w & 0xffffu);
            return;
        }
        node = w;
    }
    top[cell] = node;
}

// wide[table*64 + e]: e = (ex<<4)|(ey<<2)|ez, two octree levels per axis (high bit first level).
__global__ void build_wide_kernel(const uint32_t* __restrict__ nodes, const int* __restrict__ depth,
                                  const uint32_t* __restrict__ tid, uint32_t* __restrict__ wide,
                                  uint32_t* __restrict__ wslot, long long capacity) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = g >> 6;
    const int e = (int)(g & 63);
    if (n >= capacity) return;
    const int d = depth[n];
    if (d < 0 || (d & 1)) return;            // unreachable node, or odd depth: folded into its parent's table
    // leaf entry = kLeafBit | (103 + leaf depth) << 23 | sigma: the exponent field of the cube size (vr_march.cuh)
    const uint32_t ex = (e >> 4) & 3, ey = (e >> 2) & 3, ez = e & 3;
    const uint32_t oct1 = ((ex >> 1) << 2) | ((ey >> 1) << 1) | (ez >> 1);
    const size_t o = (size_t)tid[n] * 64 + e;
    const uint32_t s1 = (uint32_t)n * 8u + oct1;
    const uint32_t w1 = nodes[s1];
    if (w1 & kLeafBit) {
        wide[o] = kLeafBit | ((uint32_t)(103 + d + 1) << 23) | (w1 & 0xffffu);
        wslot[o] = s1;
        return;
    }
    const uint32_t oct2 = ((ex & 1) << 2) | ((ey & 1) << 1) | (ez & 1);
    const uint32_t s2 = w1 * 8u + oct2;
    const uint32_t w2 = nodes[s2];
    if (w2 & kLeafBit) {
        wide[o] = kLeafBit | ((uint32_t)(103 + d + 2) << 23) | (w2 & 0xffffu);
        wslot[o] = s2;
    } else {
        wide[o] = tid[w2];
        wslot[o] = 0;
    }
}

// wrecs[entry] = recs[wslot[entry]] for leaf entries (16-byte chunks; 8-byte records as one chunk)
__global__ void build_wrecs_kernel(const uint32_t* __restrict__ wide, const uint32_t* __restrict__ wslot,
                                   const unsigned char* __restrict__ recs, unsigned char* __restrict__ wrecs,
                                   long long n_entries, int rec_bytes) {
    const int chunks = rec_bytes >= 16 ? rec_bytes / 16 : 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gid / chunks;
    const int c = (int)(gid % chunks);
    if (e >= n_entries) return;
    const bool leaf = (wide[e] & kLeafBit) != 0;
    if (rec_bytes == 8) {
        reinterpret_cast<uint2*>(wrecs)[e] = leaf ? reinterpret_cast<const uint2*>(recs)[wslot[e]] : make_uint2(0u, 0u);
        return;
    }
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (leaf) v = reinterpret_cast<const uint4*>(recs + (size_t)wslot[e] * rec_bytes)[c];
    reinterpret_cast<uint4*>(wrecs + (size_t)e * rec_bytes)[c] = v;
}

__global__ void probe_kernel(TreeDev tree, float x, float y, float z, int n_out, float* __restrict__ out) {
    // retrieve_cursor_lumisphere_kernel (volrend.cu:175-191)
    float p[3] = {tree.offset[0] + tree.scale[0] * x, tree.offset[1] + tree.scale[1] * y,
                  tree.offset[2] + tree.scale[2] * z};
    uint32_t u[3];
    for (int i = 0; i < 3; ++i) {
        p[i] = fmaxf(fminf(p[i], 1.f - 1e-6f), 0.f);
        u[i] = __float2uint_rz(p[i] * 16777216.f);
    }
    uint32_t node = 0, idx = 0;
    for (int l = 0;; ++l) {
        const int sh = 23 - l;
        const uint32_t oct = (((u[0] >> sh) & 1u) << 2) | (((u[1] >> sh) & 1u) << 1) | ((u[2] >> sh) & 1u);
        idx = node * 8u + oct;
        const uint32_t w = tree.nodes[idx];
        if (w & kLeafBit) break;
        node = w;
    }
    const unsigned short* rec = reinterpret_cast<const unsigned short*>(tree.recs + (size_t)idx * tree.rec_bytes);
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = __half2float(__ushort_as_half(rec[i]));
}

}  // namespace

// ------------------------------------------------------------------------------------ C-ABI
extern "C" {

const char* vr_last_error(void) { return g_err.c_str(); }
const char* vr_version(void) { return "volrend_b200 0.1 (sm_100a)"; }
int vr_set_variant(int variant) {
    if (variant < 0 || (variant & 15) > 6 || variant > 65535) return fail(VR_EINVAL, "variant must be kind 0..6 (+16*tune)");
    g_variant.store(variant);
    return VR_OK;
}
int vr_get_variant(void) {
    const int v = g_variant.load();
    return v == 0 ? kDefaultVariant : v;
}
unsigned long long vr_launch_count(void) { return g_launches.load(); }

void vr_default_options(vr_options* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->step_size = 1e-4f; o->sigma_thresh = 1e-2f; o->stop_thresh = 1e-2f;  // render_options.hpp:14-23
    o->background_brightness = 1.f;
    o->render_bbox[3] = o->render_bbox[4] = o->render_bbox[5] = 1.f;
    o->basis_minmax[0] = 0; o->basis_minmax[1] = VR_BASIS_MAX - 1;
}

void vr_tree_destroy(vr_tree* t) {
    if (!t) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(t->device);
    cudaFree(t->nodes); cudaFree(t->recs); cudaFree(t->top); cudaFree(t->extra); cudaFree(t->queues);
    cudaFree(t->wide); cudaFree(t->wslot); cudaFree(t->wrecs);
    cudaFree(t->cam_ring);
    for (int i = 0; i < vr_tree::HostPath::kRing; ++i) {
        cudaFree(t->host.buf[i]);
        if (t->host.rendered[i]) cudaEventDestroy(t->host.rendered[i]);
        if (t->host.copied[i]) cudaEventDestroy(t->host.copied[i]);
    }
    for (auto st : t->host.sr) if (st) cudaStreamDestroy(st);
    if (t->host.sc) cudaStreamDestroy(t->host.sc);
    cudaSetDevice(prev);
    delete t;
}

static int tree_create_impl(const vr_tree_desc* d, const vr_tree_quant_desc* q, vr_tree** out) {
    if (!d || !out) return fail(VR_EINVAL, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VR_ENODEVICE, "no CUDA device: volrend_b200 has no CPU fallback");
    }
    if (d->N != 2) return fail(VR_EUNSUPPORTED, "N=%d: only N=2 octrees are supported (as in the reference)", d->N);
    if (d->capacity < 1 || d->capacity >= (1ll << 28)) return fail(VR_EINVAL, "capacity %lld out of range", (long long)d->capacity);
    if (!d->child || (!q && !d->data)) return fail(VR_EINVAL, "child/data arrays missing");
    if (q) {
        if (!q->quant_colors || !q->quant_map || !q->sigma) return fail(VR_EINVAL, "quantised arrays missing");
        if (q->n_quant < 1 || q->n_retain < 0 || (q->n_retain > 0 && !q->data_retained))
            return fail(VR_EINVAL, "bad quantised basis counts");
        if (d->format == VR_FMT_RGBA || q->n_quant + q->n_retain != d->basis_dim)
            return fail(VR_EINVAL, "codebook and map basis numbers does not match");   // n3tree.cpp:296-299
    }
    if (d->format < VR_FMT_RGBA || d->format > VR_FMT_ASG) return fail(VR_EINVAL, "bad data format %d", d->format);
    const int kbd = kernel_basis(d->format, d->basis_dim);
    if (kbd < 0 ? d->data_dim < 4 : d->data_dim < 3 * d->basis_dim + 1)
        return fail(VR_EINVAL, "data_dim %d too small for format %d basis %d", d->data_dim, d->format, d->basis_dim);
    if ((d->format == VR_FMT_SG || d->format == VR_FMT_ASG) && !d->extra)
        return fail(VR_EINVAL, "SG/ASG trees need extra_data");
    if ((d->format == VR_FMT_SG || d->format == VR_FMT_ASG) && d->basis_dim > VR_BASIS_MAX)
        return fail(VR_EINVAL, "basis_dim %d > %d", d->basis_dim, VR_BASIS_MAX);

    vr_tree* t = new vr_tree();
    struct Guard { vr_tree*& t; bool ok = false; ~Guard() { if (!ok) { vr_tree_destroy(t); t = nullptr; } } } guard{t};
    VR_CUDA(cudaGetDevice(&t->device));
    cudaDeviceProp prop;
    VR_CUDA(cudaGetDeviceProperties(&prop, t->device));
    t->num_sms = prop.multiProcessorCount;
    t->data_dim = d->data_dim;
    const long long n_slots = d->capacity * 8;
    const int rec_bytes = rec_bytes_for(kbd);

    int32_t* raw_child = nullptr;
    unsigned short* raw_data = nullptr;
    int* flags = nullptr;
    int* depth = nullptr;
    struct Tmp { int32_t*& a; unsigned short*& b; int*& c; int*& d; ~Tmp() { cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d); } } tmp{raw_child, raw_data, flags, depth};
    const size_t child_bytes = (size_t)n_slots * 4, data_bytes = q ? 0 : (size_t)n_slots * d->data_dim * 2;
    VR_CUDA(cudaMalloc(&raw_child, child_bytes));
    if (!q) VR_CUDA(cudaMalloc(&raw_data, data_bytes));
    VR_CUDA(cudaMalloc(&flags, 2 * sizeof(int)));
    VR_CUDA(cudaMalloc(&depth, (size_t)d->capacity * sizeof(int)));
    VR_CUDA(cudaMalloc(&t->nodes, (size_t)n_slots * 4));
    VR_CUDA(cudaMalloc(&t->recs, (size_t)n_slots * rec_bytes));
    VR_CUDA(cudaMalloc(&t->top, kTopCells * 4));
    VR_CUDA(cudaMalloc(&t->queues, kQueueSlots * 2 * sizeof(unsigned int)));
    VR_CUDA(cudaMemset(t->queues, 0, kQueueSlots * 2 * sizeof(unsigned int)));
    VR_CUDA(cudaMalloc(&t->cam_ring, kCamRing * sizeof(CamDev)));
    VR_CUDA(cudaMemcpy(raw_child, d->child, child_bytes, cudaMemcpyHostToDevice));
    if (!q) VR_CUDA(cudaMemcpy(raw_data, d->data, data_bytes, cudaMemcpyHostToDevice));
    VR_CUDA(cudaMemset(flags, 0, 2 * sizeof(int)));
    if (d->extra && (d->format == VR_FMT_SG || d->format == VR_FMT_ASG)) {
        const size_t nf = (size_t)d->basis_dim * (d->format == VR_FMT_SG ? 4 : 11);
        VR_CUDA(cudaMalloc(&t->extra, nf * sizeof(float)));
        VR_CUDA(cudaMemcpy(t->extra, d->extra, nf * sizeof(float), cudaMemcpyHostToDevice));
    }
    const int TB = 256;
    const long long rec_work = n_slots * (rec_bytes >= 16 ? rec_bytes / 16 : 1);
    unsigned short *q_colors = nullptr, *q_map = nullptr, *q_sigma = nullptr, *q_ret = nullptr;
    struct QTmp { unsigned short *&a, *&b, *&c, *&d; ~QTmp() { cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d); } } qtmp{q_colors, q_map, q_sigma, q_ret};
    if (!q) {
        relayout_nodes_kernel<<<(unsigned)((n_slots + TB - 1) / TB), TB>>>(raw_child, raw_data, t->nodes, n_slots,
                                                                           d->capacity, d->data_dim, flags);
        relayout_recs_kernel<<<(unsigned)((rec_work + TB - 1) / TB), TB>>>(raw_data, t->recs, n_slots, d->data_dim,
                                                                           d->basis_dim, kbd, rec_bytes);
    } else {
        const size_t cb = (size_t)q->n_quant * 65536 * 3 * 2, mb = (size_t)q->n_quant * n_slots * 2,
                     sb = (size_t)n_slots * 2, rb = (size_t)q->n_retain * n_slots * 3 * 2;
        VR_CUDA(cudaMalloc(&q_colors, cb));
        VR_CUDA(cudaMalloc(&q_map, mb));
        VR_CUDA(cudaMalloc(&q_sigma, sb));
        VR_CUDA(cudaMemcpy(q_colors, q->quant_colors, cb, cudaMemcpyHostToDevice));
        VR_CUDA(cudaMemcpy(q_map, q->quant_map, mb, cudaMemcpyHostToDevice));
        VR_CUDA(cudaMemcpy(q_sigma, q->sigma, sb, cudaMemcpyHostToDevice));
        if (q->n_retain > 0) {
            VR_CUDA(cudaMalloc(&q_ret, rb));
            VR_CUDA(cudaMemcpy(q_ret, q->data_retained, rb, cudaMemcpyHostToDevice));
        }
        quant_nodes_kernel<<<(unsigned)((n_slots + TB - 1) / TB), TB>>>(raw_child, q_sigma, t->nodes, n_slots,
                                                                        d->capacity, flags);
        quant_recs_kernel<<<(unsigned)((rec_work + TB - 1) / TB), TB>>>(q_colors, q_map, q_ret, t->recs, n_slots,
                                                                        d->basis_dim, q->n_retain, kbd, rec_bytes);
    }
    VR_CUDA(cudaGetLastError());
    int h_flags[2] = {0, 0};
    VR_CUDA(cudaMemcpy(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost));
    if (h_flags[0]) return fail(VR_EINVAL, "child array has links outside [1, capacity)");
    // node depths -> max leaf depth
    VR_CUDA(cudaMemset(depth, 0xff, (size_t)d->capacity * sizeof(int)));
    VR_CUDA(cudaMemset(depth, 0, sizeof(int)));
    int max_node_depth = 0;
    for (int level = 0;; ++level) {
        if (level >= kMaxTreeDepth) return fail(VR_EUNSUPPORTED, "tree deeper than %d levels", kMaxTreeDepth);
        VR_CUDA(cudaMemset(flags + 1, 0, sizeof(int)));
        node_depth_kernel<<<(unsigned)((d->capacity + TB - 1) / TB), TB>>>(t->nodes, depth, d->capacity, level, flags + 1);
        int changed = 0;
        VR_CUDA(cudaMemcpy(&changed, flags + 1, sizeof(int), cudaMemcpyDeviceToHost));
        if (!changed) break;
        max_node_depth = level + 1;
    }
    build_top_kernel<<<(kTopCells + TB - 1) / TB, TB>>>(t->nodes, t->top);
    VR_CUDA(cudaGetLastError());
    {   // two-levels-per-step tables: ids = running count of even-depth internal nodes, in node order
        std::vector<int> h_depth((size_t)d->capacity);
        VR_CUDA(cudaMemcpy(h_depth.data(), depth, (size_t)d->capacity * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<uint32_t> h_tid((size_t)d->capacity, 0u);
        uint32_t n_tab = 0;
        for (long long n = 0; n < d->capacity; ++n)
            if (h_depth[(size_t)n] >= 0 && !(h_depth[(size_t)n] & 1)) h_tid[(size_t)n] = n_tab++;
        if (n_tab >= (1u << 30)) return fail(VR_EUNSUPPORTED, "too many nodes for the wide tables");
        uint32_t* d_tid = nullptr;
        struct T2 { uint32_t*& p; ~T2() { cudaFree(p); } } t2{d_tid};
        VR_CUDA(cudaMalloc(&d_tid, (size_t)d->capacity * 4));
        VR_CUDA(cudaMemcpy(d_tid, h_tid.data(), (size_t)d->capacity * 4, cudaMemcpyHostToDevice));
        VR_CUDA(cudaMalloc(&t->wide, (size_t)n_tab * 64 * 4));
        VR_CUDA(cudaMalloc(&t->wslot, (size_t)n_tab * 64 * 4));
        const long long work = d->capacity * 64;
        build_wide_kernel<<<(unsigned)((work + TB - 1) / TB), TB>>>(t->nodes, depth, d_tid, t->wide, t->wslot, d->capacity);
        VR_CUDA(cudaGetLastError());
        const long long n_entries = (long long)n_tab * 64;
        VR_CUDA(cudaMalloc(&t->wrecs, (size_t)n_entries * rec_bytes));
        const long long rwork = n_entries * (rec_bytes >= 16 ? rec_bytes / 16 : 1);
        build_wrecs_kernel<<<(unsigned)((rwork + TB - 1) / TB), TB>>>(t->wide, t->wslot, t->recs, t->wrecs, n_entries, rec_bytes);
        VR_CUDA(cudaGetLastError());
        VR_CUDA(cudaDeviceSynchronize());
        t->n_tables = n_tab;
    }
    VR_CUDA(cudaDeviceSynchronize());
    g_launches += 5;

    TreeDev& D = t->dev;
    D.nodes = t->nodes; D.recs = t->recs; D.top = t->top; D.extra = t->extra;
    D.wide = t->wide; D.wslot = t->wslot; D.wrecs = t->wrecs;
    for (int i = 0; i < 3; ++i) { D.offset[i] = d->offset[i]; D.scale[i] = d->scale[i]; }
    D.ndc_width = d->use_ndc ? d->ndc_width : -1.f;  // data_spec.hpp:47
    D.ndc_height = d->ndc_height; D.ndc_focal = d->ndc_focal;
    D.N = d->N; D.format = d->format; D.basis_dim = d->basis_dim; D.kbd = kbd;
    D.rec_bytes = rec_bytes; D.max_depth = max_node_depth + 1;
    t->info.capacity = d->capacity; t->info.max_depth = D.max_depth; t->info.rec_bytes = rec_bytes;
    t->info.node_bytes = n_slots * 4; t->info.rec_total_bytes = n_slots * (long long)rec_bytes;
    t->info.top_bytes = kTopCells * 4;
    if (const char* e = getenv("VR_L2_PERSIST")) {
        if (atoi(e) > 0) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, t->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, t->device);
            size_t want = (size_t)n_slots * 4;
            if (want > (size_t)max_window) want = (size_t)max_window;
            size_t carve = want < (size_t)max_persist ? want : (size_t)max_persist;
            if (carve > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess)
                t->l2_window_bytes = want;
            cudaGetLastError();
        }
    }
    guard.ok = true;
    *out = t;
    return VR_OK;
}

int vr_tree_create(const vr_tree_desc* d, vr_tree** out) { return tree_create_impl(d, nullptr, out); }

int vr_tree_create_quantized(const vr_tree_quant_desc* q, vr_tree** out) {
    if (!q) return fail(VR_EINVAL, "null argument");
    if (q->base.data) return fail(VR_EINVAL, "quantised descriptor must not carry decoded data");
    return tree_create_impl(&q->base, q, out);
}

int vr_tree_get_info(const vr_tree* t, vr_tree_info* info) {
    if (!t || !info) return fail(VR_EINVAL, "null argument");
    *info = t->info;
    return VR_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ launch plumbing
namespace {

void fill_opt(OptDev& o, const vr_options* s) {
    o.step_size = s->step_size; o.sigma_thresh = s->sigma_thresh; o.stop_thresh = s->stop_thresh;
    o.background_brightness = s->background_brightness;
    for (int i = 0; i < 6; ++i) o.render_bbox[i] = s->render_bbox[i];
    o.basis_min = s->basis_minmax[0]; o.basis_max = s->basis_minmax[1];
    for (int i = 0; i < 3; ++i) o.rot_dirs[i] = s->rot_dirs[i];
    o.render_depth = s->render_depth;
}
void fill_cam(CamDev& c, const vr_camera* s) {
    c.width = s->width; c.height = s->height; c.fx = s->fx; c.fy = s->fy;
    memcpy(c.c2w, s->c2w, sizeof(c.c2w));
}

int dispatch(const vr_tree* t, LaunchDev& P, bool count, bool surface, cudaStream_t stream) {
    LaunchCfg cfg;
    cfg.variant = vr_get_variant();
    cfg.count = count; cfg.surface = surface; cfg.num_sms = t->num_sms; cfg.stream = stream;
    vr_tree* mt = const_cast<vr_tree*>(t);
    cfg.queue = mt->queues + 2 * (mt->next_queue.fetch_add(1) % kQueueSlots);
    cfg.l2_window = t->nodes;
    cfg.l2_window_bytes = t->l2_window_bytes;
    static const bool no_pdl = getenv("VR_NO_PDL") != nullptr && atoi(getenv("VR_NO_PDL")) > 0;
    cfg.pdl = !no_pdl;
    cudaError_t e;
    switch (t->dev.kbd) {
        case -1: e = launch_march<-1>(P, cfg); break;
        case 1: e = launch_march<1>(P, cfg); break;
        case 4: e = launch_march<4>(P, cfg); break;
        case 9: e = launch_march<9>(P, cfg); break;
        case 16: e = launch_march<16>(P, cfg); break;
        case 25: e = launch_march<25>(P, cfg); break;
        default: return fail(VR_EUNSUPPORTED, "unsupported kernel basis %d", t->dev.kbd);
    }
    if (e != cudaSuccess) return fail(VR_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    g_launches += 1;
    return VR_OK;
}

// Upper bound of the warp tiles of one view for every tile shape the kernels are built with
// (2x16, 4x8, 8x4 pixels): the persistent kernels index tiles of a whole batch with 31 bits.
long long tile_bound(const vr_rect& r) { return ((long long)r.w / 2 + 1) * ((long long)r.h / 4 + 1); }

int check_common(const vr_tree* t, const vr_camera* cam, const vr_options* opt, const vr_rect* tile, vr_rect& r) {
    if (!t || !cam || !opt) return fail(VR_EINVAL, "null argument");
    if (cam->width <= 0 || cam->height <= 0) return fail(VR_EINVAL, "bad camera size %dx%d", cam->width, cam->height);
    if (tile) {
        r = *tile;
        if (r.w < 0 || r.h < 0 || r.x0 < 0 || r.y0 < 0 || r.x0 + r.w > cam->width || r.y0 + r.h > cam->height)
            return fail(VR_EINVAL, "tile (%d,%d,%d,%d) outside %dx%d frame", r.x0, r.y0, r.w, r.h, cam->width, cam->height);
    } else {
        r.x0 = r.y0 = 0; r.w = cam->width; r.h = cam->height;
    }
    if (tile_bound(r) > 0x7fffffffLL)
        return fail(VR_EUNSUPPORTED, "%dx%d pixels: more tiles than the 31-bit work queue can index", r.w, r.h);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(VR_ENODEVICE, "no CUDA device");
    if (dev != t->device) return fail(VR_EINVAL, "tree lives on device %d but device %d is current", t->device, dev);
    return VR_OK;
}

}  // namespace

extern "C" {

int vr_render_batch(const vr_tree* t, const vr_camera* cams, int n_views, const vr_options* opt, const vr_rect* tile,
                    uint8_t* rgba8_dev, float* rgba32f_dev, vr_counters* counters_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_views < 0) return fail(VR_EINVAL, "n_views < 0");
    if (n_views == 0) return VR_OK;
    vr_rect r;
    int rc = check_common(t, cams, opt, tile, r);
    if (rc) return rc;
    for (int i = 1; i < n_views; ++i)
        if (cams[i].width != cams[0].width || cams[i].height != cams[0].height)
            return fail(VR_EINVAL, "all views of a batch must share one image size");
    if (r.w == 0 || r.h == 0) return VR_OK;
    LaunchDev P{};
    P.tree = t->dev;
    fill_opt(P.opt, opt);
    fill_cam(P.cam, &cams[0]);
    P.n_views = n_views;
    P.x0 = r.x0; P.y0 = r.y0; P.w = r.w; P.h = r.h;
    P.rgba8 = rgba8_dev; P.rgbaf = reinterpret_cast<float4*>(rgba32f_dev);
    P.counters = counters_dev;
    int max_views = kCamRing;  // split very large batches: camera ring size, 31-bit tile index
    if (tile_bound(r) * max_views > 0x7fffffffLL) max_views = (int)(0x7fffffffLL / tile_bound(r));
    if (n_views > max_views) {
        const size_t tile_px = (size_t)r.w * r.h;
        for (int v0 = 0; v0 < n_views; v0 += max_views) {
            const int nv = n_views - v0 < max_views ? n_views - v0 : max_views;
            rc = vr_render_batch(t, cams + v0, nv, opt, tile, rgba8_dev ? rgba8_dev + 4 * tile_px * v0 : nullptr,
                                 rgba32f_dev ? rgba32f_dev + 4 * tile_px * v0 : nullptr, counters_dev, stream_);
            if (rc) return rc;
        }
        return VR_OK;
    }
    if (n_views > 1) {
        // cameras go through a device ring owned by the tree (no allocation on the launch path: a
        // stream-ordered pool would hand memory back to the OS at every synchronisation)
        vr_tree* mt = const_cast<vr_tree*>(t);
        unsigned int slot;
        {
            std::lock_guard<std::mutex> lk(mt->cam_mu);
            if (mt->cam_pos + (unsigned int)n_views > (unsigned int)kCamRing) mt->cam_pos = 0;  // no wrap inside a batch
            slot = mt->cam_pos;
            mt->cam_pos += (unsigned int)n_views;
        }
        std::vector<CamDev> h(n_views);
        for (int i = 0; i < n_views; ++i) fill_cam(h[i], &cams[i]);
        CamDev* dcams = mt->cam_ring + slot;
        VR_CUDA(cudaMemcpyAsync(dcams, h.data(), sizeof(CamDev) * n_views, cudaMemcpyHostToDevice, stream));
        // pageable source: the copy has been staged when cudaMemcpyAsync returns
        P.cams = dcams;
    }
    return dispatch(t, P, counters_dev != nullptr, false, stream);
}

int vr_render_bands(const vr_tree* t, const vr_camera* cam, const vr_options* opt, int band_h, int n_parts,
                    int part, uint8_t* rgba8_dev, float* rgba32f_dev, void* stream_) {
    vr_rect r;
    int rc = check_common(t, cam, opt, nullptr, r);
    if (rc) return rc;
    if (band_h < 4 || band_h % 4 || n_parts < 1 || part < 0 || part >= n_parts)
        return fail(VR_EINVAL, "bands: band_h must be a positive multiple of 4 and 0 <= part < n_parts");
    const int rows = vr_band_rows(cam->height, band_h, n_parts, part);
    if (rows == 0) return VR_OK;
    LaunchDev P{};
    P.tree = t->dev;
    fill_opt(P.opt, opt);
    fill_cam(P.cam, cam);
    P.n_views = 1;
    P.x0 = 0; P.y0 = 0; P.w = r.w; P.h = rows;
    P.band_h = band_h; P.band_parts = n_parts; P.band_part = part;
    P.rgba8 = rgba8_dev; P.rgbaf = reinterpret_cast<float4*>(rgba32f_dev);
    return dispatch(t, P, false, false, (cudaStream_t)stream_);
}

int vr_band_rows(int height, int band_h, int n_parts, int part) {
    if (height <= 0 || band_h <= 0 || n_parts <= 0 || part < 0 || part >= n_parts) return 0;
    const int n_bands = (height + band_h - 1) / band_h;
    int rows = 0;
    for (int b = part; b < n_bands; b += n_parts) {
        const int y = b * band_h;
        rows += (height - y < band_h) ? height - y : band_h;
    }
    return rows;
}

int vr_debug_trace(const vr_tree* t, const vr_camera* cam, const vr_options* opt, uint8_t* rgba8_dev,
                   vr_counters* counters_dev, unsigned long long* trace_dev, void* stream_) {
    vr_rect r;
    int rc = check_common(t, cam, opt, nullptr, r);
    if (rc) return rc;
    if (!counters_dev || !trace_dev) return fail(VR_EINVAL, "trace needs counters and a trace buffer");
    LaunchDev P{};
    P.tree = t->dev;
    fill_opt(P.opt, opt);
    fill_cam(P.cam, cam);
    P.n_views = 1;
    P.x0 = 0; P.y0 = 0; P.w = r.w; P.h = r.h;
    P.rgba8 = rgba8_dev; P.counters = counters_dev; P.trace = trace_dev;
    return dispatch(t, P, true, false, (cudaStream_t)stream_);
}

int vr_render(const vr_tree* t, const vr_camera* cam, const vr_options* opt, const vr_rect* tile, uint8_t* rgba8_dev,
              float* rgba32f_dev, vr_counters* counters_dev, void* stream) {
    return vr_render_batch(t, cam, 1, opt, tile, rgba8_dev, rgba32f_dev, counters_dev, stream);
}

int vr_render_composite(const vr_tree* t, const vr_camera* cam, const vr_options* opt, const vr_rect* tile,
                        uint8_t* rgba8_dev, const float* depth_dev, float* rgba32f_dev, void* stream_) {
    vr_rect r;
    int rc = check_common(t, cam, opt, tile, r);
    if (rc) return rc;
    if (!rgba8_dev || !depth_dev) return fail(VR_EINVAL, "composite mode needs colour and depth inputs");
    if (r.w == 0 || r.h == 0) return VR_OK;
    LaunchDev P{};
    P.tree = t->dev;
    fill_opt(P.opt, opt);
    fill_cam(P.cam, cam);
    P.n_views = 1;
    P.x0 = r.x0; P.y0 = r.y0; P.w = r.w; P.h = r.h;
    P.rgba8 = rgba8_dev; P.rgbaf = reinterpret_cast<float4*>(rgba32f_dev);
    P.depth_in = depth_dev; P.composite = 1;
    return dispatch(t, P, false, false, (cudaStream_t)stream_);
}

int vr_render_surface(const vr_tree* t, const vr_camera* cam, const vr_options* opt, unsigned long long rgba8_surf,
                      unsigned long long depth_surf, void* stream_) {
    vr_rect r;
    int rc = check_common(t, cam, opt, nullptr, r);
    if (rc) return rc;
    if (!rgba8_surf) return fail(VR_EINVAL, "null surface");
    LaunchDev P{};
    P.tree = t->dev;
    fill_opt(P.opt, opt);
    fill_cam(P.cam, cam);
    P.n_views = 1;
    P.x0 = 0; P.y0 = 0; P.w = r.w; P.h = r.h;
    P.surf = (cudaSurfaceObject_t)rgba8_surf; P.dsurf = (cudaSurfaceObject_t)depth_surf;
    P.composite = depth_surf != 0;
    return dispatch(t, P, false, true, (cudaStream_t)stream_);
}

int vr_render_frames_host(const vr_tree* t, const vr_camera* cams, int n_views, const vr_options* opt,
                          uint8_t* rgba8_host) {
    // main_headless.cpp:208-223 with -o: every frame goes back to host memory.  Frames are
    // rendered in chunks of a few views per launch on two alternating streams (the tail of
    // one chunk overlaps the head of the next) and copied out on a third stream while the
    // following chunks render.
    if (n_views < 0) return fail(VR_EINVAL, "n_views < 0");
    if (n_views == 0) return VR_OK;
    if (!rgba8_host) return fail(VR_EINVAL, "null host buffer");
    vr_rect r;
    int rc = check_common(t, cams, opt, nullptr, r);
    if (rc) return rc;
    for (int i = 1; i < n_views; ++i)
        if (cams[i].width != cams[0].width || cams[i].height != cams[0].height)
            return fail(VR_EINVAL, "all views must share one image size");
    const size_t frame = (size_t)4 * r.w * r.h;
    int chunk = 8;
    if (const char* e = getenv("VR_HOST_CHUNK")) chunk = atoi(e) > 0 ? atoi(e) : chunk;
    if (chunk > n_views) chunk = n_views;
    vr_tree* mt = const_cast<vr_tree*>(t);
    std::lock_guard<std::mutex> host_lock(mt->host_mu);   // one host-path sweep per tree at a time
    vr_tree::HostPath& R = mt->host;
    constexpr int kRing = vr_tree::HostPath::kRing;
    if (!R.ready) {
        VR_CUDA(cudaStreamCreateWithFlags(&R.sr[0], cudaStreamNonBlocking));
        VR_CUDA(cudaStreamCreateWithFlags(&R.sr[1], cudaStreamNonBlocking));
        VR_CUDA(cudaStreamCreateWithFlags(&R.sc, cudaStreamNonBlocking));
        for (int i = 0; i < kRing; ++i) {
            VR_CUDA(cudaEventCreateWithFlags(&R.rendered[i], cudaEventDisableTiming));
            VR_CUDA(cudaEventCreateWithFlags(&R.copied[i], cudaEventDisableTiming));
        }
        R.ready = true;
    }
    if (R.buf_bytes < frame * chunk) {
        for (int i = 0; i < kRing; ++i) { cudaFree(R.buf[i]); R.buf[i] = nullptr; }
        R.buf_bytes = 0;
        for (int i = 0; i < kRing; ++i) VR_CUDA(cudaMalloc(&R.buf[i], frame * chunk));
        R.buf_bytes = frame * chunk;
    }
    int c = 0;
    for (int v0 = 0; v0 < n_views; v0 += chunk, ++c) {
        const int nv = n_views - v0 < chunk ? n_views - v0 : chunk;
        const int s = c % kRing;
        cudaStream_t sr = R.sr[c & 1];
        if (c >= kRing) VR_CUDA(cudaStreamWaitEvent(sr, R.copied[s], 0));
        rc = vr_render_batch(t, cams + v0, nv, opt, nullptr, R.buf[s], nullptr, nullptr, sr);
        if (rc) return rc;
        VR_CUDA(cudaEventRecord(R.rendered[s], sr));
        VR_CUDA(cudaStreamWaitEvent(R.sc, R.rendered[s], 0));
        VR_CUDA(cudaMemcpyAsync(rgba8_host + (size_t)v0 * frame, R.buf[s], frame * nv, cudaMemcpyDeviceToHost, R.sc));
        VR_CUDA(cudaEventRecord(R.copied[s], R.sc));
    }
    VR_CUDA(cudaStreamSynchronize(R.sc));
    VR_CUDA(cudaStreamSynchronize(R.sr[0]));
    VR_CUDA(cudaStreamSynchronize(R.sr[1]));
    return VR_OK;
}

int vr_probe_lumisphere(const vr_tree* t, const float xyz[3], float* out_dev, void* stream_) {
    if (!t || !xyz || !out_dev) return fail(VR_EINVAL, "null argument");
    const int n_out = t->data_dim - 1;
    if (t->dev.kbd > 0 && t->dev.kbd != t->dev.basis_dim)
        return fail(VR_EUNSUPPORTED, "probe unavailable for basis_dim %d (only coefficient 0 is resident)", t->dev.basis_dim);
    probe_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(t->dev, xyz[0], xyz[1], xyz[2], n_out, out_dev);
    VR_CUDA(cudaGetLastError());
    g_launches += 1;
    return VR_OK;
}

}  // extern "C"
