// vr_kernels_inst.cu -- instantiates the march kernels for one basis size (-DVR_KBD=...).
// Compiled six times (RGBA, SH/SG/ASG 1, 4, 9, 16, 25) so the builds run in parallel.
#include "vr_kernels.h"
#include "vr_march.cuh"

#ifndef VR_KBD
#error "compile with -DVR_KBD=<-1|1|4|9|16|25>"
#endif

namespace vrb {

namespace {

template <typename K>
int resident_ctas(K kernel, size_t smem, int num_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * num_sms;
}

template <int KBD, bool TOP, bool COUNT, int OUT>
cudaError_t launch_tile(const LaunchDev& P, const LaunchCfg& cfg) {
    const size_t smem = march_smem_bytes<TOP>(P.tree.max_depth) + basis_smem_bytes<KBD>();
    dim3 grid((P.w + kTileW - 1) / kTileW, (P.h + kTileH - 1) / kTileH, P.n_views);
    march_tile_kernel<KBD, TOP, COUNT, OUT><<<grid, kBlock, smem, cfg.stream>>>(P);
    return cudaGetLastError();
}

template <int KBD, bool TOP, bool COUNT, int OUT, int TUNE = 0>
cudaError_t launch_persistent(LaunchDev& P, const LaunchCfg& cfg) {
    const size_t smem = march_smem_bytes<TOP, (TUNE & kTuneWide) != 0 && !TOP>(P.tree.max_depth) + basis_smem_bytes<KBD>();
    static int cached_ctas = 0, cached_depth = -1;
    if (cached_ctas == 0 || cached_depth != P.tree.max_depth) {
        cached_ctas = resident_ctas(march_persistent_kernel<KBD, TOP, COUNT, OUT, TUNE>, smem, cfg.num_sms);
        cached_depth = P.tree.max_depth;
    }
    P.tiles_x = (P.w + kTW - 1) / kTW;
    P.tiles_y = (P.h + kTH - 1) / kTH;
    P.n_tiles = P.tiles_x * P.tiles_y * P.n_views;
    set_div((uint32_t)(P.tiles_x * P.tiles_y), P.div_view_mul, P.div_view_shift);
    set_div((uint32_t)P.tiles_x, P.div_row_mul, P.div_row_shift);
    P.work_counter = cfg.queue;
    int grid = cached_ctas;
    const int need = (P.n_tiles + (kBlock / 32) - 1) / (kBlock / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kBlock); lc.dynamicSmemBytes = smem; lc.stream = cfg.stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cfg.pdl) {  // programmatic dependent launch: see pdl_wait_predecessor() in vr_march.cuh
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cfg.l2_window_bytes) {  // keep the node table resident in the persisting part of L2
        attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[na].val.accessPolicyWindow.base_ptr = const_cast<void*>(cfg.l2_window);
        attr[na].val.accessPolicyWindow.num_bytes = cfg.l2_window_bytes;
        attr[na].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++na;
    }
    lc.attrs = attr; lc.numAttrs = na;
    return cudaLaunchKernelEx(&lc, march_persistent_kernel<KBD, TOP, COUNT, OUT, TUNE>, P);
}

template <int KBD, bool TOP, bool COUNT, int OUT>
cudaError_t launch_deferred(LaunchDev& P, const LaunchCfg& cfg) {
    const size_t smem = deferred_smem_bytes<TOP>(P.tree.max_depth);
    static int cached_ctas = 0, cached_depth = -1;
    if (cached_ctas == 0 || cached_depth != P.tree.max_depth) {
        cached_ctas = resident_ctas(march_deferred_kernel<KBD, TOP, COUNT, OUT>, smem, cfg.num_sms);
        cached_depth = P.tree.max_depth;
    }
    P.tiles_x = (P.w + kTW - 1) / kTW;
    P.tiles_y = (P.h + kTH - 1) / kTH;
    P.n_tiles = P.tiles_x * P.tiles_y * P.n_views;
    set_div((uint32_t)(P.tiles_x * P.tiles_y), P.div_view_mul, P.div_view_shift);
    set_div((uint32_t)P.tiles_x, P.div_row_mul, P.div_row_shift);
    P.work_counter = cfg.queue;
    int grid = cached_ctas;
    const int need = (P.n_tiles + (kBlock / 32) - 1) / (kBlock / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    march_deferred_kernel<KBD, TOP, COUNT, OUT><<<grid, kBlock, smem, cfg.stream>>>(P);
    return cudaGetLastError();
}

}  // namespace

template <int KBD>
cudaError_t launch_march(LaunchDev& P, const LaunchCfg& cfg) {
    const int kind = cfg.variant & 15, tune = cfg.variant >> 4;
    const bool top = (kind == 2 || kind == 4 || kind == 6);
    const bool persistent = (kind == 3 || kind == 4);
    const bool deferred = kind >= 5;
    if (kind == 3 && tune && !cfg.surface && !cfg.count) {  // measurement knobs on the persistent kernel
        switch (tune) {
            case 1: return launch_persistent<KBD, false, false, kOutLinear, 1>(P, cfg);
            case 2: return launch_persistent<KBD, false, false, kOutLinear, 2>(P, cfg);
            case 3: return launch_persistent<KBD, false, false, kOutLinear, 3>(P, cfg);
            case 8: return launch_persistent<KBD, false, false, kOutLinear, 8>(P, cfg);
            case 10: return launch_persistent<KBD, false, false, kOutLinear, 10>(P, cfg);
            case 17: return launch_persistent<KBD, false, false, kOutLinear, 17>(P, cfg);
            case 65: return launch_persistent<KBD, false, false, kOutLinear, 65>(P, cfg);
            case 193: return launch_persistent<KBD, false, false, kOutLinear, 193>(P, cfg);
            case 64: return launch_persistent<KBD, false, false, kOutLinear, 64>(P, cfg);
            case 16: return launch_persistent<KBD, false, false, kOutLinear, 16>(P, cfg);
            default: return cudaErrorInvalidValue;
        }
    }
    if (cfg.surface) {  // drop-in launch_renderer path: default kernel writing the caller's cudaArray
        return launch_persistent<KBD, false, false, kOutSurface, 193>(P, cfg);
    }
    if (cfg.count) {  // instrumented build of the default kernel
        return launch_persistent<KBD, false, true, kOutLinear, 193>(P, cfg);
    }
    if (deferred) {
        return top ? launch_deferred<KBD, true, false, kOutLinear>(P, cfg)
                   : launch_deferred<KBD, false, false, kOutLinear>(P, cfg);
    }
    if (persistent) {
        return top ? launch_persistent<KBD, true, false, kOutLinear>(P, cfg)
                   : launch_persistent<KBD, false, false, kOutLinear>(P, cfg);
    }
    return top ? launch_tile<KBD, true, false, kOutLinear>(P, cfg)
               : launch_tile<KBD, false, false, kOutLinear>(P, cfg);
}

template cudaError_t launch_march<VR_KBD>(LaunchDev&, const LaunchCfg&);

}  // namespace vrb
