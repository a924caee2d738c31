/* march_oracle.h -- C interface of the CPU restatement (test infrastructure only; see
 * march_oracle.c).  Plain-old-data mirrors of the reference's kernel arguments:
 *   orc_tree    <- internal::TreeSpec   (include/volrend/internal/data_spec.hpp:23-50)
 *   orc_camera  <- internal::CameraSpec (data_spec.hpp:11-22) with c2w by value
 *   orc_options <- RenderOptions        (include/volrend/render_options.hpp:11-53) */
#ifndef MARCH_ORACLE_H_
#define MARCH_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_BASIS_MAX 25 /* VOLREND_GLOBAL_BASIS_MAX, render_options.hpp:6 */

enum { ORC_FMT_RGBA = 0, ORC_FMT_SH = 1, ORC_FMT_SG = 2, ORC_FMT_ASG = 3 }; /* data_format.hpp:9-15 */

typedef struct {
    const int32_t* child;  /* [capacity*N^3] relative node offsets, 0 = leaf */
    const uint16_t* data;  /* [capacity*N^3*data_dim] fp16 bits */
    const float* extra;    /* SG: [basis_dim*4], ASG: [basis_dim*11], else NULL */
    int64_t capacity;
    int32_t N;             /* 0 = no tree loaded (pure background), else 2 */
    int32_t data_dim;
    int32_t format;        /* ORC_FMT_* */
    int32_t basis_dim;     /* -1 for RGBA */
    float offset[3];
    float scale[3];
    float ndc_width;       /* <= 0: NDC disabled (data_spec.hpp:47) */
    float ndc_height;
    float ndc_focal;
} orc_tree;

typedef struct {
    int32_t width, height;
    float fx, fy;
    float c2w[12];         /* glm::mat4x3 column-major: right, up, back, centre */
} orc_camera;

typedef struct {
    float step_size, sigma_thresh, stop_thresh, background_brightness;
    float render_bbox[6];
    int32_t basis_minmax[2];
    float rot_dirs[3];
    int32_t render_depth;
} orc_options;

typedef struct {
    uint64_t samples;      /* S: march-loop iterations (rt_core.cuh:108) */
    uint64_t child_loads;  /* sum of d_s: child[] reads (n3tree_query.hpp:36-37) */
    uint64_t shaded;       /* S_shaded: samples with sigma > sigma_thresh */
    uint64_t rays_hit;     /* rays that enter the march loop set-up */
} orc_counters;

/* Render the tile [x0,x0+w) x [y0,y0+h) of the cam->width x cam->height frame.
 * rgba_in/depth_in NULL => offscreen (background_brightness compositing, t_max=1e9);
 * otherwise tile-sized RGBA8 / float buffers to composite over (volrend.cu:92-96,143-146).
 * rgba_f32: tile-sized float4 (premultiplied RGB + bg, alpha) or NULL.
 * rgba8: tile-sized truncated bytes (volrend.cu:166) or NULL.  Returns 0 on success. */
int orc_render(const orc_tree* tree, const orc_camera* cam, const orc_options* opt,
               int x0, int y0, int w, int h, const uint8_t* rgba_in, const float* depth_in,
               float* rgba_f32, uint8_t* rgba8, orc_counters* counters, int nthreads);

void orc_default_options(orc_options* o);

#ifdef __cplusplus
}
#endif
#endif
