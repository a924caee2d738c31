/*
 * march_oracle.c -- CPU restatement of volrend's PlenOctree ray march.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the reported CPU
 * baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (volrend_b200/csrc) never does.
 *
 * Pinning: the reference ships no golden vectors for this path.  The
 * oracle is pinned against outputs of the reference's own CUDA kernel
 * (oracle/_ref/libvolrend_ref.so, built by oracle/Makefile from /root/reference and run
 * on the B200) committed under tests/golden/ by tools/make_golden.py.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Floating-point evaluation order follows the SASS nvcc 12.9 emits
 * for the reference at -O3 (default -fmad=true): where the compiler contracts a*b+c
 * the oracle calls fmaf(), everywhere else plain rounded ops; compile with
 * -ffp-contract=off so gcc adds no contractions of its own.  All operations on the
 * sample-position path (ray set-up, slab test, descent, cell exit, t update) are
 * IEEE-exact on both sides, so the visited leaf sequence is bit-identical to the GPU's;
 * colours differ only through expf/cosf ulps (glibc vs libdevice).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "march_oracle.h"

/* ---------------------------------------------------------------- fp16 -> fp32 */
static float g_h2f[65536];
static pthread_once_t g_h2f_once = PTHREAD_ONCE_INIT;

static float half_bits_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal half -> normal float */
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static void init_h2f(void) {
    for (int i = 0; i < 65536; ++i) g_h2f[i] = half_bits_to_float((uint16_t)i);
}
static inline float h2f(uint16_t h) { return g_h2f[h]; }

/* ---------------------------------------------------------------- small helpers */
/* include/volrend/cuda/common.cuh:12-16 : sqrtf(x*x + y*y + z*z), contracted by nvcc
 * as fma(z,z, fma(x,x, y*y)). */
static inline float norm3(const float* v) {
    return sqrtf(fmaf(v[2], v[2], fmaf(v[0], v[0], v[1] * v[1])));
}
/* common.cuh:18-23 */
static inline void normalize3(float* v) {
    float inv = 1.f / norm3(v);
    float x = v[0] * inv, y = inv * v[1], z = inv * v[2];
    v[0] = x; v[1] = y; v[2] = z;
}
/* common.cuh:43-47 : u0*v0 + u1*v1 + u2*v2 -> fma(u2,v2, fma(u0,v0, u1*v1)) */
static inline float dot3(const float* u, const float* v) {
    return fmaf(u[2], v[2], fmaf(u[0], v[0], u[1] * v[1]));
}

/* src/cuda/volrend.cu:22-32 screen2worlddir + common.cuh:26-33 _mv3 (v[2] = -1 folds
 * the third product into a subtraction). */
static void screen2worlddir(int ix, int iy, const orc_camera* cam, float* dir, float* cen) {
    const float* m = cam->c2w;
    float vx = ((float)ix - (float)cam->width * 0.5f) / cam->fx;
    float vy = -((float)iy - (float)cam->height * 0.5f) / cam->fy;
    dir[0] = fmaf(vx, m[0], vy * m[3]) - m[6];
    dir[1] = fmaf(vx, m[1], vy * m[4]) - m[7];
    dir[2] = fmaf(vx, m[2], vy * m[5]) - m[8];
    normalize3(dir);
    cen[0] = m[9]; cen[1] = m[10]; cen[2] = m[11];
}

/* src/cuda/volrend.cu:34-54 maybe_world2ndc */
static void maybe_world2ndc(const orc_tree* tree, float* dir, float* cen) {
    if (!(tree->ndc_width > 0.f)) return;
    float t = -(cen[2] + 1.f) / dir[2];
    for (int i = 0; i < 3; ++i) cen[i] = fmaf(t, dir[i], cen[i]);
    float m2f = tree->ndc_focal * -2.0f;
    float kx = m2f / tree->ndc_width;
    float ky = m2f / tree->ndc_height;
    float dx = dir[0] / dir[2], cx = cen[0] / cen[2];
    float dy = dir[1] / dir[2], cy = cen[1] / cen[2];
    float n0 = kx * (dx - cx);
    float n1 = ky * (dy - cy);
    float n2 = -2.0f / cen[2];
    float c0 = kx * cx;
    float c1 = ky * cy;
    float c2 = 2.0f / cen[2] + 1.0f;
    dir[0] = n0; dir[1] = n1; dir[2] = n2;
    cen[0] = c0; cen[1] = c1; cen[2] = c2;
    normalize3(dir);
}

/* src/cuda/volrend.cu:57-71 rodrigues (mixed fp32 / fp64 as in the reference) */
static void rodrigues(const float* aa, float* dir) {
    float tmp[3] = {aa[0], aa[1], aa[2]};
    float angle = norm3(tmp);
    if ((double)angle < 1e-6) return;
    float k[3];
    for (int i = 0; i < 3; ++i) k[i] = aa[i] / angle;
    float ca = cosf(angle), sa = sinf(angle);
    float cr[3] = {k[1] * dir[2] - k[2] * dir[1], k[2] * dir[0] - k[0] * dir[2],
                   k[0] * dir[1] - k[1] * dir[0]};
    float d = dot3(k, dir);
    double omc = 1.0 - (double)ca;
    for (int i = 0; i < 3; ++i) {
        float a = fmaf(cr[i], sa, dir[i] * ca);
        dir[i] = (float)((double)a + (double)(k[i] * d) * omc);
    }
}

/* include/volrend/internal/lumisphere.hpp:9-87 maybe_precalc_basis.  The SH constants
 * are double literals in the reference, so each product is evaluated in double and
 * rounded to float on store; the float sub-expressions keep nvcc's contraction. */
static void precalc_basis(const orc_tree* tree, const float* dir, float* out) {
    const int bd = tree->basis_dim;
    if (tree->format == ORC_FMT_ASG) {
        const float* p = tree->extra;
        for (int i = 0; i < bd; ++i) {
            float S = dot3(dir, p + 8), dx = dot3(dir, p + 2), dy = dot3(dir, p + 5);
            float a = dx * (dx * -p[0]);
            float b = dy * (p[1] * dy);
            out[i] = S * expf(a - b) / (float)bd;
            p += 11;
        }
    } else if (tree->format == ORC_FMT_SG) {
        const float* p = tree->extra;
        for (int i = 0; i < bd; ++i) {
            out[i] = expf(p[0] * (dot3(dir, p + 1) - 1.f)) / (float)bd;
            p += 4;
        }
    } else if (tree->format == ORC_FMT_SH) {
        out[0] = (float)0.28209479177387814;
        const float x = dir[0], y = dir[1], z = dir[2];
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, yz = y * z, xz = x * z;
        const double dxy = xy, dyz = yz, dxz = xz, dx = x, dy = y, dz = z;
        /* float sub-expressions as ptxas fuses them in the reference build (SASS of lumisphere.hpp:47-66) */
        const float xx_yy = xx - yy;
        const float t3xx_yy = fmaf(xx, 3.f, -yy);
        const float xx_3yy = fmaf(yy, -3.f, xx);
        if (bd >= 25) {
            const float z7_1 = fmaf(zz, 7.f, -1.f);
            const float z7_3 = fmaf(zz, 7.f, -3.f);
            out[16] = (float)((dxy * 2.5033429417967046) * (double)xx_yy);
            out[17] = (float)((dyz * -1.7701307697799304) * (double)t3xx_yy);
            out[18] = (float)((dxy * 0.9461746957575601) * (double)z7_1);
            out[19] = (float)((dyz * -0.6690465435572892) * (double)z7_3);
            out[20] = (float)((double)fmaf(zz, fmaf(zz, 35.f, -30.f), 3.f) * 0.10578554691520431);
            out[21] = (float)((dxz * -0.6690465435572892) * (double)z7_3);
            out[22] = (float)(((double)xx_yy * 0.47308734787878004) * (double)z7_1);
            out[23] = (float)((dxz * -1.7701307697799304) * (double)xx_3yy);
            out[24] = (float)((double)fmaf(xx, xx_3yy, -(yy * t3xx_yy)) * 0.6258357354491761);
        }
        if (bd >= 16) {
            const float z4 = -yy + fmaf(zz, 4.f, -xx);
            const float q = fmaf(yy, -3.f, fmaf(xx, -3.f, zz + zz));
            out[9] = (float)((dy * -0.5900435899266435) * (double)t3xx_yy);
            out[10] = (float)((dxy * 2.890611442640554) * dz);
            out[11] = (float)((dy * -0.4570457994644658) * (double)z4);
            out[12] = (float)((dz * 0.3731763325901154) * (double)q);
            out[13] = (float)((dx * -0.4570457994644658) * (double)z4);
            out[14] = (float)((dz * 1.445305721320277) * (double)xx_yy);
            out[15] = (float)((dx * -0.5900435899266435) * (double)xx_3yy);
        }
        if (bd >= 9) {
            out[4] = (float)(dxy * 1.0925484305920792);
            out[5] = (float)(dyz * -1.0925484305920792);
            out[6] = (float)((((double)zz + (double)zz) - (double)xx - (double)yy) * 0.31539156525252005);
            out[7] = (float)(dxz * -1.0925484305920792);
            out[8] = (float)((double)(xx - yy) * 0.5462742152960396);
        }
        if (bd >= 4) {
            out[1] = (float)(dy * -0.4886025119029199);
            out[2] = (float)(dz * 0.4886025119029199);
            out[3] = (float)(dx * -0.4886025119029199);
        }
    }
}

/* One group of the colour dot product, include/volrend/cuda/rt_core.cuh:130-162:
 * `tmp += B[lo]*k[lo] + ... + B[hi]*k[hi]` compiles to
 * s = B[lo+1]*k[lo+1]; s = fma(B[lo],k[lo],s); s = fma(B[j],k[j],s) j=lo+2..hi; tmp += s */
static inline float sh_group(const float* B, const uint16_t* k, int lo, int hi) {
    float s = B[lo + 1] * h2f(k[lo + 1]);
    s = fmaf(B[lo], h2f(k[lo]), s);
    for (int j = lo + 2; j <= hi; ++j) s = fmaf(B[j], h2f(k[j]), s);
    return s;
}

/* include/volrend/cuda/rt_core.cuh:66-196 trace_ray, with
 * include/volrend/internal/n3tree_query.hpp:13-48 query_single_from_root,
 * rt_core.cuh:18-34 _dda_world, :37-49 _dda_unit, :52-63 _get_delta_scale inlined. */
static void trace_ray(const orc_tree* tree, float* dir, const float* vdir, const float* cen,
                      const orc_options* opt, float tmax_bg, float* out, orc_counters* cnt) {
    /* _get_delta_scale */
    dir[0] = tree->scale[0] * dir[0];
    dir[1] = tree->scale[1] * dir[1];
    dir[2] = tree->scale[2] * dir[2];
    const float delta_scale = 1.f / norm3(dir);
    {
        float x = dir[0] * delta_scale, y = delta_scale * dir[1], z = delta_scale * dir[2];
        dir[0] = x; dir[1] = y; dir[2] = z;
    }
    tmax_bg = tmax_bg / delta_scale;

    float invdir[3];
    for (int i = 0; i < 3; ++i) invdir[i] = (float)(1.0 / ((double)dir[i] + 1e-9)); /* :83 [f64] */
    /* _dda_world :18-34 [f64] */
    float tmin = 0.f, tmax = 1e4f;
    for (int i = 0; i < 3; ++i) {
        float t1 = (float)((((double)opt->render_bbox[i] + 1e-6) - (double)cen[i]) * (double)invdir[i]);
        float t2 = (float)((((double)opt->render_bbox[i + 3] - 1e-6) - (double)cen[i]) * (double)invdir[i]);
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
    }
    tmax = fminf(tmax, tmax_bg);

    if (tmax < 0.f || tmin > tmax) {                                   /* :88-92 */
        if (opt->render_depth) out[3] = 1.f;
        return;
    }
    if (cnt) cnt->rays_hit += 1;

    float basis_fn[ORC_BASIS_MAX];
    memset(basis_fn, 0, sizeof(basis_fn));
    precalc_basis(tree, vdir, basis_fn);
    for (int i = 0; i < opt->basis_minmax[0] && i < ORC_BASIS_MAX; ++i) basis_fn[i] = 0.f;  /* :98-100 */
    for (int i = opt->basis_minmax[1] + 1; i < ORC_BASIS_MAX; ++i)                           /* :101-103 */
        if (i >= 0) basis_fn[i] = 0.f;

    const float fN = (float)tree->N;
    const int64_t N3 = (int64_t)tree->N * tree->N * tree->N;
    const int bd = tree->basis_dim;
    float light = 1.f;
    float t = tmin;
    while (t < tmax) {                                                 /* :108 */
        float pos[3];
        for (int i = 0; i < 3; ++i) pos[i] = fmaf(t, dir[i], cen[i]);  /* :109-111 */

        /* query_single_from_root, n3tree_query.hpp:16-47 */
        for (int i = 0; i < 3; ++i) pos[i] = fmaxf(fminf(pos[i], 1.f - 1e-6f), 0.f);
        int64_t ptr = 0, sub_ptr;
        float cube_sz = fN;
        uint64_t depth = 0;
        for (;;) {
            float index = 0.f;
            for (int i = 0; i < 3; ++i) {
                pos[i] = pos[i] * fN;
                const float k = floorf(pos[i]);
                index = fmaf(index, fN, k);
                pos[i] = pos[i] - k;
            }
            sub_ptr = ptr + (int32_t)index;
            const int64_t skip = tree->child[sub_ptr];
            ++depth;
            if (skip == 0) break;
            cube_sz = cube_sz * fN;
            ptr += skip * N3;
        }
        const uint16_t* leaf = tree->data + sub_ptr * tree->data_dim;
        if (cnt) { cnt->samples += 1; cnt->child_loads += depth; }

        /* _dda_unit :37-49 */
        float tsub = 1e4f;
        for (int i = 0; i < 3; ++i) {
            float t1 = invdir[i] * -pos[i];
            float t2 = invdir[i] + t1;
            tsub = fminf(tsub, fmaxf(t1, t2));
        }
        const float t_subcube = tsub / cube_sz;                        /* :116 */
        const float delta_t = t_subcube + opt->step_size;              /* :117 */
        const float sigma = h2f(leaf[tree->data_dim - 1]);
        if (sigma > opt->sigma_thresh) {                               /* :118 */
            const float att = expf(((-delta_t) * delta_scale) * sigma);   /* :119 */
            const float weight = light * (1.f - att);                  /* :120 */
            if (cnt) cnt->shaded += 1;
            if (opt->render_depth) {
                out[0] = fmaf(weight, t, out[0]);                       /* :122-123 */
            } else if (bd >= 0) {                                      /* :125-165 */
                for (int c = 0; c < 3; ++c) {
                    const uint16_t* k = leaf + c * bd;
                    float tmp = basis_fn[0] * h2f(k[0]);
                    if (bd == 25) tmp = tmp + sh_group(basis_fn, k, 16, 24);
                    if (bd == 25 || bd == 16) tmp = tmp + sh_group(basis_fn, k, 9, 15);
                    if (bd == 25 || bd == 16 || bd == 9) tmp = tmp + sh_group(basis_fn, k, 4, 8);
                    if (bd == 25 || bd == 16 || bd == 9 || bd == 4) tmp = tmp + sh_group(basis_fn, k, 1, 3);
                    out[c] = out[c] + weight / (1.f + expf(-tmp));      /* :163 */
                }
            } else {
                for (int c = 0; c < 3; ++c) out[c] = fmaf(h2f(leaf[c]), weight, out[c]);  /* :167-171 */
            }
            light = light * att;                                        /* :174 */
            if (light < opt->stop_thresh) {                             /* :176-185 */
                if (opt->render_depth) out[0] = out[1] = out[2] = fminf(out[0] * 0.3f, 1.0f);
                const float scale = 1.f / (1.f - light);
                out[0] *= scale; out[1] *= scale; out[2] *= scale;
                out[3] = 1.f;
                return;
            }
        }
        t = t + delta_t;                                                /* :187 */
    }
    if (opt->render_depth) {                                            /* :189-194 */
        out[0] = out[1] = out[2] = fminf(out[0] * 0.3f, 1.0f);
        out[3] = 1.f;
    } else {
        out[3] = 1.f - light;
    }
}

/* src/cuda/volrend.cu:78-173 render_kernel for one pixel (probe overlay excluded: GUI). */
static void render_pixel(const orc_tree* tree, const orc_camera* cam, const orc_options* opt,
                         int x, int y, const uint8_t* rgba_in, const float* depth_in,
                         float* out, orc_counters* cnt) {
    float dir[3], cen[3];
    out[0] = out[1] = out[2] = out[3] = 0.f;
    if (tree->N > 0) {
        screen2worlddir(x, y, cam, dir, cen);                            /* :136 */
        float vdir[3] = {dir[0], dir[1], dir[2]};                        /* :137 */
        maybe_world2ndc(tree, dir, cen);                                 /* :138 */
        for (int i = 0; i < 3; ++i) cen[i] = fmaf(tree->scale[i], cen[i], tree->offset[i]);  /* :139-141 */
        float t_max = 1e9f;                                              /* :143 */
        if (depth_in) t_max = *depth_in;                                 /* :144-146 */
        rodrigues(opt->rot_dirs, vdir);                                  /* :148 */
        trace_ray(tree, dir, vdir, cen, opt, t_max, out, cnt);           /* :150 */
    }
    const float nalpha = 1.f - out[3];                                   /* :153 */
    if (!rgba_in) {                                                      /* offscreen :154-158 */
        const float remain = nalpha * opt->background_brightness;
        out[0] = remain + out[0]; out[1] = remain + out[1]; out[2] = remain + out[2];
    } else {                                                             /* :159-163 */
        for (int c = 0; c < 3; ++c) out[c] = fmaf((float)rgba_in[c] / 255.f, nalpha, out[c]);
    }
}

/* volrend.cu:166 : uint8_t(out*255) -- F2I.U32.TRUNC then & 0xff in the SASS. */
static inline uint8_t quant8(float v) {
    float s = v * 255.f;
    if (!(s > 0.f)) return 0;
    if (s >= 4294967296.f) return 0xff;
    return (uint8_t)((uint32_t)s & 0xffu);
}

typedef struct {
    const orc_tree* tree; const orc_camera* cam; const orc_options* opt;
    int x0, y0, w, h;
    const uint8_t* rgba_in; const float* depth_in;
    float* rgba_f32; uint8_t* rgba8;
    orc_counters cnt; int want_cnt;
    volatile int* next_row;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        int r = __sync_fetch_and_add(j->next_row, 1);
        if (r >= j->h) break;
        for (int c = 0; c < j->w; ++c) {
            float out[4];
            size_t o = (size_t)r * j->w + c;
            render_pixel(j->tree, j->cam, j->opt, j->x0 + c, j->y0 + r,
                         j->rgba_in ? j->rgba_in + 4 * o : NULL,
                         j->depth_in ? j->depth_in + o : NULL, out, j->want_cnt ? &j->cnt : NULL);
            if (j->rgba_f32) memcpy(j->rgba_f32 + 4 * o, out, 16);
            if (j->rgba8) {
                j->rgba8[4 * o + 0] = quant8(out[0]);
                j->rgba8[4 * o + 1] = quant8(out[1]);
                j->rgba8[4 * o + 2] = quant8(out[2]);
                j->rgba8[4 * o + 3] = 255;
            }
        }
    }
    return NULL;
}

int orc_render(const orc_tree* tree, const orc_camera* cam, const orc_options* opt,
               int x0, int y0, int w, int h, const uint8_t* rgba_in, const float* depth_in,
               float* rgba_f32, uint8_t* rgba8, orc_counters* counters, int nthreads) {
    if (!tree || !cam || !opt || w < 0 || h < 0) return -1;
    if (tree->N != 0 && tree->N != 2) return -2; /* reference: "N != 2 probably doesn't work" */
    pthread_once(&g_h2f_once, init_h2f);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > h && h > 0) nthreads = h;
    volatile int next_row = 0;
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; ++i) {
        job_t* j = &jobs[i];
        j->tree = tree; j->cam = cam; j->opt = opt;
        j->x0 = x0; j->y0 = y0; j->w = w; j->h = h;
        j->rgba_in = rgba_in; j->depth_in = depth_in;
        j->rgba_f32 = rgba_f32; j->rgba8 = rgba8;
        j->want_cnt = counters != NULL; j->next_row = &next_row;
        if (i > 0) pthread_create(&th[i], NULL, worker, j);
    }
    worker(&jobs[0]);
    for (int i = 1; i < nthreads; ++i) pthread_join(th[i], NULL);
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (int i = 0; i < nthreads; ++i) {
            counters->samples += jobs[i].cnt.samples;
            counters->child_loads += jobs[i].cnt.child_loads;
            counters->shaded += jobs[i].cnt.shaded;
            counters->rays_hit += jobs[i].cnt.rays_hit;
        }
    }
    free(jobs); free(th);
    return 0;
}

void orc_default_options(orc_options* o) {
    /* include/volrend/render_options.hpp:11-53 */
    memset(o, 0, sizeof(*o));
    o->step_size = 1e-4f; o->sigma_thresh = 1e-2f; o->stop_thresh = 1e-2f;
    o->background_brightness = 1.f;
    o->render_bbox[3] = o->render_bbox[4] = o->render_bbox[5] = 1.f;
    o->basis_minmax[0] = 0; o->basis_minmax[1] = ORC_BASIS_MAX - 1;
}
