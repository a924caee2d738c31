/* lane_sim.c -- design tool (not product, not a test): replays the exact per-ray sample
 * sequence of a frame (oracle arithmetic) and counts WARP-level instruction issue under
 * different lane-scheduling policies, so that kernel restructurings can be ranked before
 * GPU time is spent on them.
 *
 *   P0  one 4x8 tile per warp pass, inline shading (the round-1 kernel)
 *   P1  P0 + warp-shared shading queue drained 32 items at a time
 *   P2  ray-granular refill from a per-warp pool of set-up rays (threshold = idle lanes) + P1 queue
 *
 * Cost constants (warp instructions) come from the SASS of the round-1 kernel.
 * Usage: lane_sim <dump.bin> ; the dump is written by tools/lane_sim.py.
 */
#include "../oracle/march_oracle.c"

#include <stdio.h>

typedef struct {
    uint32_t off;   /* first sample in the trace */
    uint32_t n;     /* samples */
    uint8_t hit;
} ray_t;

static double g_act_hist[33];
static const int g_p3_theta[6] = {4, 8, 12, 8, 8, 4}, g_p3_pat[6] = {1, 1, 1, 8, 16, 16};
static int g_b[32];   /* cumulative levels resolved after table k */
static uint8_t* g_trace;  /* per sample: fetches (low 4 bits) | shaded << 7 */
static size_t g_trace_n, g_trace_cap;

static void push(uint8_t v) {
    if (g_trace_n == g_trace_cap) {
        g_trace_cap = g_trace_cap ? g_trace_cap * 2 : (1u << 24);
        g_trace = (uint8_t*)realloc(g_trace, g_trace_cap);
    }
    g_trace[g_trace_n++] = v;
}

/* the march of trace_ray() above without colours; records wide-table fetch counts */
static void trace_one(const orc_tree* tree, const orc_camera* cam, const orc_options* opt, int x, int y, ray_t* R) {
    float dir[3], cen[3];
    R->off = (uint32_t)g_trace_n; R->n = 0; R->hit = 0;
    screen2worlddir(x, y, cam, dir, cen);
    maybe_world2ndc(tree, dir, cen);
    for (int i = 0; i < 3; ++i) cen[i] = fmaf(tree->scale[i], cen[i], tree->offset[i]);
    float ds;
    {
        for (int i = 0; i < 3; ++i) dir[i] = tree->scale[i] * dir[i];
        ds = 1.f / norm3(dir);
        float a = dir[0] * ds, b = ds * dir[1], c = ds * dir[2];
        dir[0] = a; dir[1] = b; dir[2] = c;
    }
    float invdir[3];
    for (int i = 0; i < 3; ++i) invdir[i] = (float)(1.0 / ((double)dir[i] + 1e-9));
    float tmin = 0.f, tmax = 1e4f;
    for (int i = 0; i < 3; ++i) {
        float t1 = (float)((((double)opt->render_bbox[i] + 1e-6) - (double)cen[i]) * (double)invdir[i]);
        float t2 = (float)((((double)opt->render_bbox[i + 3] - 1e-6) - (double)cen[i]) * (double)invdir[i]);
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
    }
    tmax = fminf(tmax, 1e9f / ds);
    if (tmax < 0.f || tmin > tmax) return;
    R->hit = 1;
    float light = 1.f, t = tmin;
    uint32_t pu[3] = {0, 0, 0};
    int pdepth = 1;
    while (t < tmax) {
        float pos[3];
        uint32_t u[3];
        for (int i = 0; i < 3; ++i) {
            pos[i] = fmaf(t, dir[i], cen[i]);
            pos[i] = fmaxf(fminf(pos[i], 1.f - 1e-6f), 0.f);
            u[i] = (uint32_t)(pos[i] * 16777216.f);
        }
        int64_t ptr = 0, sub;
        float cube = 2.f;
        int depth = 0;
        for (;;) {
            float index = 0.f;
            for (int i = 0; i < 3; ++i) {
                pos[i] = pos[i] * 2.f;
                const float k = floorf(pos[i]);
                index = fmaf(index, 2.f, k);
                pos[i] -= k;
            }
            sub = ptr + (int32_t)index;
            const int64_t skip = tree->child[sub];
            ++depth;
            if (!skip) break;
            cube *= 2.f;
            ptr += skip * 8;
        }
        /* wide-table fetches (vr_march.cuh find_leaf_wide) */
        const uint32_t diff = (u[0] ^ pu[0]) | (u[1] ^ pu[1]) | (u[2] ^ pu[2]);
        int common = diff ? (__builtin_clz(diff) - 8) : 24;
        /* table k covers the levels (g_b[k-1], g_b[k]] and hangs off a node of depth g_b[k-1] (LEVELS / PHASE env) */
        int shared = common < pdepth - 1 ? common : pdepth - 1;   /* deepest node depth shared and on the previous path */
        int j0 = 0, jl = 0;
        while (g_b[j0] <= shared) ++j0;        /* table j0 is rooted at depth g_b[j0-1] <= shared */
        while (g_b[jl] < depth) ++jl;
        int fetches = jl - j0 + 1;
        if (fetches < 1) fetches = 1;
        pu[0] = u[0]; pu[1] = u[1]; pu[2] = u[2]; pdepth = depth;
        float tsub = 1e4f;
        for (int i = 0; i < 3; ++i) {
            float t1 = invdir[i] * -pos[i];
            float t2 = invdir[i] + t1;
            tsub = fminf(tsub, fmaxf(t1, t2));
        }
        const float dt = tsub / cube + opt->step_size;
        const float sigma = h2f(tree->data[sub]);  /* data here = sigma only (data_dim 1) */
        int shaded = 0, stop = 0;
        if (sigma > opt->sigma_thresh) {
            shaded = 1;
            const float att = expf(((-dt) * ds) * sigma);
            light *= att;
            if (light < opt->stop_thresh) stop = 1;
        }
        push((uint8_t)(fetches | (shaded << 7)));
        R->n++;
        if (stop) break;
        t += dt;
    }
}

/* ---- cost model (warp instructions) */
static double C_SETUP = 330, C_OUT = 25, C_BODY = 105, C_FETCH = 20, C_SHADE = 190;   /* body/fetch: SASS of the r2 queue kernel */
static double C_ENQ = 55, C_DRAIN = 250, C_ACC = 10, C_VOTE = 4, C_REFILL = 90, C_BATCH_EXTRA = 40, C_TILE = 30;

typedef struct { double instr, body_iters, body_lanes, shade_iters, shade_lanes, fetch_iters, fetch_lanes; } stats_t;

/* P0 / P1: tiles of 4x8 pixels */
static void sim_tiles(const ray_t* rays, int W, int H, int shared_queue, stats_t* S) {
    const int TW = 4, TH = 8;
    for (int ty = 0; ty < (H + TH - 1) / TH; ++ty)
        for (int tx = 0; tx < (W + TW - 1) / TW; ++tx) {
            const ray_t* lane[32];
            uint32_t pos[32];
            int alive = 0;
            for (int l = 0; l < 32; ++l) {
                int x = tx * TW + l % TW, y = ty * TH + l / TW;
                lane[l] = (x < W && y < H) ? &rays[(size_t)y * W + x] : NULL;
                pos[l] = 0;
                if (lane[l] && lane[l]->n) ++alive;
            }
            S->instr += C_TILE + C_SETUP + C_OUT;
            int q = 0;
            while (alive) {
                int act = 0, maxf = 0, nsh = 0;
                int fl[8] = {0};
                for (int l = 0; l < 32; ++l) {
                    if (!lane[l] || pos[l] >= lane[l]->n) continue;
                    uint8_t v = g_trace[lane[l]->off + pos[l]++];
                    int f = v & 15;
                    ++act;
                    if (f > maxf) maxf = f;
                    for (int k = 1; k < f && k < 8; ++k) fl[k]++;
                    if (v & 128) ++nsh;
                    if (pos[l] >= lane[l]->n) --alive;
                }
                S->instr += C_BODY + C_FETCH * (maxf - 1);
                S->body_iters += 1; S->body_lanes += act;
                if (!shared_queue) g_act_hist[act]++;
                for (int k = 1; k < maxf; ++k) { S->fetch_iters += 1; S->fetch_lanes += fl[k]; }
                if (nsh) {
                    if (!shared_queue) {
                        S->instr += C_SHADE; S->shade_iters += 1; S->shade_lanes += nsh;
                    } else {
                        S->instr += C_ENQ;
                        q += nsh;
                        if (q >= 32) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += 32; q -= 32; }
                    }
                }
            }
            if (shared_queue && q) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += q; }
        }
}

/* P2: per-warp pool of set-up rays, refill when idle lanes >= theta.  A warp handles `run`
 * consecutive tiles (stand-in for the dynamic global queue). */
static void sim_refill(const ray_t* rays, int W, int H, int theta, int run, int shared_queue, stats_t* S) {
    const int TW = 4, TH = 8;
    const int ntx = (W + TW - 1) / TW, nty = (H + TH - 1) / TH, ntiles = ntx * nty;
    for (int t0 = 0; t0 < ntiles; t0 += run) {
        int tnext = t0, tend = t0 + run < ntiles ? t0 + run : ntiles;
        const ray_t* pool[128];
        int npool = 0;
        const ray_t* lane[32] = {0};
        uint32_t pos[32] = {0};
        int alive = 0, q = 0;
        S->instr += C_TILE;
        for (;;) {
            int idle = 32 - alive;
            if (idle >= theta || alive == 0) {
                /* top the pool up with full-warp batch set-ups */
                while (npool < idle && tnext < tend) {
                    int tx = tnext % ntx, ty = tnext / ntx;
                    ++tnext;
                    S->instr += C_SETUP + C_BATCH_EXTRA + C_OUT;  /* misses are written here */
                    for (int l = 0; l < 32; ++l) {
                        int x = tx * TW + l % TW, y = ty * TH + l / TW;
                        if (x < W && y < H && rays[(size_t)y * W + x].n) pool[npool++] = &rays[(size_t)y * W + x];
                    }
                }
                if (npool == 0 && alive == 0) break;
                if (npool && idle) {
                    /* finished rays need their queued colours before they are written */
                    if (shared_queue && q) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += q; q = 0; }
                    S->instr += C_REFILL + C_OUT;
                    for (int l = 0; l < 32 && npool; ++l)
                        if (!lane[l] || pos[l] >= lane[l]->n) { lane[l] = pool[--npool]; pos[l] = 0; ++alive; }
                }
            }
            if (!alive) continue;
            int act = 0, maxf = 0, nsh = 0;
            int fl[8] = {0};
            for (int l = 0; l < 32; ++l) {
                if (!lane[l] || pos[l] >= lane[l]->n) continue;
                uint8_t v = g_trace[lane[l]->off + pos[l]++];
                int f = v & 15;
                ++act;
                if (f > maxf) maxf = f;
                for (int k = 1; k < f && k < 8; ++k) fl[k]++;
                if (v & 128) ++nsh;
                if (pos[l] >= lane[l]->n) --alive;
            }
            S->instr += C_BODY + C_VOTE + C_FETCH * (maxf - 1);
            S->body_iters += 1; S->body_lanes += act;
            for (int k = 1; k < maxf; ++k) { S->fetch_iters += 1; S->fetch_lanes += fl[k]; }
            if (nsh) {
                if (!shared_queue) { S->instr += C_SHADE; S->shade_iters += 1; S->shade_lanes += nsh; }
                else {
                    S->instr += C_ENQ;
                    q += nsh;
                    if (q >= 32) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += 32; q -= 32; }
                }
            }
        }
        if (shared_queue && q) { S->instr += C_DRAIN + 2 * C_ACC + C_OUT; S->shade_iters += 1; S->shade_lanes += q; }
    }
}

/* P3: P1 + tail parking.  When a tile has had <= theta rays alive for `patience` iterations, the warp flushes its
 * shading queue, parks the surviving rays in a pool and takes the next tile; whenever 32 rays are parked some warp
 * marches them as a dense pass (which may park its own tail again while the pool is not empty). */
static double C_PARK = 70, C_RESUME = 70;
static double g_park_events[16], g_park_rays[16], g_pool_passes[16];
static int g_p3_cur;
static void sim_park(const ray_t* rays, int W, int H, int theta, int patience, stats_t* S) {
    const int TW = 4, TH = 8;
    static const ray_t* pool[1 << 16];
    static uint32_t pool_pos[1 << 16];
    int npool = 0;
    const int ntx = (W + TW - 1) / TW, nty = (H + TH - 1) / TH;
    int tile = 0;
    const int ntiles = ntx * nty;
    for (;;) {
        const ray_t* lane[32];
        uint32_t pos[32];
        int alive = 0, from_pool = 0;
        if (npool >= 32 || (tile >= ntiles && npool > 0)) {
            from_pool = 1;
            S->instr += C_RESUME;
            g_pool_passes[g_p3_cur] += 1;
            for (int l = 0; l < 32; ++l) {
                if (npool) { --npool; lane[l] = pool[npool]; pos[l] = pool_pos[npool]; ++alive; }
                else { lane[l] = NULL; pos[l] = 0; }
            }
        } else if (tile < ntiles) {
            const int tx = tile % ntx, ty = tile / ntx;
            ++tile;
            for (int l = 0; l < 32; ++l) {
                int x = tx * TW + l % TW, y = ty * TH + l / TW;
                lane[l] = (x < W && y < H) ? &rays[(size_t)y * W + x] : NULL;
                pos[l] = 0;
                if (lane[l] && lane[l]->n) ++alive;
            }
            S->instr += C_TILE + C_SETUP + C_OUT;
        } else break;
        int q = 0, low = 0;
        while (alive) {
            int act = 0, maxf = 0, nsh = 0;
            int fl[8] = {0};
            for (int l = 0; l < 32; ++l) {
                if (!lane[l] || pos[l] >= lane[l]->n) continue;
                uint8_t v = g_trace[lane[l]->off + pos[l]++];
                int f = v & 15;
                ++act;
                if (f > maxf) maxf = f;
                for (int k = 1; k < f && k < 8; ++k) fl[k]++;
                if (v & 128) ++nsh;
                if (pos[l] >= lane[l]->n) --alive;
            }
            S->instr += C_BODY + C_FETCH * (maxf - 1);
            S->body_iters += 1; S->body_lanes += act;
            for (int k = 1; k < maxf; ++k) { S->fetch_iters += 1; S->fetch_lanes += fl[k]; }
            if (nsh) {
                S->instr += C_ENQ;
                q += nsh;
                if (q >= 32) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += 32; q -= 32; }
            }
            low = (alive <= theta) ? low + 1 : 0;
            const int more_work = tile < ntiles || npool > 0;
            if (alive && low >= patience && more_work) {   /* park the tail */
                if (q) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += q; q = 0; }
                S->instr += C_PARK + (from_pool ? 0 : 0);
                g_park_events[g_p3_cur] += 1; g_park_rays[g_p3_cur] += alive;
                for (int l = 0; l < 32; ++l)
                    if (lane[l] && pos[l] < lane[l]->n) { pool[npool] = lane[l]; pool_pos[npool] = pos[l]; ++npool; }
                alive = 0;
            }
        }
        if (q) { S->instr += C_DRAIN + 2 * C_ACC; S->shade_iters += 1; S->shade_lanes += q; }
        if (from_pool) S->instr += C_OUT;
    }
}

static void report(const char* name, const stats_t* S, int frames) {
    printf("%-34s %7.2f Minstr/frame  body %5.1f/32 (%6.2fM it)  fetch+ %5.1f/32 (%5.2fM)  shade %5.1f/32 (%5.3fM it)\n", name,
           S->instr / frames / 1e6, S->body_lanes / (S->body_iters + 1e-9), S->body_iters / frames / 1e6,
           S->fetch_lanes / (S->fetch_iters + 1e-9), S->fetch_iters / frames / 1e6,
           S->shade_lanes / (S->shade_iters + 1e-9), S->shade_iters / frames / 1e6);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: lane_sim dump.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    int64_t hdr[4];  /* capacity, W, H, n_cams */
    if (fread(hdr, 8, 4, f) != 4) return 1;
    const int64_t cap = hdr[0];
    const int W = (int)hdr[1], H = (int)hdr[2], ncam = (int)hdr[3];
    float meta[8];  /* offset[3], scale[3], fx, fy */
    if (fread(meta, 4, 8, f) != 8) return 1;
    int32_t* child = (int32_t*)malloc((size_t)cap * 8 * 4);
    uint16_t* sigma = (uint16_t*)malloc((size_t)cap * 8 * 2);
    float* cams = (float*)malloc((size_t)ncam * 12 * 4);
    if (fread(child, 4, (size_t)cap * 8, f) != (size_t)cap * 8) return 1;
    if (fread(sigma, 2, (size_t)cap * 8, f) != (size_t)cap * 8) return 1;
    if (fread(cams, 4, (size_t)ncam * 12, f) != (size_t)ncam * 12) return 1;
    fclose(f);
    {
        const int lv = getenv("LEVELS") ? atoi(getenv("LEVELS")) : 2, g0 = getenv("PHASE") ? atoi(getenv("PHASE")) : lv;
        for (int k = 0; k < 32; ++k) g_b[k] = g0 + k * lv;
        printf("tables: %d levels per step, root table resolves %d\n", lv, g0);
    }
    pthread_once(&g_h2f_once, init_h2f);
    orc_tree tree;
    memset(&tree, 0, sizeof(tree));
    tree.child = child; tree.data = sigma; tree.capacity = cap; tree.N = 2; tree.data_dim = 1;
    tree.format = ORC_FMT_SH; tree.basis_dim = 16; tree.ndc_width = -1.f;
    for (int i = 0; i < 3; ++i) { tree.offset[i] = meta[i]; tree.scale[i] = meta[3 + i]; }
    orc_options opt;
    orc_default_options(&opt);

    stats_t P0 = {0}, P1 = {0}, P2a[6], P2b[6], P3[6];
    memset(P3, 0, sizeof(P3));
    memset(P2a, 0, sizeof(P2a)); memset(P2b, 0, sizeof(P2b));
    const int thetas[6] = {4, 8, 12, 16, 24, 32};
    ray_t* rays = (ray_t*)malloc((size_t)W * H * sizeof(ray_t));
    double nsamp = 0, nshade = 0, nhit = 0, nmarch = 0;
    for (int c = 0; c < ncam; ++c) {
        orc_camera cam;
        cam.width = W; cam.height = H; cam.fx = meta[6]; cam.fy = meta[7];
        memcpy(cam.c2w, cams + 12 * c, 48);
        g_trace_n = 0;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) trace_one(&tree, &cam, &opt, x, y, &rays[(size_t)y * W + x]);
        for (size_t i = 0; i < (size_t)W * H; ++i) { nhit += rays[i].hit; nmarch += rays[i].n > 0; }
        {   /* ray-length distribution (single-frame tail bound) */
            static uint32_t hist[4096];
            memset(hist, 0, sizeof(hist));
            uint32_t mx = 0;
            for (size_t i = 0; i < (size_t)W * H; ++i) { uint32_t n = rays[i].n < 4095 ? rays[i].n : 4095; hist[n]++; if (n > mx) mx = n; }
            size_t acc = 0; const double qs[5] = {0.5, 0.9, 0.99, 0.999, 0.9999}; int qi = 0;
            printf("view %d samples/ray:", c);
            for (uint32_t n = 0; n <= mx && qi < 5; ++n) { acc += hist[n]; while (qi < 5 && acc >= qs[qi] * W * H) { printf(" p%g=%u", qs[qi] * 100, n); ++qi; } }
            printf(" max=%u\n", mx);
        }
        nsamp += g_trace_n;
        for (size_t i = 0; i < g_trace_n; ++i) nshade += g_trace[i] >> 7;
        sim_tiles(rays, W, H, 0, &P0);
        sim_tiles(rays, W, H, 1, &P1);
        for (int k = 0; k < 6; ++k) { g_p3_cur = k; sim_park(rays, W, H, g_p3_theta[k], g_p3_pat[k], &P3[k]); }
        for (int k = 0; k < 6; ++k) {
            sim_refill(rays, W, H, thetas[k], 64, 0, &P2a[k]);
            sim_refill(rays, W, H, thetas[k], 64, 1, &P2b[k]);
        }
    }
    printf("frames %d: samples %.2fM shaded %.3fM per frame, rays hit %.0f, rays with samples %.0f\n", ncam, nsamp / ncam / 1e6,
           nshade / ncam / 1e6, nhit / ncam, nmarch / ncam);
    {
        double tot = 0, cum = 0;
        for (int a = 1; a <= 32; ++a) tot += g_act_hist[a];
        printf("body iterations by active lanes (cumulative %%):");
        for (int a = 1; a <= 32; ++a) { cum += g_act_hist[a]; if (a % 4 == 0) printf(" <=%d:%.1f", a, 100 * cum / tot); }
        printf("\n");
    }
    report("P0 tiles, inline shading", &P0, ncam);
    report("P1 tiles + shared shade queue", &P1, ncam);
    for (int k = 0; k < 6; ++k) {
        char nm[64];
        snprintf(nm, sizeof nm, "P3 queue + tail parking th=%d pat=%d", g_p3_theta[k], g_p3_pat[k]);
        report(nm, &P3[k], ncam);
        printf("      park events %.0f/frame (%.1f rays each), pool passes %.0f/frame\n", g_park_events[k] / ncam, g_park_rays[k] / (g_park_events[k] + 1e-9), g_pool_passes[k] / ncam);
    }
    for (int k = 0; k < 6; ++k) {
        char nm[64];
        snprintf(nm, sizeof nm, "P2 refill theta=%d, inline", thetas[k]);
        report(nm, &P2a[k], ncam);
        snprintf(nm, sizeof nm, "P2 refill theta=%d + queue", thetas[k]);
        report(nm, &P2b[k], ncam);
    }
    return 0;
}
