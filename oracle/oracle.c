/*
 * oracle.c -- CPU restatement of the oflibpytorch dense flow-field hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (oflibpytorch_b200/)
 * may import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function
 * below against (a) golden vectors produced by the unmodified reference
 * (tests/golden/make_golden.py, run in the build container against
 * /root/reference/src) and (b) the reference's own known-answer tests
 * (test/test_utils.py, test/test_flow_class.py).
 *
 * The arithmetic follows, operation by operation and in fp32:
 *   W (backward warp, ref 't')  reference utils.py:541-555 (grid - flow,
 *       normalise_coords utils.py:462-465, grid_sample(bilinear, zeros,
 *       align_corners=True)) composed with ATen's un-normalisation
 *       (ATen/native/GridSampler.h:28-31) and the ATen CPU kernel's
 *       accumulation, which contracts to  fma(se,d, fma(sw,c, fma(ne,b, nw*a)))
 *       (established by bit-comparison, see DESIGN.md "Pinning the oracle").
 *   S (forward splat, ref 's')  reference utils.py:1045-1058 (endpoints),
 *       utils.py:1061-1154 (grid_from_unstructured_data), utils.py:1157-1205
 *       (apply_s_flow).  The four corner planes (y0,x0),(y0,x1),(y1,x0),(y1,x1)
 *       are each accumulated over source pixels in raster order and then summed
 *       ((P0+P1)+P2)+P3 -- the order torch's CPU scatter_add_/sum produce.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -mfma -fopenmp).
 * -ffp-contract=off matters: every fma below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_THRESHOLD 1e-3f /* utils.py:23 DEFAULT_THRESHOLD, compared in fp32 */
#define ORC_MIN_DENSITY 1e-3f /* utils.py:1144 clamp_min(density, 1e-3) */

/* ------------------------------------------------------------------------- */
/* W: backward warp                                                          */
/* ------------------------------------------------------------------------- */

/* Pixel coordinate -> normalised [-1,1] (utils.py:462-465) -> back to pixels
 * (GridSampler.h:31: ((coord + 1) / 2) * (size - 1)).  The round trip is NOT
 * the identity in fp32 and the validity masks depend on reproducing it. */
static inline float orc_round_trip(float c, float size_m1)
{
    float g = c * 2.0f;
    g = g / size_m1;
    g = g - 1.0f;
    return ((g + 1.0f) / 2.0f) * size_m1;
}

static inline int orc_inb(float x, float y, int H, int W)
{
    return x >= 0.0f && x <= (float)(W - 1) && y >= 0.0f && y <= (float)(H - 1);
}

/* flow (N,2,H,W); tgt (N,C,H,W); out (N,C,H,W).  The sample position is
 * p - sign*flow(p); sign = -1 restates the reference's apply_flow(-vecs, ...)
 * call sites (flow_class.py:1155,1166) without materialising -vecs. */
void orc_warp_t_fwd(const float *flow, float sign, const float *tgt, float *out,
                    int N, int C, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int i = 0; i < H; ++i) {
            const float *fx = flow + ((int64_t)n * 2 + 0) * HW + (int64_t)i * W;
            const float *fy = flow + ((int64_t)n * 2 + 1) * HW + (int64_t)i * W;
            for (int j = 0; j < W; ++j) {
                float cx = (float)j - sign * fx[j]; /* utils.py:548-550 */
                float cy = (float)i - sign * fy[j];
                float ix = orc_round_trip(cx, wm1);
                float iy = orc_round_trip(cy, hm1);
                float x0 = floorf(ix), y0 = floorf(iy);
                float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
                float nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy);
                float sw = (x1 - ix) * (iy - y0), se = (ix - x0) * (iy - y0);
                int b_nw = orc_inb(x0, y0, H, W), b_ne = orc_inb(x1, y0, H, W);
                int b_sw = orc_inb(x0, y1, H, W), b_se = orc_inb(x1, y1, H, W);
                int64_t o_nw = b_nw ? (int64_t)y0 * W + (int64_t)x0 : 0;
                int64_t o_ne = b_ne ? (int64_t)y0 * W + (int64_t)x1 : 0;
                int64_t o_sw = b_sw ? (int64_t)y1 * W + (int64_t)x0 : 0;
                int64_t o_se = b_se ? (int64_t)y1 * W + (int64_t)x1 : 0;
                for (int c = 0; c < C; ++c) {
                    const float *t = tgt + ((int64_t)n * C + c) * HW;
                    float a = b_nw ? t[o_nw] : 0.0f, b = b_ne ? t[o_ne] : 0.0f;
                    float cc = b_sw ? t[o_sw] : 0.0f, d = b_se ? t[o_se] : 0.0f;
                    float r = a * nw;
                    r = fmaf(b, ne, r);
                    r = fmaf(cc, sw, r);
                    r = fmaf(d, se, r);
                    out[((int64_t)n * C + c) * HW + (int64_t)i * W + j] = r;
                }
            }
        }
    }
}

/* Adjoint of orc_warp_t_fwd.  gflow (N,2,H,W) and gtgt (N,C,H,W) are
 * overwritten; either may be NULL.  gtgt accumulates in raster order of the
 * output pixels, corners nw,ne,sw,se (ATen grid_sampler_2d_backward, one
 * batch item per thread).  gflow replays the chain  grid_sample' -> x(W-1)/2
 * -> normalise_coords' (/(W-1), x2) -> minus. */
typedef struct { float nw, ne, sw, se, ex, wx, sy, ny; int64_t o_nw, o_ne, o_sw, o_se; int b_nw, b_ne, b_sw, b_se; } orc_taps;

static inline orc_taps orc_make_taps(const float *flow, float sign, int n, int i, int j, int H, int W)
{
    const int64_t HW = (int64_t)H * W, p = (int64_t)i * W + j;
    orc_taps t;
    float cx = (float)j - sign * flow[((int64_t)n * 2 + 0) * HW + p];
    float cy = (float)i - sign * flow[((int64_t)n * 2 + 1) * HW + p];
    float ix = orc_round_trip(cx, (float)(W - 1));
    float iy = orc_round_trip(cy, (float)(H - 1));
    float x0 = floorf(ix), y0 = floorf(iy);
    float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    t.ex = x1 - ix; t.wx = ix - x0; t.sy = y1 - iy; t.ny = iy - y0;
    t.nw = t.ex * t.sy; t.ne = t.wx * t.sy; t.sw = t.ex * t.ny; t.se = t.wx * t.ny;
    t.b_nw = orc_inb(x0, y0, H, W); t.b_ne = orc_inb(x1, y0, H, W);
    t.b_sw = orc_inb(x0, y1, H, W); t.b_se = orc_inb(x1, y1, H, W);
    t.o_nw = t.b_nw ? (int64_t)y0 * W + (int64_t)x0 : 0;
    t.o_ne = t.b_ne ? (int64_t)y0 * W + (int64_t)x1 : 0;
    t.o_sw = t.b_sw ? (int64_t)y1 * W + (int64_t)x0 : 0;
    t.o_se = t.b_se ? (int64_t)y1 * W + (int64_t)x1 : 0;
    return t;
}

void orc_warp_t_bwd(const float *flow, float sign, const float *tgt, const float *gout,
                    float *gflow, float *gtgt, int N, int C, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    if (gflow) {
        /* a gather: any order */
#pragma omp parallel for collapse(2) schedule(static)
        for (int n = 0; n < N; ++n)
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    int64_t p = (int64_t)i * W + j;
                    orc_taps t = orc_make_taps(flow, sign, n, i, j, H, W);
                    float gix = 0.0f, giy = 0.0f;
                    for (int c = 0; c < C; ++c) {
                        const float *s = tgt + ((int64_t)n * C + c) * HW;
                        float g = gout[((int64_t)n * C + c) * HW + p];
                        float a = t.b_nw ? s[t.o_nw] : 0.0f, b = t.b_ne ? s[t.o_ne] : 0.0f;
                        float cc = t.b_sw ? s[t.o_sw] : 0.0f, d = t.b_se ? s[t.o_se] : 0.0f;
                        gix += g * (-a * t.sy + b * t.sy - cc * t.ny + d * t.ny);
                        giy += g * (-a * t.ex - b * t.wx + cc * t.ex + d * t.wx);
                    }
                    float ggx = ((gix * (wm1 / 2.0f)) / wm1) * 2.0f;
                    float ggy = ((giy * (hm1 / 2.0f)) / hm1) * 2.0f;
                    gflow[((int64_t)n * 2 + 0) * HW + p] = -sign * ggx;
                    gflow[((int64_t)n * 2 + 1) * HW + p] = -sign * ggy;
                }
    }
    if (gtgt) {
        /* a scatter: every (n, c) plane accumulates in raster order of the output pixels, taps nw,ne,sw,se;
         * planes are independent, so they run in parallel without changing any sum */
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < C; ++c) {
                float *gs = gtgt + ((int64_t)n * C + c) * HW;
                const float *go = gout + ((int64_t)n * C + c) * HW;
                memset(gs, 0, sizeof(float) * (size_t)HW);
                for (int i = 0; i < H; ++i)
                    for (int j = 0; j < W; ++j) {
                        orc_taps t = orc_make_taps(flow, sign, n, i, j, H, W);
                        float g = go[(int64_t)i * W + j];
                        if (t.b_nw) gs[t.o_nw] += t.nw * g;
                        if (t.b_ne) gs[t.o_ne] += t.ne * g;
                        if (t.b_sw) gs[t.o_sw] += t.sw * g;
                        if (t.b_se) gs[t.o_se] += t.se * g;
                    }
            }
    }
}

/* ------------------------------------------------------------------------- */
/* S: forward splat                                                          */
/* ------------------------------------------------------------------------- */

typedef struct {
    float wx[2], wy[2];   /* utils.py:1110-1111, already times the in-bounds test */
    float inbx[2], inby[2];
    int64_t X[2], Y[2];   /* clamped corner coordinates, utils.py:1106-1107 */
} orc_corners;

static inline float orc_clampf(float v, float lo, float hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

static inline void orc_corner_setup(float x, float y, int H, int W, orc_corners *k)
{
    float x0 = floorf(x), y0 = floorf(y);
    float xx[2] = {x0, x0 + 1.0f}, yy[2] = {y0, y0 + 1.0f};
    float dx[2] = {xx[1] - x, x - xx[0]}, dy[2] = {yy[1] - y, y - yy[0]};
    for (int b = 0; b < 2; ++b) {
        float xs = orc_clampf(xx[b], 0.0f, (float)(W - 1));
        float ys = orc_clampf(yy[b], 0.0f, (float)(H - 1));
        k->inbx[b] = (xx[b] == xs) ? 1.0f : 0.0f;
        k->inby[b] = (yy[b] == ys) ? 1.0f : 0.0f;
        k->wx[b] = dx[b] * k->inbx[b];
        k->wy[b] = dy[b] * k->inby[b];
        k->X[b] = (int64_t)xs;
        k->Y[b] = (int64_t)ys;
    }
}

/* Accumulate one batch item: planes[k][ (C+1) ][HW], k = 2a+b, the last
 * "channel" is the density.  fm may be NULL (all true). */
static void orc_splat_item(const float *x, const float *y, const float *data,
                           const uint8_t *fm, int C, int H, int W, float *planes)
{
    const int64_t HW = (int64_t)H * W;
    memset(planes, 0, sizeof(float) * 4 * (size_t)(C + 1) * HW);
    /* One raster pass feeding four separate planes keeps every plane's own
     * accumulation in raster order, which is all the reference's order needs. */
    for (int64_t s = 0; s < HW; ++s) {
        orc_corners k;
        orc_corner_setup(x[s], y[s], H, W, &k);
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                float *P = planes + (int64_t)(2 * a + b) * (C + 1) * HW;
                float w = k.wy[a] * k.wx[b];              /* utils.py:1114 */
                if (fm) w = w * (float)(fm[s] ? 1 : 0);   /* utils.py:1123 */
                int64_t p = (int64_t)W * k.Y[a] + k.X[b]; /* utils.py:1118 */
                P[(int64_t)C * HW + p] += w;              /* utils.py:1133 */
                for (int c = 0; c < C; ++c)               /* utils.py:1139-1140 */
                    P[(int64_t)c * HW + p] += w * data[(int64_t)c * HW + s];
            }
    }
}

static void orc_planes_sum(const float *planes, int C, int64_t HW, float *num, float *den)
{
    const int64_t S = (int64_t)(C + 1) * HW;
    for (int c = 0; c <= C; ++c) {
        float *dst = (c < C) ? num + (int64_t)c * HW : den;
        const float *p0 = planes + (int64_t)c * HW;
        for (int64_t p = 0; p < HW; ++p)
            dst[p] = ((p0[p] + p0[S + p]) + p0[2 * S + p]) + p0[3 * S + p]; /* utils.py:1134,1144 */
    }
}

/* grid_from_unstructured_data (utils.py:1061-1154).
 * x,y (N,H,W); data (N,C,H,W); mask (N,H,W) bytes or NULL -> out (N,C,H,W), density (N,H,W) */
void orc_grid_from_unstructured(const float *x, const float *y, const float *data,
                                const uint8_t *mask, float *out, float *density,
                                int N, int C, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; ++n) {
        float *planes = (float *)malloc(sizeof(float) * 4 * (size_t)(C + 1) * HW);
        float *num = (float *)malloc(sizeof(float) * (size_t)C * HW);
        float *den = density + (int64_t)n * HW;
        orc_splat_item(x + (int64_t)n * HW, y + (int64_t)n * HW, data + (int64_t)n * C * HW,
                       mask ? mask + (int64_t)n * HW : NULL, C, H, W, planes);
        orc_planes_sum(planes, C, HW, num, den);
        for (int c = 0; c < C; ++c)
            for (int64_t p = 0; p < HW; ++p)
                out[((int64_t)n * C + c) * HW + p] =
                    num[(int64_t)c * HW + p] / fmaxf(den[p], ORC_MIN_DENSITY);
        free(planes);
        free(num);
    }
}

/* Adjoint of orc_grid_from_unstructured with upstream gradients gout (N,C,H,W)
 * and gdensity (N,H,W, may be NULL) -> gx, gy (N,H,W), gdata (N,C,H,W).
 * If zero_at (N,H,W bytes, may be NULL) is given, gout is treated as zero at
 * those target pixels (the un-occlude overwrite of apply_s_flow, utils.py:1203). */
void orc_grid_from_unstructured_bwd(const float *x, const float *y, const float *data,
                                    const uint8_t *mask, const float *gout,
                                    const float *gdensity, const uint8_t *zero_at,
                                    float *gx, float *gy, float *gdata,
                                    int N, int C, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; ++n) {
        const float *xn = x + (int64_t)n * HW, *yn = y + (int64_t)n * HW;
        const float *dn = data + (int64_t)n * C * HW;
        const uint8_t *mn = mask ? mask + (int64_t)n * HW : NULL;
        float *planes = (float *)malloc(sizeof(float) * 4 * (size_t)(C + 1) * HW);
        float *num = (float *)malloc(sizeof(float) * (size_t)C * HW);
        float *den = (float *)malloc(sizeof(float) * (size_t)HW);
        float *gnum = (float *)malloc(sizeof(float) * (size_t)C * HW);
        float *gden = (float *)malloc(sizeof(float) * (size_t)HW);
        float *gd_c = (float *)malloc(sizeof(float) * (size_t)(C > 0 ? C : 1));
        orc_splat_item(xn, yn, dn, mn, C, H, W, planes);
        orc_planes_sum(planes, C, HW, num, den);
        for (int64_t p = 0; p < HW; ++p) {
            float dc = fmaxf(den[p], ORC_MIN_DENSITY);
            int dead = zero_at && zero_at[(int64_t)n * HW + p];
            float acc = 0.0f;
            for (int c = 0; c < C; ++c) {
                float g = dead ? 0.0f : gout[((int64_t)n * C + c) * HW + p];
                gnum[(int64_t)c * HW + p] = g / dc;
                acc += g * num[(int64_t)c * HW + p] / (dc * dc);
            }
            float gd = gdensity ? gdensity[(int64_t)n * HW + p] : 0.0f;
            gden[p] = gd + ((den[p] >= ORC_MIN_DENSITY) ? -acc : 0.0f);
        }
        for (int64_t s = 0; s < HW; ++s) {
            orc_corners k;
            orc_corner_setup(xn[s], yn[s], H, W, &k);
            float fm = (mn && !mn[s]) ? 0.0f : 1.0f;
            float G[2][2];
            for (int c = 0; c < C; ++c) gd_c[c] = 0.0f;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    int64_t p = (int64_t)W * k.Y[a] + k.X[b];
                    float w = k.wy[a] * k.wx[b] * fm;
                    float g = gden[p];
                    for (int c = 0; c < C; ++c) {
                        float gn = gnum[(int64_t)c * HW + p];
                        g += gn * dn[(int64_t)c * HW + s];
                        gd_c[c] += w * gn;
                    }
                    G[a][b] = g;
                }
            for (int c = 0; c < C; ++c) gdata[((int64_t)n * C + c) * HW + s] = gd_c[c];
            float sx = 0.0f, sy = 0.0f;
            for (int a = 0; a < 2; ++a) sx += k.wy[a] * (k.inbx[1] * G[a][1] - k.inbx[0] * G[a][0]);
            for (int b = 0; b < 2; ++b) sy += k.wx[b] * (k.inby[1] * G[1][b] - k.inby[0] * G[0][b]);
            gx[(int64_t)n * HW + s] = fm * sx;
            gy[(int64_t)n * HW + s] = fm * sy;
        }
        free(planes); free(num); free(den); free(gnum); free(gden); free(gd_c);
    }
}

/* get_flow_endpoints (utils.py:1045-1058): x = s*f0 + j, y = s*f1 + i. */
void orc_flow_endpoints(const float *flow, float s, float *x, float *y, int N, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                int64_t p = (int64_t)i * W + j;
                x[(int64_t)n * HW + p] = s * flow[((int64_t)n * 2 + 0) * HW + p] + (float)j;
                y[(int64_t)n * HW + p] = s * flow[((int64_t)n * 2 + 1) * HW + p] + (float)i;
            }
}

/* zero_mask of apply_s_flow (utils.py:1191 via threshold_vectors utils.py:642):
 * both components strictly inside (-1e-3, 1e-3). */
void orc_zero_mask(const float *flow, uint8_t *zero, int N, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
    for (int n = 0; n < N; ++n)
        for (int64_t p = 0; p < HW; ++p) {
            float a = flow[((int64_t)n * 2 + 0) * HW + p], b = flow[((int64_t)n * 2 + 1) * HW + p];
            zero[(int64_t)n * HW + p] =
                (a < ORC_THRESHOLD && a > -ORC_THRESHOLD && b < ORC_THRESHOLD && b > -ORC_THRESHOLD);
        }
}

/* is_zero_flow (utils.py:919-938): per batch item, all components zero
 * (after thresholding when thresholded != 0). */
void orc_is_zero_flow(const float *flow, int thresholded, uint8_t *out, int N, int H, int W)
{
    const int64_t n2 = 2 * (int64_t)H * W;
    for (int n = 0; n < N; ++n) {
        int all = 1;
        for (int64_t q = 0; q < n2 && all; ++q) {
            float v = flow[(int64_t)n * n2 + q];
            if (thresholded) all = (v < ORC_THRESHOLD && v > -ORC_THRESHOLD);
            else all = (v == 0.0f);
        }
        out[n] = (uint8_t)all;
    }
}

/* ------------------------------------------------------------------------- */
/* point sampling (track_pts, utils.py:1003-1019)                            */
/* ------------------------------------------------------------------------- */

/* out[n,m,:] = (flow_y, flow_x) interpolated at pts[n,m,:] = (y, x): grid_sample(flow, normalise_coords(pts flipped),
 * align_corners=True) followed by .flip(1) (utils.py:1010-1018), with ATen's accumulation order. */
void orc_sample_pts_fwd(const float *flow, const float *pts, float *out, int N, int M, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < M; ++m) {
            const float *pp = pts + ((int64_t)n * M + m) * 2;
            float ix = orc_round_trip(pp[1], (float)(W - 1)), iy = orc_round_trip(pp[0], (float)(H - 1));
            float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
            float nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy), sw = (x1 - ix) * (iy - y0), se = (ix - x0) * (iy - y0);
            int b_nw = orc_inb(x0, y0, H, W), b_ne = orc_inb(x1, y0, H, W), b_sw = orc_inb(x0, y1, H, W), b_se = orc_inb(x1, y1, H, W);
            for (int c = 0; c < 2; ++c) {
                const float *t = flow + ((int64_t)n * 2 + c) * HW;
                float a = b_nw ? t[(int64_t)y0 * W + (int64_t)x0] : 0.0f, b = b_ne ? t[(int64_t)y0 * W + (int64_t)x1] : 0.0f;
                float cc = b_sw ? t[(int64_t)y1 * W + (int64_t)x0] : 0.0f, d = b_se ? t[(int64_t)y1 * W + (int64_t)x1] : 0.0f;
                float r = a * nw;
                r = fmaf(b, ne, r);
                r = fmaf(cc, sw, r);
                r = fmaf(d, se, r);
                out[((int64_t)n * M + m) * 2 + (1 - c)] = r;
            }
        }
}

/* adjoint: gflow (N,2,H,W) and gpts (N,M,2, (y,x) order) are overwritten */
void orc_sample_pts_bwd(const float *flow, const float *pts, const float *gout, float *gflow, float *gpts,
                        int N, int M, int H, int W)
{
    const int64_t HW = (int64_t)H * W;
    memset(gflow, 0, sizeof(float) * (size_t)N * 2 * HW);
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < M; ++m) {
            const float *pp = pts + ((int64_t)n * M + m) * 2;
            float ix = orc_round_trip(pp[1], (float)(W - 1)), iy = orc_round_trip(pp[0], (float)(H - 1));
            float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
            float ex = x1 - ix, wx = ix - x0, sy = y1 - iy, ny = iy - y0;
            int b_nw = orc_inb(x0, y0, H, W), b_ne = orc_inb(x1, y0, H, W), b_sw = orc_inb(x0, y1, H, W), b_se = orc_inb(x1, y1, H, W);
            float gx = 0.0f, gy = 0.0f;
            for (int c = 0; c < 2; ++c) {
                float g = gout[((int64_t)n * M + m) * 2 + (1 - c)];
                const float *t = flow + ((int64_t)n * 2 + c) * HW;
                float *gt = gflow + ((int64_t)n * 2 + c) * HW;
                float a = 0, b = 0, cc = 0, d = 0;
                if (b_nw) { a = t[(int64_t)y0 * W + (int64_t)x0]; gt[(int64_t)y0 * W + (int64_t)x0] += ex * sy * g; }
                if (b_ne) { b = t[(int64_t)y0 * W + (int64_t)x1]; gt[(int64_t)y0 * W + (int64_t)x1] += wx * sy * g; }
                if (b_sw) { cc = t[(int64_t)y1 * W + (int64_t)x0]; gt[(int64_t)y1 * W + (int64_t)x0] += ex * ny * g; }
                if (b_se) { d = t[(int64_t)y1 * W + (int64_t)x1]; gt[(int64_t)y1 * W + (int64_t)x1] += wx * ny * g; }
                gx += g * ((b - a) * sy + (d - cc) * ny);
                gy += g * ((cc - a) * ex + (d - b) * wx);
            }
            gpts[((int64_t)n * M + m) * 2] = gy;
            gpts[((int64_t)n * M + m) * 2 + 1] = gx;
        }
}
