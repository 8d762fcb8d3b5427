// S: normalised bilinear forward splat ('s' flows) and its adjoint.
//
// Replaces reference utils.py:1045-1058 (get_flow_endpoints), utils.py:1061-1154
// (grid_from_unstructured_data), utils.py:1157-1205 (apply_s_flow) and the 's' branch of the mask
// tracking in Flow.apply (flow_class.py:880-898,922).
//
// Mask bits without ordered sums.  The reference decides validity by comparing two sums accumulated
// in the same order (numerator of the mask channel and density).  Atomics reorder sums, so the kernel
// carries an order-independent side channel instead: bad = sum of w over contributors whose mask
// channel is 0.  "mask channel == 1"  <=>  bad == 0 (and density >= 1e-3); both tests, and
// density > 0, are sums of non-negative terms compared against zero and therefore independent of the
// accumulation order (probe C9).
#include "common.cuh"

namespace ofb {

struct Corner2 {
    float wx[2], wy[2];
    float inbx[2], inby[2];
    int X[2], Y[2];
};

// utils.py:1098-1111: floor, the two integer neighbours, clamp, weights times the in-bounds test
__device__ __forceinline__ Corner2 make_corners(float x, float y, int H, int W)
{
    Corner2 k;
    const float x0 = floorf(x), y0 = floorf(y);
    const float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);
    const float wm = (float)(W - 1), hm = (float)(H - 1);
    const float xs0 = fminf(fmaxf(x0, 0.0f), wm), xs1 = fminf(fmaxf(x1, 0.0f), wm);
    const float ys0 = fminf(fmaxf(y0, 0.0f), hm), ys1 = fminf(fmaxf(y1, 0.0f), hm);
    k.inbx[0] = (x0 == xs0) ? 1.0f : 0.0f;
    k.inbx[1] = (x1 == xs1) ? 1.0f : 0.0f;
    k.inby[0] = (y0 == ys0) ? 1.0f : 0.0f;
    k.inby[1] = (y1 == ys1) ? 1.0f : 0.0f;
    k.wx[0] = __fmul_rn(__fsub_rn(x1, x), k.inbx[0]);
    k.wx[1] = __fmul_rn(__fsub_rn(x, x0), k.inbx[1]);
    k.wy[0] = __fmul_rn(__fsub_rn(y1, y), k.inby[0]);
    k.wy[1] = __fmul_rn(__fsub_rn(y, y0), k.inby[1]);
    // NaN coordinates fail every comparison above; keep the indices in range regardless
    k.X[0] = (k.inbx[0] != 0.0f) ? (int)xs0 : 0;
    k.X[1] = (k.inbx[1] != 0.0f) ? (int)xs1 : 0;
    k.Y[0] = (k.inby[0] != 0.0f) ? (int)ys0 : 0;
    k.Y[1] = (k.inby[1] != 0.0f) ? (int)ys1 : 0;
    return k;
}

struct SplatParams {
    const float *cx, *cy; int64_t coord_bs; int relative; float sign;
    const float *data; int64_t data_bs;
    const uint8_t *data_mask; int64_t data_mask_bs;
    const uint8_t *flow_mask; int64_t flow_mask_bs;
    int occlude_zero;
    float *acc;
    float *out, *density; uint8_t *warped_mask, *unocc, *valid; int valid_mode;
    int N, C, H, W;
};

__device__ __forceinline__ void endpoint(const SplatParams &p, int n, int pix, int x, int y,
                                         float &X, float &Y, bool &zero)
{
    const float fx = __ldg(p.cx + (int64_t)n * p.coord_bs + pix);
    const float fy = __ldg(p.cy + (int64_t)n * p.coord_bs + pix);
    if (p.relative) {
        X = __fadd_rn(__fmul_rn(p.sign, fx), (float)x);   // utils.py:1056-1057
        Y = __fadd_rn(__fmul_rn(p.sign, fy), (float)y);
        zero = p.occlude_zero && is_zero_vec(fx, fy);     // utils.py:1191
    } else {
        X = fx;
        Y = fy;
        zero = false;
    }
}

// v0 accumulation: one thread per source pixel, one red.global.add.f32 per (tap, channel).
__global__ void __launch_bounds__(256) splat_accumulate_kernel(const SplatParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W;
    const int pix = y * p.W + x;
    float X, Y;
    bool zero;
    endpoint(p, n, pix, x, y, X, Y, zero);
    bool fm = !zero;                                                         // utils.py:1192
    if (p.flow_mask) fm = fm && (__ldg(p.flow_mask + (int64_t)n * p.flow_mask_bs + pix) != 0);
    if (!fm) return;                                                         // weight * 0, utils.py:1123
    const Corner2 k = make_corners(X, Y, p.H, p.W);
    const bool is_bad = p.data_mask && (__ldg(p.data_mask + (int64_t)n * p.data_mask_bs + pix) == 0);
    float *acc = p.acc + (int64_t)n * (p.C + 2) * HW;
    const float *data = p.data + (int64_t)n * p.data_bs + pix;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const float w = __fmul_rn(k.wy[a], k.wx[b]);                     // utils.py:1114
            if (!(w > 0.0f)) continue;                                       // adding +0 changes nothing
            const int q = p.W * k.Y[a] + k.X[b];                             // utils.py:1118
            atomicAdd(acc + (int64_t)p.C * HW + q, w);                       // utils.py:1133
            if (is_bad) atomicAdd(acc + (int64_t)(p.C + 1) * HW + q, w);
            for (int c = 0; c < p.C; ++c)                                    // utils.py:1139-1140
                atomicAdd(acc + (int64_t)c * HW + q, __fmul_rn(w, __ldg(data + (int64_t)c * HW)));
        }
}

// normalise by max(density, 1e-3) (utils.py:1144), density > 0 (utils.py:1197), un-occlude fill
// (utils.py:1202-1203) and the validity bit (flow_class.py:922 / :1117,1170), one pass over the targets
__global__ void __launch_bounds__(256) splat_normalise_kernel(const SplatParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W;
    const int pix = y * p.W + x;
    const float *acc = p.acc + (int64_t)n * (p.C + 2) * HW + pix;
    const float den = acc[(int64_t)p.C * HW];
    const float bad = p.data_mask ? acc[(int64_t)(p.C + 1) * HW] : 0.0f;
    const float dc = fmaxf(den, OFB_MIN_DENSITY);
    const bool wm = den > 0.0f;
    bool unocc = false;
    if (p.occlude_zero && p.relative && !wm) {
        const float fx = __ldg(p.cx + (int64_t)n * p.coord_bs + pix);
        const float fy = __ldg(p.cy + (int64_t)n * p.coord_bs + pix);
        unocc = is_zero_vec(fx, fy);
        if (p.flow_mask) unocc = unocc && (__ldg(p.flow_mask + (int64_t)n * p.flow_mask_bs + pix) != 0);
    }
    const float *data = p.data + (int64_t)n * p.data_bs + pix;
    float *out = p.out + (int64_t)n * p.C * HW + pix;
    for (int c = 0; c < p.C; ++c) {
        const float v = unocc ? __ldg(data + (int64_t)c * HW) : __fdiv_rn(acc[(int64_t)c * HW], dc);
        out[(int64_t)c * HW] = v;
    }
    if (p.density) p.density[(int64_t)n * HW + pix] = den;
    if (p.warped_mask) p.warped_mask[(int64_t)n * HW + pix] = wm ? 1 : 0;
    if (p.unocc) p.unocc[(int64_t)n * HW + pix] = unocc ? 1 : 0;
    if (p.valid) {
        float mch;
        if (unocc) {
            mch = (p.data_mask && __ldg(p.data_mask + (int64_t)n * p.data_mask_bs + pix) == 0) ? 0.0f : 1.0f;
        } else {
            const float mnum = (bad == 0.0f) ? den : __fsub_rn(den, bad);
            mch = __fdiv_rn(mnum, dc);
        }
        const bool ok = (p.valid_mode == OFB_VALID_EQ1) ? (mch == 1.0f) : (mch > 0.99999f);
        p.valid[(int64_t)n * HW + pix] = ok ? 1 : 0;
    }
}

struct SplatBwdParams {
    const float *cx, *cy; int64_t coord_bs; int relative; float sign;
    const float *data; int64_t data_bs;
    const uint8_t *flow_mask; int64_t flow_mask_bs; int occlude_zero;
    const float *out, *density; const uint8_t *unocc;
    const float *gout, *gdensity;
    float *gcx, *gcy; int64_t gcoord_bs; float *gdata;
    int N, C, H, W;
};

constexpr int kMaxC = 8;   // channels held in registers by the v0 adjoint; larger C is chunked

// v0 adjoint: a pure gather per source pixel.
__global__ void __launch_bounds__(256) splat_bwd_kernel(const SplatBwdParams p, const int c0, const int cn,
                                                        const bool first_chunk)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W;
    const int pix = y * p.W + x;
    const float fx = __ldg(p.cx + (int64_t)n * p.coord_bs + pix);
    const float fy = __ldg(p.cy + (int64_t)n * p.coord_bs + pix);
    float X = fx, Y = fy;
    bool zero = false;
    if (p.relative) {
        X = __fadd_rn(__fmul_rn(p.sign, fx), (float)x);
        Y = __fadd_rn(__fmul_rn(p.sign, fy), (float)y);
        zero = p.occlude_zero && is_zero_vec(fx, fy);
    }
    bool fm = !zero;
    if (p.flow_mask) fm = fm && (__ldg(p.flow_mask + (int64_t)n * p.flow_mask_bs + pix) != 0);

    const float *gout = p.gout + (int64_t)n * p.C * HW;
    const float *outp = p.out + (int64_t)n * p.C * HW;
    const float *den = p.density + (int64_t)n * HW;
    const uint8_t *uno = p.unocc ? p.unocc + (int64_t)n * HW : nullptr;
    const float *data = p.data + (int64_t)n * p.data_bs + pix;

    float gd[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) gd[c] = 0.0f;
    float gx = 0.0f, gy = 0.0f;

    if (fm) {
        const Corner2 k = make_corners(X, Y, p.H, p.W);
        float dat[kMaxC];
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) dat[c] = (c < cn) ? __ldg(data + (int64_t)(c0 + c) * HW) : 0.0f;
        float G[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                G[a][b] = 0.0f;
                if (k.inbx[b] == 0.0f || k.inby[a] == 0.0f) continue;   // contributes to nothing
                const int q = p.W * k.Y[a] + k.X[b];
                const float dq = __ldg(den + q);
                const float dc = fmaxf(dq, OFB_MIN_DENSITY);
                const bool dead = uno && (__ldg(uno + q) != 0);          // utils.py:1203 overwrote it
                const float w = __fmul_rn(k.wy[a], k.wx[b]);
                float g = 0.0f;
                // density path: every channel of the quotient contributes, so it is evaluated over ALL
                // channels in the first chunk only
                if (first_chunk) {
                    float gden = p.gdensity ? __ldg(p.gdensity + (int64_t)n * HW + q) : 0.0f;
                    if (!dead && dq >= OFB_MIN_DENSITY) {
                        float s = 0.0f;
                        for (int c = 0; c < p.C; ++c)
                            s += __ldg(gout + (int64_t)c * HW + q) * __ldg(outp + (int64_t)c * HW + q);
                        gden -= s / dc;
                    }
                    g = gden;
                }
                if (!dead) {
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c) {
                        if (c < cn) {
                            const float gnum = __ldg(gout + (int64_t)(c0 + c) * HW + q) / dc;
                            g += gnum * dat[c];
                            gd[c] += w * gnum;
                        }
                    }
                }
                G[a][b] = g;
            }
        // d w / d x: wx[0] = (x1 - x)*inbx[0] -> -inbx[0];  wx[1] = (x - x0)*inbx[1] -> +inbx[1]
        gx = k.wy[0] * (k.inbx[1] * G[0][1] - k.inbx[0] * G[0][0]) +
             k.wy[1] * (k.inbx[1] * G[1][1] - k.inbx[0] * G[1][0]);
        gy = k.wx[0] * (k.inby[1] * G[1][0] - k.inby[0] * G[0][0]) +
             k.wx[1] * (k.inby[1] * G[1][1] - k.inby[0] * G[0][1]);
        if (p.relative) {
            gx *= p.sign;
            gy *= p.sign;
        }
    }
    if (p.gdata) {
        const bool self_unocc = uno && (__ldg(uno + pix) != 0);
        float *gdp = p.gdata + (int64_t)n * p.C * HW + pix;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c)
            if (c < cn) {
                float v = gd[c];
                if (self_unocc) v += __ldg(gout + (int64_t)(c0 + c) * HW + pix);
                gdp[(int64_t)(c0 + c) * HW] = v;
            }
    }
    if (p.gcx) {
        float *px = p.gcx + (int64_t)n * p.gcoord_bs + pix;
        float *py = p.gcy + (int64_t)n * p.gcoord_bs + pix;
        if (first_chunk) {
            *px = gx;
            *py = gy;
        } else {
            *px += gx;
            *py += gy;
        }
    }
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_splat_s_fwd(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                               const float *data, int64_t data_bs,
                               const uint8_t *data_mask, int64_t data_mask_bs,
                               const uint8_t *flow_mask, int64_t flow_mask_bs,
                               int occlude_zero, float *acc,
                               float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                               uint8_t *valid, int valid_mode,
                               int N, int C, int H, int W, void *stream)
{
    if (!cx || !cy || !acc || N <= 0 || C < 0 || (C > 0 && (!data || !out))) return OFB_EINVAL;
    if (occlude_zero && !relative) return OFB_EINVAL;
    if (valid && valid_mode != OFB_VALID_GT && valid_mode != OFB_VALID_EQ1) return OFB_EINVAL;
    if (H < 1 || W < 1) return OFB_ESHAPE;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)N * (C + 2) * H * W, st);
    if (e != cudaSuccess) return (int)e;
    SplatParams p{cx, cy, coord_bs, relative, sign, data, data_bs, data_mask, data_mask_bs,
                  flow_mask, flow_mask_bs, occlude_zero, acc, out, density, warped_mask, unocc,
                  valid, valid_mode, N, C, H, W};
    const dim3 block(32, 8);
    const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    splat_accumulate_kernel<<<grid, block, 0, st>>>(p);
    OFB_CHECK_LAUNCH();
    splat_normalise_kernel<<<grid, block, 0, st>>>(p);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_splat_s_bwd(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                               const float *data, int64_t data_bs,
                               const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                               const float *out, const float *density, const uint8_t *unocc,
                               const float *gout, const float *gdensity,
                               float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                               int N, int C, int H, int W, void *stream)
{
    if (!cx || !cy || !density || N <= 0 || C < 0) return OFB_EINVAL;
    if (C > 0 && (!data || !out || !gout)) return OFB_EINVAL;
    if ((gcx == nullptr) != (gcy == nullptr)) return OFB_EINVAL;
    if (H < 1 || W < 1) return OFB_ESHAPE;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    if (!gcx && !gdata) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplatBwdParams p{cx, cy, coord_bs, relative, sign, data, data_bs, flow_mask, flow_mask_bs, occlude_zero,
                     out, density, unocc, gout, gdensity, gcx, gcy, gcoord_bs, gdata, N, C, H, W};
    const dim3 block(32, 8);
    const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    int c0 = 0;
    do {
        const int cn = (C - c0 < kMaxC) ? (C - c0) : kMaxC;
        splat_bwd_kernel<<<grid, block, 0, st>>>(p, c0, cn, c0 == 0);
        OFB_CHECK_LAUNCH();
        c0 += kMaxC;
    } while (c0 < C);
    return 0;
}
