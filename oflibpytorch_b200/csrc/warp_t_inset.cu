// W with an INSET flow: the flow (and its mask) cover an fh x fw window at (fy0, fx0) of the H x W frame of the
// payload and the output; outside the window the flow is zero and the result is not valid.
//
// Replaces Flow.apply(target, padding=[t, b, l, r], cut=False) for 't' flows (reference flow_class.py:905-934): there the
// flow is zero-padded to the target's size (Flow.pad, mode 'constant'), the whole frame is warped, and the valid area is
// (warped mask channel > 0.99999) & flow mask inside the window, False outside.  Here neither the padded flow nor the
// padded mask is materialised; the arithmetic per pixel is that of the plain warp (warp_t.cuh), so the outputs are
// bit-identical to warping with the padded flow.  This is a convenience path (one thread per pixel, plain REDs in the
// adjoint), not a tuned one.
#include "common.cuh"
#include "warp_t.cuh"

namespace ofb {

struct InsetParams {
    const float *flow; int64_t flow_bs;                 // (N|1, 2, fh, fw)
    const float *src; int64_t src_bs;                   // (N|1, C, H, W)
    const uint8_t *src_mask; int64_t src_mask_bs;       // (N|1, H, W) or NULL
    const uint8_t *flow_mask; int64_t flow_mask_bs;     // (N|1, fh, fw) or NULL
    float thresh;
    float *out; uint8_t *valid;                         // (N, C, H, W), (N, H, W) or NULL
    const float *gout; float *gflow; float *gsrc; int64_t gsrc_bs;   // adjoint: (N, C, H, W), (N, 2, fh, fw), (N|1, C, H, W)
    int N, C, H, W, fy0, fx0, fh, fw;
};

__global__ void __launch_bounds__(256) warp_t_inset_fwd_kernel(const InsetParams p, const CoordCtx cc)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W, FHW = p.fh * p.fw, pix = y * p.W + x;
    const bool in = (unsigned)(y - p.fy0) < (unsigned)p.fh && (unsigned)(x - p.fx0) < (unsigned)p.fw;
    const int fpix = (y - p.fy0) * p.fw + (x - p.fx0);
    float fx = 0.0f, fy = 0.0f;
    if (in) {
        const float *fl = p.flow + (int64_t)n * p.flow_bs;
        fx = __ldg(fl + fpix);
        fy = __ldg(fl + FHW + fpix);
    }
    float ix, iy;
    sample_pos(cc, fx, fy, (float)x, (float)y, ix, iy);
    const Window wn = identity_window(p.H, p.W);
    const PixGeom g = pix_geom<false>(ix, iy, wn);
    if (p.valid) {
        float a, b, c, d;
        if (p.src_mask) {
            const uint8_t *m = p.src_mask + (int64_t)n * p.src_mask_bs + g.o;
            uint8_t ua, ub, uc, ud;
            load_taps<false, uint8_t>(m, m + p.W, g, ua, ub, uc, ud);
            a = mask_to_float(ua); b = mask_to_float(ub); c = mask_to_float(uc); d = mask_to_float(ud);
        } else {
            a = g.pnw ? 1.0f : 0.0f; b = g.pne ? 1.0f : 0.0f; c = g.psw ? 1.0f : 0.0f; d = g.pse ? 1.0f : 0.0f;
        }
        bool ok = in && blend4(g, a, b, c, d) > p.thresh;                           // flow_class.py:922, :927-932
        if (ok && p.flow_mask) ok = __ldg(p.flow_mask + (int64_t)n * p.flow_mask_bs + fpix) != 0;
        p.valid[(int64_t)n * HW + pix] = ok ? 1 : 0;
    }
    const float *src = p.src + (int64_t)n * p.src_bs + g.o;
    float *out = p.out + (int64_t)n * p.C * HW + pix;
    for (int c = 0; c < p.C; ++c) {
        float a, b, cq, d;
        load_taps<false, float>(src + (int64_t)c * HW, src + (int64_t)c * HW + p.W, g, a, b, cq, d);
        out[(int64_t)c * HW] = blend4(g, a, b, cq, d);
    }
}

__global__ void __launch_bounds__(256) warp_t_inset_bwd_kernel(const InsetParams p, const CoordCtx cc)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W, FHW = p.fh * p.fw, pix = y * p.W + x;
    const bool in = (unsigned)(y - p.fy0) < (unsigned)p.fh && (unsigned)(x - p.fx0) < (unsigned)p.fw;
    const int fpix = (y - p.fy0) * p.fw + (x - p.fx0);
    float fx = 0.0f, fy = 0.0f;
    if (in) {
        const float *fl = p.flow + (int64_t)n * p.flow_bs;
        fx = __ldg(fl + fpix);
        fy = __ldg(fl + FHW + fpix);
    }
    float ix, iy;
    sample_pos(cc, fx, fy, (float)x, (float)y, ix, iy);
    const Window wn = identity_window(p.H, p.W);
    const PixGeom g = pix_geom<false>(ix, iy, wn);
    const float *go = p.gout + (int64_t)n * p.C * HW + pix;
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < p.C; ++c) {
        const float gv = __ldg(go + (int64_t)c * HW);
        if (p.gflow && in) {                          // the padded vectors are constants: no gradient outside the window
            const float *src = p.src + (int64_t)n * p.src_bs + (int64_t)c * HW + g.o;
            float a, b, cq, d;
            load_taps<false, float>(src, src + p.W, g, a, b, cq, d);
            gix = __fmaf_rn(gv, __fmaf_rn(b - a, g.sy, (d - cq) * g.ny), gix);
            giy = __fmaf_rn(gv, __fmaf_rn(cq - a, g.ex, (d - b) * g.wx), giy);
        }
        if (p.gsrc) {
            const float *t = p.gsrc + (int64_t)n * p.gsrc_bs + (int64_t)c * HW + g.o;
            red_add<0>(t, g.wnw * gv, g.pnw);
            red_add<4>(t, g.wne * gv, g.pne);
            red_add<0>(t + p.W, g.wsw * gv, g.psw);
            red_add<4>(t + p.W, g.wse * gv, g.pse);
        }
    }
    if (p.gflow && in) {
        float *gf = p.gflow + (int64_t)n * 2 * FHW + fpix;
        gf[0] = -gix;                                 // d(ix)/d(flow_x) = -1 (flow_sign = 1 on this path)
        gf[FHW] = -giy;
    }
}

}  // namespace ofb

using namespace ofb;

static int inset_check(int N, int C, int H, int W, const int *fwin)
{
    if (!fwin || N <= 0 || N > 65535 || C < 0) return OFB_EINVAL;
    if (H < 2 || W < 2) return OFB_ESHAPE;
    if (fwin[0] < 0 || fwin[1] < 0 || fwin[2] < 1 || fwin[3] < 1 || fwin[0] + fwin[2] > H || fwin[1] + fwin[3] > W)
        return OFB_EINVAL;
    if ((int64_t)H * W > 0x1fffffff) return OFB_EINVAL;
    return 0;
}

extern "C" int ofb_warp_t_inset_fwd(const float *flow, int64_t flow_bs,
                                    const float *src, int64_t src_bs,
                                    const uint8_t *src_mask, int64_t src_mask_bs,
                                    const uint8_t *flow_mask, int64_t flow_mask_bs,
                                    float valid_thresh, float *out, uint8_t *valid,
                                    int N, int C, int H, int W, const int *fwin, void *stream)
{
    const int st = inset_check(N, C, H, W, fwin);
    if (st) return st;
    if (!flow || (C > 0 && (!src || !out)) || (C == 0 && !valid)) return OFB_EINVAL;
    const InsetParams p{flow, flow_bs, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs, valid_thresh, out, valid,
                        nullptr, nullptr, nullptr, 0, N, C, H, W, fwin[0], fwin[1], fwin[2], fwin[3]};
    const CoordCtx cc{{(float)(W - 1), 1.0f / (float)(W - 1)}, {(float)(H - 1), 1.0f / (float)(H - 1)}, 1.0f};
    const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    warp_t_inset_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, cc);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_warp_t_inset_bwd(const float *flow, int64_t flow_bs,
                                    const float *src, int64_t src_bs,
                                    const float *gout, float *gflow, float *gsrc, int64_t gsrc_bs,
                                    int N, int C, int H, int W, const int *fwin, void *stream)
{
    const int st = inset_check(N, C, H, W, fwin);
    if (st) return st;
    if (!flow || !gout || C <= 0 || (gflow && !src)) return OFB_EINVAL;
    if (!gflow && !gsrc) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (gsrc) {
        const cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)(gsrc_bs == 0 ? 1 : N) * C * H * W, s);
        if (e != cudaSuccess) return (int)e;
    }
    const InsetParams p{flow, flow_bs, src, src_bs, nullptr, 0, nullptr, 0, 0.0f, nullptr, nullptr,
                        gout, gflow, gsrc, gsrc_bs, N, C, H, W, fwin[0], fwin[1], fwin[2], fwin[3]};
    const CoordCtx cc{{(float)(W - 1), 1.0f / (float)(W - 1)}, {(float)(H - 1), 1.0f / (float)(H - 1)}, 1.0f};
    const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    warp_t_inset_bwd_kernel<<<grid, 256, 0, s>>>(p, cc);
    OFB_CHECK_LAUNCH();
    return 0;
}
