// W: backward warp ('t' flows) with fused validity, and its adjoint.
//
// Replaces reference utils.py:541-555 (apply_flow ref 't') and the mask tracking of Flow.apply
// (flow_class.py:896-898,922-934).  One thread owns VEC horizontally adjacent output pixels: the two
// flow planes are read with one 128-bit load each, every payload channel is written with one 128-bit
// store, and the 4*VEC bilinear taps per channel go through the read-only path (neighbouring threads
// hit the same 128 B lines, so the source is fetched from HBM once).
#include "common.cuh"

namespace ofb {

struct WarpFwdParams {
    const float *flow; int64_t flow_bs; float sign;
    const float *src; int64_t src_bs;
    const uint8_t *src_mask; int64_t src_mask_bs;
    const uint8_t *flow_mask; int64_t flow_mask_bs;
    float add_flow, thresh;
    float *out; uint8_t *valid;
    int N, C, H, W;
};

template <int VEC> struct FVec;
template <> struct FVec<1> { float v[1]; };
template <> struct FVec<4> { float v[4]; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float *__restrict__ p, float (&v)[VEC])
{
    if constexpr (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float *__restrict__ p, const float (&v)[VEC])
{
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        p[0] = v[0];
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) warp_t_fwd_kernel(const WarpFwdParams p)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W;
    const int pix = y * p.W + x;
    const float wm1 = (float)(p.W - 1), hm1 = (float)(p.H - 1);

    float fx[VEC], fy[VEC];
    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    load_vec<VEC>(fl + pix, fx);
    load_vec<VEC>(fl + HW + pix, fy);

    Taps t[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float cx = __fsub_rn((float)(x + k), __fmul_rn(p.sign, fx[k]));   // utils.py:548-550
        const float cy = __fsub_rn((float)y, __fmul_rn(p.sign, fy[k]));
        t[k] = make_taps(round_trip(cx, wm1), round_trip(cy, hm1), p.H, p.W);
    }

    if (p.valid) {
        const uint8_t *sm = p.src_mask ? p.src_mask + (int64_t)n * p.src_mask_bs : nullptr;
        const uint8_t *fm = p.flow_mask ? p.flow_mask + (int64_t)n * p.flow_mask_bs + pix : nullptr;
        uint8_t v[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float a, b, c, d;
            if (sm) {
                a = t[k].o_nw >= 0 ? (float)__ldg(sm + t[k].o_nw) : 0.0f;
                b = t[k].o_ne >= 0 ? (float)__ldg(sm + t[k].o_ne) : 0.0f;
                c = t[k].o_sw >= 0 ? (float)__ldg(sm + t[k].o_sw) : 0.0f;
                d = t[k].o_se >= 0 ? (float)__ldg(sm + t[k].o_se) : 0.0f;
            } else {
                a = t[k].o_nw >= 0 ? 1.0f : 0.0f;
                b = t[k].o_ne >= 0 ? 1.0f : 0.0f;
                c = t[k].o_sw >= 0 ? 1.0f : 0.0f;
                d = t[k].o_se >= 0 ? 1.0f : 0.0f;
            }
            const float mch = blend(t[k], a, b, c, d);                     // flow_class.py:898
            bool ok = mch > p.thresh;                                       // flow_class.py:922
            if (fm) ok = ok && (fm[k] != 0);                                // flow_class.py:934
            v[k] = ok ? 1 : 0;
        }
        uint8_t *vp = p.valid + (int64_t)n * HW + pix;
        if constexpr (VEC == 4) {
            *reinterpret_cast<uchar4 *>(vp) = make_uchar4(v[0], v[1], v[2], v[3]);
        } else {
            vp[0] = v[0];
        }
    }

    const float *src = p.src + (int64_t)n * p.src_bs;
    float *out = p.out + (int64_t)n * p.C * HW + pix;
#pragma unroll 1
    for (int c = 0; c < p.C; ++c) {
        const float *s = src + (int64_t)c * HW;
        float r[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float a = ldg_or_zero(s, t[k].o_nw), b = ldg_or_zero(s, t[k].o_ne);
            const float cc = ldg_or_zero(s, t[k].o_sw), d = ldg_or_zero(s, t[k].o_se);
            r[k] = blend(t[k], a, b, cc, d);
        }
        if (p.add_flow != 0.0f && c < 2) {                                  // flow_class.py:1804,1808
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                r[k] = __fadd_rn(__fmul_rn(p.add_flow, c == 0 ? fx[k] : fy[k]), r[k]);
        }
        store_vec<VEC>(out + (int64_t)c * HW, r);
    }
}

struct WarpBwdParams {
    const float *flow; int64_t flow_bs; float sign;
    const float *src; int64_t src_bs;
    const float *gout; float add_flow;
    float *gflow; float *gsrc; int64_t gsrc_bs;
    int N, C, H, W;
};

// v0 adjoint: gflow is a gather; gsrc is scattered with one red.global.add.f32 per tap.
template <int VEC>
__global__ void __launch_bounds__(256) warp_t_bwd_kernel(const WarpBwdParams p)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W;
    const int pix = y * p.W + x;
    const float wm1 = (float)(p.W - 1), hm1 = (float)(p.H - 1);

    float fx[VEC], fy[VEC];
    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    load_vec<VEC>(fl + pix, fx);
    load_vec<VEC>(fl + HW + pix, fy);
    Taps t[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float cx = __fsub_rn((float)(x + k), __fmul_rn(p.sign, fx[k]));
        const float cy = __fsub_rn((float)y, __fmul_rn(p.sign, fy[k]));
        t[k] = make_taps(round_trip(cx, wm1), round_trip(cy, hm1), p.H, p.W);
    }

    float gix[VEC], giy[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) gix[k] = giy[k] = 0.0f;

    const float *src = p.src ? p.src + (int64_t)n * p.src_bs : nullptr;
    float *gsrc = p.gsrc ? p.gsrc + (int64_t)n * p.gsrc_bs : nullptr;
    const float *go = p.gout + (int64_t)n * p.C * HW + pix;
#pragma unroll 1
    for (int c = 0; c < p.C; ++c) {
        float g[VEC];
        load_vec<VEC>(go + (int64_t)c * HW, g);
        if (p.gflow) {
            const float *s = src + (int64_t)c * HW;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float a = ldg_or_zero(s, t[k].o_nw), b = ldg_or_zero(s, t[k].o_ne);
                const float cc = ldg_or_zero(s, t[k].o_sw), d = ldg_or_zero(s, t[k].o_se);
                // d out / d ix and d out / d iy  
                gix[k] += g[k] * ((b - a) * t[k].sy + (d - cc) * t[k].ny);
                giy[k] += g[k] * ((cc - a) * t[k].ex + (d - b) * t[k].wx);
                if (p.add_flow != 0.0f && c < 2) {
                    if (c == 0) gix[k] -= p.sign * p.add_flow * g[k];
                    else giy[k] -= p.sign * p.add_flow * g[k];
                }
            }
        }
        if (gsrc) {
            float *gs = gsrc + (int64_t)c * HW;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (t[k].o_nw >= 0) atomicAdd(gs + t[k].o_nw, t[k].nw * g[k]);
                if (t[k].o_ne >= 0) atomicAdd(gs + t[k].o_ne, t[k].ne * g[k]);
                if (t[k].o_sw >= 0) atomicAdd(gs + t[k].o_sw, t[k].sw * g[k]);
                if (t[k].o_se >= 0) atomicAdd(gs + t[k].o_se, t[k].se * g[k]);
            }
        }
    }
    if (p.gflow) {
        float rx[VEC], ry[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            rx[k] = -p.sign * gix[k];   // d(ix)/d(flow_x) = -sign (the (W-1)/2 factors cancel)
            ry[k] = -p.sign * giy[k];
        }
        float *gf = p.gflow + (int64_t)n * 2 * HW + pix;
        store_vec<VEC>(gf, rx);
        store_vec<VEC>(gf + HW, ry);
    }
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_warp_t_fwd(const float *flow, int64_t flow_bs, float flow_sign,
                              const float *src, int64_t src_bs,
                              const uint8_t *src_mask, int64_t src_mask_bs,
                              const uint8_t *flow_mask, int64_t flow_mask_bs,
                              float add_flow, float valid_thresh,
                              float *out, uint8_t *valid,
                              int N, int C, int H, int W, void *stream)
{
    if (!flow || N <= 0 || C < 0 || (C > 0 && (!src || !out)) || (C == 0 && !valid)) return OFB_EINVAL;
    if (add_flow != 0.0f && C != 2) return OFB_EINVAL;
    if (H < 2 || W < 2) return OFB_ESHAPE;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    WarpFwdParams p{flow, flow_bs, flow_sign, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs,
                    add_flow, valid_thresh, out, valid, N, C, H, W};
    const bool vec4 = (W % 4 == 0) && aligned16(flow) && aligned16(out) && (flow_bs % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(valid) % 4 == 0);
    const dim3 block(32, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec4) {
        const dim3 grid((W / 4 + 31) / 32, (H + 7) / 8, N);
        warp_t_fwd_kernel<4><<<grid, block, 0, st>>>(p);
    } else {
        const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
        warp_t_fwd_kernel<1><<<grid, block, 0, st>>>(p);
    }
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_warp_t_bwd(const float *flow, int64_t flow_bs, float flow_sign,
                              const float *src, int64_t src_bs,
                              const float *gout, float add_flow,
                              float *gflow, float *gsrc, int64_t gsrc_bs,
                              int N, int C, int H, int W, void *stream)
{
    if (!flow || !gout || N <= 0 || C <= 0 || (gflow && !src)) return OFB_EINVAL;
    if (add_flow != 0.0f && C != 2) return OFB_EINVAL;
    if (H < 2 || W < 2) return OFB_ESHAPE;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!gflow && !gsrc) return 0;
    if (gsrc) {
        const int64_t items = (gsrc_bs == 0) ? 1 : N;
        cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(float) * (size_t)items * C * H * W, st);
        if (e != cudaSuccess) return (int)e;
    }
    WarpBwdParams p{flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W};
    const bool vec4 = (W % 4 == 0) && aligned16(flow) && aligned16(gout) && (flow_bs % 4 == 0) &&
                      (!gflow || aligned16(gflow));
    const dim3 block(32, 8);
    if (vec4) {
        const dim3 grid((W / 4 + 31) / 32, (H + 7) / 8, N);
        warp_t_bwd_kernel<4><<<grid, block, 0, st>>>(p);
    } else {
        const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
        warp_t_bwd_kernel<1><<<grid, block, 0, st>>>(p);
    }
    OFB_CHECK_LAUNCH();
    return 0;
}
