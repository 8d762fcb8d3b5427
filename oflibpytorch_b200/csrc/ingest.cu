// The callers and data formats either side of the warp/splat path:
//
//   flow extent   the masked min/max of the flow end points, i.e. everything Flow.get_padding needs
//                 (flow_class.py:1174-1224; the reference loops over the batch with four host syncs per item)
//   .flo ingest   Sintel .flo payload (H, W, 2) interleaved -> (2, H, W) planes   (utils.py:844-854 + to_tensor)
//   KITTI ingest  uint16 BGR png payload -> flow planes + valid mask              (utils.py:821-829)
//   resize        bilinear resize (align_corners = False, scale factor given) of a flow with its vectors
//                 rescaled, optionally with its mask (utils.py:878-916, flow_class.py:694-712), and the adjoint
//
// All of it is streaming byte/float work bound by HBM; nothing here is shaped like a GEMM.
#include <cmath>

#include "common.cuh"

namespace ofb {

// ---------------------------------------------------------------------------------------------------------------
// flow extent
// ---------------------------------------------------------------------------------------------------------------
// order-preserving float <-> int map so that integer atomicMin/atomicMax order floats (no NaNs here: flows are
// validated finite before any of this runs)
__device__ __forceinline__ int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void extent_init_kernel(int *ext, int n4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) ext[i] = (i & 1) ? f2ord(-INFINITY) : f2ord(INFINITY);      // [min_y, max_y, min_x, max_x] per item
}

// end point of every masked flow vector: pos = -((s * thresholded(v)) - grid), s = +1 for 't', -1 for 's'
// (flow_class.py:1195-1207, evaluated in fp32 like the reference's in-place tensor ops)
__global__ void __launch_bounds__(256) flow_extent_kernel(const float *__restrict__ flow,
                                                          const uint8_t *__restrict__ mask, const float s,
                                                          int *__restrict__ ext, const int H, const int W)
{
    const int n = blockIdx.y;
    const int HW = H * W;
    const float *f = flow + (int64_t)n * 2 * HW;
    const uint8_t *m = mask ? mask + (int64_t)n * HW : nullptr;
    float mny = INFINITY, mxy = -INFINITY, mnx = INFINITY, mxx = -INFINITY;
    for (int y = blockIdx.x; y < H; y += gridDim.x) {
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            const int i = y * W + x;
            if (m && __ldg(m + i) == 0) continue;
            float fx = __ldg(f + i), fy = __ldg(f + HW + i);
            if (fx < OFB_THRESHOLD && fx > -OFB_THRESHOLD) fx = 0.0f;      // utils.py:642
            if (fy < OFB_THRESHOLD && fy > -OFB_THRESHOLD) fy = 0.0f;
            const float px = -__fsub_rn(__fmul_rn(s, fx), (float)x);
            const float py = -__fsub_rn(__fmul_rn(s, fy), (float)y);
            mnx = fminf(mnx, px); mxx = fmaxf(mxx, px);
            mny = fminf(mny, py); mxy = fmaxf(mxy, py);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        int *e = ext + 4 * n;
        if (mny != INFINITY) atomicMin(e + 0, f2ord(mny));
        if (mxy != -INFINITY) atomicMax(e + 1, f2ord(mxy));
        if (mnx != INFINITY) atomicMin(e + 2, f2ord(mnx));
        if (mxx != -INFINITY) atomicMax(e + 3, f2ord(mxx));
    }
}

__global__ void extent_decode_kernel(int *ext, int n4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) reinterpret_cast<float *>(ext)[i] = ord2f(ext[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// ingest
// ---------------------------------------------------------------------------------------------------------------
// (N, HW, 2) interleaved -> (N, 2, HW) planes; VEC = 4 pixels per thread with 16-byte accesses when HW % 4 == 0
template <int VEC>
__global__ void __launch_bounds__(256) flo_to_planar_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                            const int HW)
{
    const int n = blockIdx.y;
    const float *src = in + (int64_t)n * 2 * HW;
    float *dst = out + (int64_t)n * 2 * HW;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= HW) return;
    if constexpr (VEC == 4) {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(src + 2 * i));
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(src + 2 * i + 4));
        *reinterpret_cast<float4 *>(dst + i) = make_float4(a.x, a.z, b.x, b.z);
        *reinterpret_cast<float4 *>(dst + HW + i) = make_float4(a.y, a.w, b.y, b.w);
    } else {
        dst[i] = __ldcs(src + 2 * i);
        dst[HW + i] = __ldcs(src + 2 * i + 1);
    }
}

// (N, HW, 3) uint16 in png/cv2 channel order [valid, v, u] -> flow planes ((u16 - 2^15) / 64, exact in fp32) and a
// 0/1 valid mask (utils.py:826-828)
__global__ void __launch_bounds__(256) kitti_decode_kernel(const uint16_t *__restrict__ in, float *__restrict__ flow,
                                                           uint8_t *__restrict__ valid, const int HW)
{
    const int n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const uint16_t *px = in + ((int64_t)n * HW + i) * 3;
    const int b = __ldcs(px), g = __ldcs(px + 1), r = __ldcs(px + 2);
    float *f = flow + (int64_t)n * 2 * HW;
    f[i] = (float)(r - 32768) * 0.015625f;
    f[HW + i] = (float)(g - 32768) * 0.015625f;
    if (valid) valid[(int64_t)n * HW + i] = b > 0 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// resize
// ---------------------------------------------------------------------------------------------------------------
struct ResizeGeom {
    int i0, i1;
    float l0, l1;
};

// ATen's bilinear source index with align_corners = False and a given scale factor: src = rs * (dst + 0.5) - 0.5
// clamped at 0, rs = float(1 / scale) (UpSample.h: area_pixel_compute_source_index, guard_index_and_lambda).  The
// CPU build of ATen contracts the multiply-subtract into one FMA and the two-term blends into  fma(w0, a, w1 * b);
// both were found by bit comparison (oracle/oracle.py: resize_bilinear) and are reproduced here so that the rounded
// mask of Flow.resize is bit-identical to the reference's.
__device__ __forceinline__ ResizeGeom resize_geom(int dst, float rs, int in_size)
{
    float src = __fmaf_rn(rs, __fadd_rn((float)dst, 0.5f), -0.5f);
    if (src < 0.0f) src = 0.0f;
    ResizeGeom g;
    g.i0 = min((int)src, in_size - 1);
    g.i1 = g.i0 + ((g.i0 < in_size - 1) ? 1 : 0);
    g.l1 = fminf(fmaxf(__fsub_rn(src, (float)g.i0), 0.0f), 1.0f);
    g.l0 = __fsub_rn(1.0f, g.l1);
    return g;
}

__device__ __forceinline__ float lerp2(float a, float b, float c, float d, const ResizeGeom &gy, const ResizeGeom &gx)
{
    // along x first, then along y
    const float top = __fmaf_rn(gx.l0, a, __fmul_rn(gx.l1, b));
    const float bot = __fmaf_rn(gx.l0, c, __fmul_rn(gx.l1, d));
    return __fmaf_rn(gy.l0, top, __fmul_rn(gy.l1, bot));
}

struct ResizeParams {
    const float *flow; const uint8_t *mask;
    float *out; uint8_t *out_mask;
    const float *gout; float *gin;
    int N, H, W, Ho, Wo;
    float rsy, rsx, sy, sx;      // reciprocal scales for the source index; vector scales (utils.py:908-909)
};

__global__ void __launch_bounds__(256) resize_fwd_kernel(const ResizeParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    if (x >= p.Wo) return;
    const ResizeGeom gy = resize_geom(y, p.rsy, p.H), gx = resize_geom(x, p.rsx, p.W);
    const int HW = p.H * p.W, HWo = p.Ho * p.Wo;
    const int o00 = gy.i0 * p.W + gx.i0, o01 = gy.i0 * p.W + gx.i1, o10 = gy.i1 * p.W + gx.i0, o11 = gy.i1 * p.W + gx.i1;
    if (p.flow) {
        const float *f = p.flow + (int64_t)n * 2 * HW;
        float *o = p.out + (int64_t)n * 2 * HWo + y * p.Wo + x;
        const float u = lerp2(__ldg(f + o00), __ldg(f + o01), __ldg(f + o10), __ldg(f + o11), gy, gx);
        const float v = lerp2(__ldg(f + HW + o00), __ldg(f + HW + o01), __ldg(f + HW + o10), __ldg(f + HW + o11), gy, gx);
        o[0] = __fmul_rn(u, p.sx);
        o[HWo] = __fmul_rn(v, p.sy);
    }
    if (p.mask) {
        // flow_class.py:707-712: bilinear resize of the 0/1 mask, then torch.round (half to even)
        const uint8_t *m = p.mask + (int64_t)n * HW;
        const float r = lerp2(m[o00] ? 1.0f : 0.0f, m[o01] ? 1.0f : 0.0f, m[o10] ? 1.0f : 0.0f, m[o11] ? 1.0f : 0.0f,
                              gy, gx);
        p.out_mask[(int64_t)n * HWo + y * p.Wo + x] = rintf(r) != 0.0f ? 1 : 0;
    }
}

// adjoint of the flow resize: every output pixel hands its gradient back to its four sources
__global__ void __launch_bounds__(256) resize_bwd_kernel(const ResizeParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    if (x >= p.Wo) return;
    const ResizeGeom gy = resize_geom(y, p.rsy, p.H), gx = resize_geom(x, p.rsx, p.W);
    const int HW = p.H * p.W, HWo = p.Ho * p.Wo;
    const float *g = p.gout + (int64_t)n * 2 * HWo + y * p.Wo + x;
    const float gu = __fmul_rn(__ldg(g), p.sx), gv = __fmul_rn(__ldg(g + HWo), p.sy);
    float *gi = p.gin + (int64_t)n * 2 * HW;
    const float w00 = gy.l0 * gx.l0, w01 = gy.l0 * gx.l1, w10 = gy.l1 * gx.l0, w11 = gy.l1 * gx.l1;
    const int o00 = gy.i0 * p.W + gx.i0, o01 = gy.i0 * p.W + gx.i1, o10 = gy.i1 * p.W + gx.i0, o11 = gy.i1 * p.W + gx.i1;
    atomicAdd(gi + o00, w00 * gu); atomicAdd(gi + o01, w01 * gu);
    atomicAdd(gi + o10, w10 * gu); atomicAdd(gi + o11, w11 * gu);
    atomicAdd(gi + HW + o00, w00 * gv); atomicAdd(gi + HW + o01, w01 * gv);
    atomicAdd(gi + HW + o10, w10 * gv); atomicAdd(gi + HW + o11, w11 * gv);
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_flow_extent(const float *flow, const uint8_t *mask, int is_source_ref, float *extent,
                               int N, int H, int W, void *stream)
{
    if (!flow || !extent || N <= 0 || H <= 0 || W <= 0) return OFB_EINVAL;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int *ext = reinterpret_cast<int *>(extent);
    const int n4 = 4 * N;
    extent_init_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(ext, n4);
    OFB_CHECK_LAUNCH();
    int rows_grid = (H < 148 * 8) ? H : 148 * 8;
    flow_extent_kernel<<<dim3(rows_grid, N), 256, 0, st>>>(flow, mask, is_source_ref ? -1.0f : 1.0f, ext, H, W);
    OFB_CHECK_LAUNCH();
    extent_decode_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(ext, n4);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_flo_to_planar(const float *interleaved, float *planar, int N, int H, int W, void *stream)
{
    if (!interleaved || !planar || N <= 0 || H <= 0 || W <= 0) return OFB_EINVAL;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    const int HW = H * W;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (HW % 4 == 0 && aligned16(interleaved) && aligned16(planar))
        flo_to_planar_kernel<4><<<dim3((HW / 4 + 255) / 256, N), 256, 0, st>>>(interleaved, planar, HW);
    else
        flo_to_planar_kernel<1><<<dim3((HW + 255) / 256, N), 256, 0, st>>>(interleaved, planar, HW);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_kitti_decode(const uint16_t *png_bgr, float *flow, uint8_t *valid, int N, int H, int W, void *stream)
{
    if (!png_bgr || !flow || N <= 0 || H <= 0 || W <= 0) return OFB_EINVAL;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    const int HW = H * W;
    kitti_decode_kernel<<<dim3((HW + 255) / 256, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(png_bgr, flow, valid, HW);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_resize_output_size(int in_size, double scale)
{
    return (int)std::floor((double)in_size * scale);        // ATen upsample: floor(input_size * scale_factor)
}

static bool resize_params(ResizeParams *p, int N, int H, int W, double scale_y, double scale_x)
{
    if (N <= 0 || H <= 0 || W <= 0 || !(scale_y > 0.0) || !(scale_x > 0.0) || N > 65535) return false;
    p->N = N; p->H = H; p->W = W;
    p->Ho = ofb_resize_output_size(H, scale_y);
    p->Wo = ofb_resize_output_size(W, scale_x);
    if (p->Ho < 1 || p->Wo < 1 || p->Ho > 65535 || (int64_t)p->Ho * p->Wo > 0x3fffffff ||
        (int64_t)H * W > 0x3fffffff)
        return false;
    p->rsy = (float)(1.0 / scale_y); p->rsx = (float)(1.0 / scale_x);
    p->sy = (float)scale_y; p->sx = (float)scale_x;
    return true;
}

extern "C" int ofb_resize_fwd(const float *flow, const uint8_t *mask, double scale_y, double scale_x,
                              float *out, uint8_t *out_mask, int N, int H, int W, void *stream)
{
    ResizeParams p{flow, mask, out, out_mask, nullptr, nullptr};
    if ((!flow && !mask) || (flow && !out) || (mask && !out_mask)) return OFB_EINVAL;
    if (!resize_params(&p, N, H, W, scale_y, scale_x)) return OFB_EINVAL;
    resize_fwd_kernel<<<dim3((p.Wo + 255) / 256, p.Ho, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_resize_bwd(const float *gout, double scale_y, double scale_x, float *gin,
                              int N, int H, int W, void *stream)
{
    ResizeParams p{nullptr, nullptr, nullptr, nullptr, gout, gin};
    if (!gout || !gin) return OFB_EINVAL;
    if (!resize_params(&p, N, H, W, scale_y, scale_x)) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)N * 2 * H * W, st);
    if (e != cudaSuccess) return (int)e;
    resize_bwd_kernel<<<dim3((p.Wo + 255) / 256, p.Ho, N), 256, 0, st>>>(p);
    OFB_CHECK_LAUNCH();
    return 0;
}
