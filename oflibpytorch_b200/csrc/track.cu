// Point sampling of a flow field (the sparse half of track_pts, reference utils.py:1003-1019):
//   flow_vecs = grid_sample(flow, normalise_coords(pts[..., (x, y)]), align_corners=True), channels flipped to (y, x).
// M points per batch item, not a dense pass: one thread per point, same coordinate round trip, tap geometry and
// accumulation order as the dense warp (common.cuh), so the result is bit-identical to the reference's CPU output.
#include "common.cuh"

namespace ofb {

struct PtGeom {
    float wnw, wne, wsw, wse, ex, wx, sy, ny;
    int o;
    bool pnw, pne, psw, pse;
};

__device__ __forceinline__ PtGeom pt_geom(float px, float py, int H, int W)
{
    PtGeom g;
    const float ix = round_trip(px, (float)(W - 1)), iy = round_trip(py, (float)(H - 1));
    const float x0 = floorf(ix), y0 = floorf(iy);
    g.ex = __fsub_rn(__fadd_rn(x0, 1.0f), ix); g.wx = __fsub_rn(ix, x0);
    g.sy = __fsub_rn(__fadd_rn(y0, 1.0f), iy); g.ny = __fsub_rn(iy, y0);
    g.wnw = __fmul_rn(g.ex, g.sy); g.wne = __fmul_rn(g.wx, g.sy);
    g.wsw = __fmul_rn(g.ex, g.ny); g.wse = __fmul_rn(g.wx, g.ny);
    const int xi = __float2int_rd(ix), yi = __float2int_rd(iy);
    const bool bx0 = (unsigned)xi < (unsigned)W, bx1 = (unsigned)(xi + 1) < (unsigned)W;
    const bool by0 = (unsigned)yi < (unsigned)H, by1 = (unsigned)(yi + 1) < (unsigned)H;
    g.pnw = bx0 && by0; g.pne = bx1 && by0; g.psw = bx0 && by1; g.pse = bx1 && by1;
    g.o = (int)((unsigned)yi * (unsigned)W + (unsigned)xi);
    return g;
}

__global__ void sample_pts_fwd_kernel(const float *__restrict__ flow, int64_t flow_bs, const float *__restrict__ pts,
                                      float *__restrict__ out, int M, int H, int W)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (m >= M) return;
    const float *pp = pts + ((int64_t)n * M + m) * 2;
    const PtGeom g = pt_geom(__ldg(pp + 1), __ldg(pp), H, W);          // pts are (y, x)
    const int HW = H * W;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float *s = flow + (int64_t)n * flow_bs + (int64_t)c * HW + g.o;
        const float a = g.pnw ? __ldg(s) : 0.0f, b = g.pne ? __ldg(s + 1) : 0.0f;
        const float cq = g.psw ? __ldg(s + W) : 0.0f, d = g.pse ? __ldg(s + W + 1) : 0.0f;
        float r = __fmul_rn(a, g.wnw);
        r = __fmaf_rn(b, g.wne, r);
        r = __fmaf_rn(cq, g.wsw, r);
        r = __fmaf_rn(d, g.wse, r);
        out[((int64_t)n * M + m) * 2 + (1 - c)] = r;                   // .flip(1): (x, y) channels -> (y, x)
    }
}

// adjoint: gflow (N,2,H,W, zero-filled by the launcher) via atomics, gpts (N,M,2) in (y, x) order
__global__ void sample_pts_bwd_kernel(const float *__restrict__ flow, int64_t flow_bs, const float *__restrict__ pts,
                                      const float *__restrict__ gout, float *gflow, int64_t gflow_bs,
                                      float *__restrict__ gpts, int M, int H, int W)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (m >= M) return;
    const float *pp = pts + ((int64_t)n * M + m) * 2;
    const PtGeom g = pt_geom(__ldg(pp + 1), __ldg(pp), H, W);
    const int HW = H * W;
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float go = __ldg(gout + ((int64_t)n * M + m) * 2 + (1 - c));
        const float *s = flow + (int64_t)n * flow_bs + (int64_t)c * HW + g.o;
        const float a = g.pnw ? __ldg(s) : 0.0f, b = g.pne ? __ldg(s + 1) : 0.0f;
        const float cq = g.psw ? __ldg(s + W) : 0.0f, d = g.pse ? __ldg(s + W + 1) : 0.0f;
        gx += go * ((b - a) * g.sy + (d - cq) * g.ny);
        gy += go * ((cq - a) * g.ex + (d - b) * g.wx);
        if (gflow) {
            float *t = gflow + (int64_t)n * gflow_bs + (int64_t)c * HW + g.o;
            if (g.pnw) atomicAdd(t, g.wnw * go);
            if (g.pne) atomicAdd(t + 1, g.wne * go);
            if (g.psw) atomicAdd(t + W, g.wsw * go);
            if (g.pse) atomicAdd(t + W + 1, g.wse * go);
        }
    }
    if (gpts) {
        gpts[((int64_t)n * M + m) * 2] = gy;
        gpts[((int64_t)n * M + m) * 2 + 1] = gx;
    }
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_sample_pts_fwd(const float *flow, int64_t flow_bs, const float *pts, float *out,
                                  int N, int M, int H, int W, void *stream)
{
    if (!flow || !pts || !out || N <= 0 || M < 0) return OFB_EINVAL;
    if (H < 2 || W < 2) return OFB_ESHAPE;
    if (M == 0) return 0;
    sample_pts_fwd_kernel<<<dim3((M + 127) / 128, N), 128, 0, static_cast<cudaStream_t>(stream)>>>(flow, flow_bs, pts, out,
                                                                                                 M, H, W);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_sample_pts_bwd(const float *flow, int64_t flow_bs, const float *pts, const float *gout,
                                  float *gflow, int64_t gflow_bs, float *gpts, int N, int M, int H, int W, void *stream)
{
    if (!flow || !pts || !gout || N <= 0 || M < 0) return OFB_EINVAL;
    if (H < 2 || W < 2) return OFB_ESHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (gflow) {
        cudaError_t e = cudaMemsetAsync(gflow, 0, sizeof(float) * 2 * (size_t)H * W * (gflow_bs == 0 ? 1 : N), st);
        if (e != cudaSuccess) return (int)e;
    }
    if (M == 0 || (!gflow && !gpts)) return 0;
    sample_pts_bwd_kernel<<<dim3((M + 127) / 128, N), 128, 0, st>>>(flow, flow_bs, pts, gout, gflow, gflow_bs, gpts, M, H, W);
    OFB_CHECK_LAUNCH();
    return 0;
}
