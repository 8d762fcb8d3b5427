// Shared device helpers for the sm_100a flow-field kernels.
//
// Arithmetic contract: fp32, round-to-nearest, IEEE division, no implicit FMA.  The file is compiled
// with -fmad=false and every fused multiply-add is written as __fmaf_rn, because the boolean validity
// masks are only bit-exact if the coordinate round trip and the bilinear accumulation are evaluated
// operation by operation like the reference's CPU path (oracle/oracle.c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/oflib_b200.h"

#define OFB_THRESHOLD 1e-3f     // utils.py:23
#define OFB_MIN_DENSITY 1e-3f   // utils.py:1144

namespace ofb {

extern std::atomic<int64_t> g_launches;   // host-side launch counter (api.cu)

// pixel -> normalised [-1,1] (utils.py:462-465) -> pixel (ATen GridSampler.h:31), in fp32
__device__ __forceinline__ float round_trip(float c, float size_m1)
{
    float g = __fmul_rn(c, 2.0f);
    g = __fdiv_rn(g, size_m1);
    g = __fsub_rn(g, 1.0f);
    g = __fadd_rn(g, 1.0f);
    g = __fmul_rn(g, 0.5f);          // == g / 2 exactly
    return __fmul_rn(g, size_m1);
}

// x / d for a loop-invariant divisor d with rcp = RN(1/d) computed once (host, IEEE division).
// q0 = RN(x*rcp) is within 2 ulp of x/d; each step  q += RN(x - d*q) * rcp  (the residual is exact with an FMA)
// brings a value within 1 ulp to the correctly rounded quotient (Markstein's theorem), so two steps from q0 give
// RN(x/d) for |x| < 2^24 as long as the quotient is a normal number.  Denormal quotients can differ from IEEE in
// exact-halfway cases, which is immaterial here: the round trip subtracts 1 next, absorbing anything below 2^-25.
// |x| >= 2^24 takes the IEEE division.  tests/test_gpu_parity.py::test_round_trip_exhaustive checks
// round_trip_fast == round_trip over EVERY float bit pattern for the image sizes in use.
struct ExactDiv {
    float d, rcp;
};

__device__ __forceinline__ float div_exact(float x, const ExactDiv &e)
{
    if (!(fabsf(x) < 16777216.0f)) return __fdiv_rn(x, e.d);
    float q = __fmul_rn(x, e.rcp);
    q = __fmaf_rn(__fmaf_rn(-e.d, q, x), e.rcp, q);
    q = __fmaf_rn(__fmaf_rn(-e.d, q, x), e.rcp, q);
    return q;
}

// pixel -> normalised -> pixel with the division done by div_exact (bit-identical to round_trip)
__device__ __forceinline__ float round_trip_fast(float c, const ExactDiv &e)
{
    float g = __fmul_rn(c, 2.0f);
    g = div_exact(g, e);
    g = __fsub_rn(g, 1.0f);
    g = __fadd_rn(g, 1.0f);
    g = __fmul_rn(g, 0.5f);
    return __fmul_rn(g, e.d);
}

// round_trip_fast without the |x| < 2^24 test of div_exact.  For d <= 65535 (the launcher checks) a coordinate with
// |2c| >= 2^24 lies at least 127 field widths outside the field whatever the last bits of the quotient are: every tap
// is out of bounds in both evaluations, the weights are finite, the output is 0 and the validity bit false, exactly
// as with the IEEE division (an infinite coordinate gives NaN weights and a NaN output in both).  Inside |2c| < 2^24
// this IS round_trip_fast.
__device__ __forceinline__ float round_trip_unchecked(float c, const ExactDiv &e)
{
    const float x = __fmul_rn(c, 2.0f);
    float q = __fmul_rn(x, e.rcp);
    q = __fmaf_rn(__fmaf_rn(-e.d, q, x), e.rcp, q);
    q = __fmaf_rn(__fmaf_rn(-e.d, q, x), e.rcp, q);
    float g = __fsub_rn(q, 1.0f);
    g = __fadd_rn(g, 1.0f);
    g = __fmul_rn(g, 0.5f);
    return __fmul_rn(g, e.d);
}

// The four bilinear taps of one sample position (zero padding: out-of-bounds taps read as 0).
struct Taps {
    float nw, ne, sw, se;       // weights
    float ex, wx, sy, ny;       // x1-ix, ix-x0, y1-iy, iy-y0
    int o_nw, o_ne, o_sw, o_se; // offsets inside one H*W plane, -1 when out of bounds
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W)
{
    Taps t;
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);
    t.ex = __fsub_rn(x1, ix);
    t.wx = __fsub_rn(ix, x0);
    t.sy = __fsub_rn(y1, iy);
    t.ny = __fsub_rn(iy, y0);
    t.nw = __fmul_rn(t.ex, t.sy);
    t.ne = __fmul_rn(t.wx, t.sy);
    t.sw = __fmul_rn(t.ex, t.ny);
    t.se = __fmul_rn(t.wx, t.ny);
    const float wm = (float)(W - 1), hm = (float)(H - 1);
    const bool bx0 = (x0 >= 0.0f) && (x0 <= wm), bx1 = (x1 >= 0.0f) && (x1 <= wm);
    const bool by0 = (y0 >= 0.0f) && (y0 <= hm), by1 = (y1 >= 0.0f) && (y1 <= hm);
    const int xi0 = bx0 ? (int)x0 : 0, xi1 = bx1 ? (int)x1 : 0;
    const int yi0 = by0 ? (int)y0 : 0, yi1 = by1 ? (int)y1 : 0;
    t.o_nw = (bx0 && by0) ? yi0 * W + xi0 : -1;
    t.o_ne = (bx1 && by0) ? yi0 * W + xi1 : -1;
    t.o_sw = (bx0 && by1) ? yi1 * W + xi0 : -1;
    t.o_se = (bx1 && by1) ? yi1 * W + xi1 : -1;
    return t;
}

// ATen's CPU accumulation order, bit for bit: fma(se,d, fma(sw,c, fma(ne,b, nw*a)))
__device__ __forceinline__ float blend(const Taps &t, float a, float b, float c, float d)
{
    float r = __fmul_rn(a, t.nw);
    r = __fmaf_rn(b, t.ne, r);
    r = __fmaf_rn(c, t.sw, r);
    r = __fmaf_rn(d, t.se, r);
    return r;
}

__device__ __forceinline__ float ldg_or_zero(const float *__restrict__ p, int off)
{
    return off >= 0 ? __ldg(p + off) : 0.0f;
}

__device__ __forceinline__ bool is_zero_vec(float a, float b)
{
    return (a < OFB_THRESHOLD) && (a > -OFB_THRESHOLD) && (b < OFB_THRESHOLD) && (b > -OFB_THRESHOLD);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace ofb

#define OFB_CHECK_LAUNCH()                      \
    do {                                        \
        ++ofb::g_launches;                      \
        cudaError_t e__ = cudaGetLastError();   \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)
