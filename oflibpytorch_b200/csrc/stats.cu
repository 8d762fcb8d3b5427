// One pass over a flow field that answers every question the reference asks of it before the hot path
// starts: finite? (utils.py:98)  all zero? (utils.py:919-938)  all zero where the mask is set?
// (flow_class.py:1226-1244).  The reference spends a clone, two compares, a masked fill, a reduction and
// N host syncs on this per call; here it is one 8 B/pixel read and one int per batch item.
#include "common.cuh"

namespace ofb {

__device__ __forceinline__ int classify(float v, bool m)
{
    int f = 0;
    if (!isfinite(v)) f |= OFB_STAT_NONFINITE;
    if (v != 0.0f) f |= OFB_STAT_NONZERO | (m ? OFB_STAT_NONZERO_M : 0);
    if (!((v < OFB_THRESHOLD) && (v > -OFB_THRESHOLD))) f |= OFB_STAT_NONZERO_T | (m ? OFB_STAT_NONZERO_TM : 0);
    return f;
}

template <int VEC>
__global__ void __launch_bounds__(256) flow_stats_kernel(const float *__restrict__ flow,
                                                         const uint8_t *__restrict__ mask,
                                                         int32_t *__restrict__ flags, int HW)
{
    const int n = blockIdx.y;
    const float *f = flow + (int64_t)n * 2 * HW;
    const uint8_t *m = mask ? mask + (int64_t)n * HW : nullptr;
    int acc = 0;
    const int stride = gridDim.x * blockDim.x * VEC;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < HW; i += stride) {
        if constexpr (VEC == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(f + i));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(f + HW + i));
            uchar4 mm = make_uchar4(1, 1, 1, 1);
            if (m) mm = __ldg(reinterpret_cast<const uchar4 *>(m + i));
            acc |= classify(a.x, mm.x) | classify(a.y, mm.y) | classify(a.z, mm.z) | classify(a.w, mm.w);
            acc |= classify(b.x, mm.x) | classify(b.y, mm.y) | classify(b.z, mm.z) | classify(b.w, mm.w);
        } else {
            const bool mm = m ? (__ldg(m + i) != 0) : true;
            acc |= classify(__ldg(f + i), mm) | classify(__ldg(f + HW + i), mm);
        }
    }
    acc = __reduce_or_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicOr(flags + n, acc);
}

// self-test: the coordinate round trip with the constant-reciprocal division must equal the one with the IEEE
// division for every float bit pattern (both signs)
__global__ void div_selftest_kernel(float d, float rcp, uint32_t lo, uint32_t hi, unsigned long long *bad)
{
    const ExactDiv e{d, rcp};
    unsigned long long local = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += stride) {
        for (int sgn = 0; sgn < 2; ++sgn) {
            const float x = __uint_as_float((uint32_t)b | (sgn ? 0x80000000u : 0u));
            const float q = round_trip_fast(x, e), r = round_trip(x, d);
            if (__float_as_uint(q) != __float_as_uint(r) && !(q != q && r != r)) ++local;
        }
    }
    if (local) atomicAdd(bad, local);
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_selftest_div(float d, uint32_t lo_bits, uint32_t hi_bits, unsigned long long *mismatches,
                                void *stream)
{
    if (!mismatches || !(d >= 1.0f)) return OFB_EINVAL;
    div_selftest_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(d, 1.0f / d, lo_bits, hi_bits,
                                                                                 mismatches);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_flow_stats(const float *flow, const uint8_t *mask, int32_t *flags,
                              int N, int H, int W, void *stream)
{
    if (!flow || !flags || N <= 0 || H <= 0 || W <= 0) return OFB_EINVAL;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    const int HW = H * W;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (HW % 4 == 0) && aligned16(flow) && (reinterpret_cast<uintptr_t>(mask) % 4 == 0);
    const int per_block = 256 * (vec4 ? 4 : 1) * 8;   // 8 iterations per thread
    int gx = (HW + per_block - 1) / per_block;
    if (gx < 1) gx = 1;
    const dim3 grid(gx, N);
    if (vec4) flow_stats_kernel<4><<<grid, 256, 0, st>>>(flow, mask, flags, HW);
    else flow_stats_kernel<1><<<grid, 256, 0, st>>>(flow, mask, flags, HW);
    OFB_CHECK_LAUNCH();
    return 0;
}
