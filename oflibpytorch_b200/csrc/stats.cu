// One pass over a flow field that answers every question the reference asks of it before the hot path
// starts: finite? (utils.py:98)  all zero? (utils.py:919-938)  all zero where the mask is set?
// (flow_class.py:1226-1244).  The reference spends a clone, two compares, a masked fill, a reduction and
// N host syncs on this per call; here it is one 8 B/pixel read and one int per batch item.
#include "common.cuh"

namespace ofb {

// Every question reduces to the LARGEST |component| as an unsigned bit pattern (IEEE order = integer order for
// non-negative floats, NaN/Inf patterns sort above every finite number):
//   non-finite  <=>  max >= 0x7f800000        != 0  <=>  max != 0        |v| >= 1e-3 or NaN  <=>  max >= bits(1e-3f)
// so the per-element work is one AND and one unsigned max (plus a multiply by the 0/1 mask byte, which runs on the
// FMA pipe, and a second max for the masked variant).  The first version classified each element with five compares
// and was bound by the ALU pipe (ncu: 90 % ALU, half the read bandwidth inside a step).
struct AbsMax {
    unsigned all, masked;
    __device__ __forceinline__ void add(float v, unsigned m)
    {
        const unsigned u = __float_as_uint(v) & 0x7fffffffu;
        all = max(all, u);
        masked = max(masked, u * m);
    }
};

constexpr int kStatsIters = 8;

template <int VEC>
__global__ void __launch_bounds__(256) flow_stats_kernel(const float *__restrict__ flow,
                                                         const uint8_t *__restrict__ mask,
                                                         int32_t *__restrict__ flags, int HW, int summary)
{
    const int n = blockIdx.y;
    const float *f = flow + (int64_t)n * 2 * HW;
    const uint8_t *m = mask ? mask + (int64_t)n * HW : nullptr;
    AbsMax am{0u, 0u};
    // a CTA owns one contiguous chunk of kStatsIters * 256 * VEC pixels
    const int base = blockIdx.x * (kStatsIters * 256 * VEC) + threadIdx.x * VEC;
#pragma unroll 4
    for (int k = 0; k < kStatsIters; ++k) {
        const int i = base + k * (256 * VEC);
        if (i >= HW) break;
        if constexpr (VEC == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(f + i));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(f + HW + i));
            uchar4 mm = make_uchar4(1, 1, 1, 1);
            if (m) mm = __ldg(reinterpret_cast<const uchar4 *>(m + i));
            const unsigned m0 = mm.x != 0, m1 = mm.y != 0, m2 = mm.z != 0, m3 = mm.w != 0;
            am.add(a.x, m0); am.add(a.y, m1); am.add(a.z, m2); am.add(a.w, m3);
            am.add(b.x, m0); am.add(b.y, m1); am.add(b.z, m2); am.add(b.w, m3);
        } else {
            const unsigned mm = m ? (__ldg(m + i) != 0) : 1u;
            am.add(__ldg(f + i), mm);
            am.add(__ldg(f + HW + i), mm);
        }
    }
    // one flag update per CTA, and none once the item's flags already hold these bits: 2000 warps OR-ing into the
    // same word serialise in one L2 slice and doubled the kernel time (tools/readbench.cu vs the first version)
    __shared__ unsigned s_all[8], s_masked[8];
    unsigned all = __reduce_max_sync(0xffffffffu, am.all), masked = __reduce_max_sync(0xffffffffu, am.masked);
    if ((threadIdx.x & 31) == 0) { s_all[threadIdx.x >> 5] = all; s_masked[threadIdx.x >> 5] = masked; }
    __syncthreads();
    if (threadIdx.x < 32) {
        all = __reduce_max_sync(0xffffffffu, threadIdx.x < 8 ? s_all[threadIdx.x] : 0u);
        masked = __reduce_max_sync(0xffffffffu, threadIdx.x < 8 ? s_masked[threadIdx.x] : 0u);
        if (threadIdx.x == 0 && all != 0) {
            const unsigned thr = __float_as_uint(OFB_THRESHOLD);
            int fl = OFB_STAT_NONZERO;
            if (all >= 0x7f800000u) fl |= OFB_STAT_NONFINITE;
            if (all >= thr) fl |= OFB_STAT_NONZERO_T;
            if (masked != 0) fl |= OFB_STAT_NONZERO_M;
            if (masked >= thr) fl |= OFB_STAT_NONZERO_TM;
            if ((*reinterpret_cast<volatile int32_t *>(flags + n) & fl) != fl) atomicOr(flags + n, fl);
            // batch summary (OR over the items): what the reference's all(is_zero_flow(..)) looks at (utils.py:497)
            if (summary >= 0 && (*reinterpret_cast<volatile int32_t *>(flags + summary) & fl) != fl)
                atomicOr(flags + summary, fl);
        }
    }
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

static int flow_stats_impl(const float *flow, const uint8_t *mask, int32_t *flags, int N, int H, int W, int summary,
                           void *stream)
{
    if (!flow || !flags || N <= 0 || H <= 0 || W <= 0) return OFB_EINVAL;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    const int HW = H * W;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (HW % 4 == 0) && aligned16(flow) && (reinterpret_cast<uintptr_t>(mask) % 4 == 0);
    const int per_block = 256 * (vec4 ? 4 : 1) * kStatsIters;
    int gx = (HW + per_block - 1) / per_block;
    if (gx < 1) gx = 1;
    const dim3 grid(gx, N);
    if (vec4) flow_stats_kernel<4><<<grid, 256, 0, st>>>(flow, mask, flags, HW, summary);
    else flow_stats_kernel<1><<<grid, 256, 0, st>>>(flow, mask, flags, HW, summary);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_flow_stats(const float *flow, const uint8_t *mask, int32_t *flags,
                              int N, int H, int W, void *stream)
{
    return flow_stats_impl(flow, mask, flags, N, H, W, -1, stream);
}

extern "C" int ofb_flow_stats_batch(const float *flow, const uint8_t *mask, int32_t *flags,
                                    int N, int H, int W, void *stream)
{
    return flow_stats_impl(flow, mask, flags, N, H, W, N, stream);
}
