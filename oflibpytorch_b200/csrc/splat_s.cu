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
#include <cstdlib>

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
    // Row-band frame (identity: src_y0 = 0, gh = H, acc_y0 = 0, acc_h = H).  The H x W tensors are rows
    // [src_y0, src_y0 + H) of a global field of gh rows; the accumulator stores global rows [acc_y0, acc_y0 + acc_h).
    // Used by the v1 kernels only (spatial sharding of one oversized field).
    int src_y0, gh, acc_y0, acc_h;
    // SPLIT layout only (C + 1 fills the vector, e.g. RGB + density): the bad-weight channel is a plane of its own
    // behind the vectors instead of a second float4 per pixel that would be four fifths padding
    float *acc_bad;
    // Column frame, same idea (identity: src_x0 = 0, gw = W): the tensors are columns [src_x0, src_x0 + W) of a
    // global field of gw columns, and the accumulator stores exactly those columns.  Used by the padded 's' apply:
    // end points are computed in the padded frame like the reference does, what lands outside the window is cut.
    int src_x0, gw;
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

// ------------------------------------------------------------------------------------------------
// deterministic forward (opt-in, ofb_set_option key 4): the merge is done in 64-bit FIXED POINT.  Integer addition is
// associative, so the accumulated sums -- and with them every float output -- are bit-identical from run to run
// whatever order the L2 applies the atomics in ("deterministic global merge", reference order utils.py:1132-1144).
//   numerators    round(w * d * 2^k),  2^k = 2^40 / 2^ceil(log2 max|d|)  (max|d| from one reduction pass; the product
//                 w * d is exact in double, the power of two too: the only rounding is the final conversion)
//   density, bad  max(1, round(w * 2^40)): a positive weight never quantises to zero, so `density > 0` (utils.py:1197)
//                 and `bad == 0` keep their exact meaning
// C + 2 planes of int64 per item, scalar 64-bit atomics: about 3x slower than the float path, which is why it is a
// mode and not the default.
// ------------------------------------------------------------------------------------------------
int g_splat_det = 0;

__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ v, int64_t count, unsigned *out)
{
    unsigned m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(__ldg(v + i)) & 0x7fffffffu);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m != 0 && *reinterpret_cast<volatile unsigned *>(out) < m) atomicMax(out, m);
}

// 2^k with 2^k * max|d| in [2^39, 2^40]; 1 when the data are all zero or not finite
__device__ __forceinline__ double det_scale(const unsigned *maxbits)
{
    const unsigned b = __ldg(maxbits);
    if (b == 0 || b >= 0x7f800000u) return 1.0;
    int e;
    frexpf(__uint_as_float(b), &e);                 // |d| = m * 2^e, m in [0.5, 1)
    return ldexp(1.0, 40 - e);
}

__device__ __forceinline__ void det_add(unsigned long long *a, long long v)
{
    atomicAdd(a, (unsigned long long)v);            // two's complement: negative contributions wrap correctly
}

__global__ void __launch_bounds__(256) splat_accumulate_det_kernel(const SplatParams p, unsigned long long *acc64,
                                                                   const unsigned *maxbits)
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
    unsigned long long *acc = acc64 + (int64_t)n * (p.C + 2) * HW;
    const float *data = p.data + (int64_t)n * p.data_bs + pix;
    const double scale = det_scale(maxbits);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const float w = __fmul_rn(k.wy[a], k.wx[b]);                     // utils.py:1114
            if (!(w > 0.0f)) continue;
            const int q = p.W * k.Y[a] + k.X[b];                             // utils.py:1118
            const long long wq = max(1ll, __double2ll_rn((double)w * 1099511627776.0));   // 2^40
            det_add(acc + (int64_t)p.C * HW + q, wq);                        // utils.py:1133
            if (is_bad) det_add(acc + (int64_t)(p.C + 1) * HW + q, wq);
            for (int c = 0; c < p.C; ++c)                                    // utils.py:1139-1140
                det_add(acc + (int64_t)c * HW + q, __double2ll_rn((double)w * (double)__ldg(data + (int64_t)c * HW) * scale));
        }
}

__global__ void __launch_bounds__(256) splat_normalise_det_kernel(const SplatParams p, const unsigned long long *acc64,
                                                                  const unsigned *maxbits)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int HW = p.H * p.W;
    const int pix = y * p.W + x;
    const unsigned long long *acc = acc64 + (int64_t)n * (p.C + 2) * HW + pix;
    const double inv40 = 1.0 / 1099511627776.0, inv_scale = 1.0 / det_scale(maxbits);
    const float den = (float)((double)(long long)acc[(int64_t)p.C * HW] * inv40);
    const float bad = p.data_mask ? (float)((double)(long long)acc[(int64_t)(p.C + 1) * HW] * inv40) : 0.0f;
    const float dc = fmaxf(den, OFB_MIN_DENSITY);
    const bool wm = acc[(int64_t)p.C * HW] != 0ull;
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
        const float num = (float)((double)(long long)acc[(int64_t)c * HW] * inv_scale);
        out[(int64_t)c * HW] = unocc ? __ldg(data + (int64_t)c * HW) : __fdiv_rn(num, dc);
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

// floats of workspace the deterministic path needs: C + 2 int64 planes per item and the max|data| word
static int64_t splat_det_floats(int N, int C, int H, int W) { return (int64_t)N * (C + 2) * H * W * 2 + 4; }

// ------------------------------------------------------------------------------------------------
// v1 forward: interleaved accumulators, vector REDs, register merging
// ------------------------------------------------------------------------------------------------
// The accumulator is this library's private workspace, so its layout is free: KP = 4 or 8 floats PER TARGET PIXEL
// ([num_0 .. num_{C-1}, den, bad, pad]) instead of C+2 planes.  Every tap site is then ONE
// red.global.add.v4.f32 (SASS REDG.E.ADD.F32x4) instead of C+2 scalar REDs, and the normalise pass reads one
// float4 per pixel.  Contributions are merged in registers before they reach the L2 exactly as in the W adjoint
// (csrc/warp_t.cu): lane l's east taps are lane l+1's west taps when the endpoints chain, and a thread carries its
// south taps into the next source row.  The whole batch goes through zero-fill, accumulate, normalise in three
// launches (OFB_SPLAT_CHUNK_MB splits it): per-item passes meant to keep the accumulator in the L2 were measured
// slower, with and without a persisting access-policy window, and so was a persistent kernel sweeping an L2-sized
// ring of accumulator rows (profiles/r01_ring_splat_experiment.patch).
#ifndef OFB_SPLAT_ROWS
#define OFB_SPLAT_ROWS 8
#endif
constexpr int kSplatRows = OFB_SPLAT_ROWS;   // source rows per thread; CTA tile = 64 x 4*rows (2 warps across x, 4 across y)

template <int KP> struct AccVec { float v[KP]; };

template <int KP>
__device__ __forceinline__ void red_vec(float *acc, int q, const AccVec<KP> &a, bool pred)
{
    float *t = acc + (int64_t)q * KP;
    if constexpr (KP == 1) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q red.global.add.f32 [%0], %1;\n\t}"
                     ::"l"(t), "f"(a.v[0]), "r"((int)pred));
    } else if constexpr (KP == 2) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q red.global.add.v2.f32 [%0], {%1, %2};\n\t}"
                     ::"l"(t), "f"(a.v[0]), "f"(a.v[1]), "r"((int)pred));
    } else {
#pragma unroll
        for (int h = 0; h < KP / 4; ++h)
            asm volatile(
                "{\n\t"
                ".reg .pred q;\n\t"
                "setp.ne.s32 q, %5, 0;\n\t"
                "@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t"
                "}" ::"l"(t + 4 * h), "f"(a.v[4 * h]), "f"(a.v[4 * h + 1]), "f"(a.v[4 * h + 2]), "f"(a.v[4 * h + 3]),
                "r"((int)pred));
    }
}

#ifndef OFB_SPLAT_PF
#define OFB_SPLAT_PF 128
#endif
#ifndef OFB_SPLAT_MINB
#define OFB_SPLAT_MINB 3
#endif
template <int C, bool HAS_BAD, int KP, bool SPLIT = false>
__global__ void __launch_bounds__(256, OFB_SPLAT_MINB) splat_accumulate_v1_kernel(const SplatParams p, const int n0)
{
    constexpr bool BAD_SLOT = HAS_BAD && !SPLIT;       // bad weight rides in the vector
    static_assert(C + 1 + (BAD_SLOT ? 1 : 0) <= KP && (!SPLIT || HAS_BAD), "accumulator slots");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kSplatRows) + (warp >> 1) * kSplatRows;
    const int n = n0 + blockIdx.z;
    const int HW = p.H * p.W;
    const bool active = x < p.W;                         // inactive lanes still take part in the shuffles
    const int xc = min(x, p.W - 1);
    const int rows = min(kSplatRows, p.H - y0);          // warp-uniform
    if (rows <= 0) return;                               // whole warp below the field (the row prefetch must not run)

    const float *cx = p.cx + (int64_t)n * p.coord_bs, *cy = p.cy + (int64_t)n * p.coord_bs;
    const float *data = (C > 0) ? p.data + (int64_t)n * p.data_bs : nullptr;
    const uint8_t *dmask = HAS_BAD ? p.data_mask + (int64_t)n * p.data_mask_bs : nullptr;
    const uint8_t *fmask = p.flow_mask ? p.flow_mask + (int64_t)n * p.flow_mask_bs : nullptr;
    float *acc = p.acc + (int64_t)blockIdx.z * p.acc_h * p.W * KP;  // workspace item = position inside the chunk
    float *accb = SPLIT ? p.acc_bad + (int64_t)blockIdx.z * p.acc_h * p.W : nullptr;
    asm volatile("" : "+l"(cx), "+l"(cy), "+l"(data), "+l"(dmask), "+l"(fmask), "+l"(acc), "+l"(accb));

    // L2 prefetch of the source rows the CTA launched OFB_SPLAT_PF slots later will stream (coordinates, payload planes,
    // masks), fire-and-forget -- the kernel is bound by memory latency at 3 CTAs per SM (see warp_t_fwd_lean_kernel)
    if (OFB_SPLAT_PF > 0) {
        const long long lin = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + OFB_SPLAT_PF;
        const long long total = (long long)gridDim.x * gridDim.y * gridDim.z;
        if (lin < total) {
            const int pz = (int)(lin / ((long long)gridDim.x * gridDim.y));
            const int rem = (int)(lin - (long long)pz * gridDim.x * gridDim.y);
            const int py = rem / (int)gridDim.x, px = rem - py * (int)gridDim.x;
            const int nn = n0 + pz;
            constexpr int kRows = 4 * kSplatRows, kPlane = 2 * kRows;        // two 128-byte lines per row and plane
            constexpr int kLines = (2 + C) * kPlane;
            for (int idx = threadIdx.x; idx < kLines + 2 * kRows; idx += 256) {
                const void *a = nullptr;
                if (idx < kLines) {
                    const int pl = idx / kPlane, j = idx - pl * kPlane;
                    const int o = min(py * kRows + (j >> 1), p.H - 1) * p.W + min(px * 64 + (j & 1) * 32, p.W - 1);
                    if (pl == 0) a = p.cx + (int64_t)nn * p.coord_bs + o;
                    else if (pl == 1) a = p.cy + (int64_t)nn * p.coord_bs + o;
                    else a = p.data + (int64_t)nn * p.data_bs + ((int64_t)(pl - 2) * HW + o);
                } else {
                    const int j = idx - kLines, which = j / kRows;
                    const int o = min(py * kRows + (j - which * kRows), p.H - 1) * p.W + min(px * 64, p.W - 1);
                    if (which == 0 && HAS_BAD) a = p.data_mask + (int64_t)nn * p.data_mask_bs + o;
                    if (which == 1 && p.flow_mask) a = p.flow_mask + (int64_t)nn * p.flow_mask_bs + o;
                }
                if (a) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
    AccVec<KP> carry;                                    // south-tap sums waiting for the next source row
#pragma unroll
    for (int j = 0; j < KP; ++j) carry.v[j] = 0.0f;
    int carry_o = 0;
    bool carry_ok = false;

    // software pipeline: the loads of source row r+1 are issued before row r is processed (the kernel is bound by
    // memory latency at 3 CTAs/SM, and the volatile REDs below would otherwise pin the next row's loads behind them)
    struct RowIn { float fx, fy, d[C > 0 ? C : 1]; uint8_t dm, fmk; };
    auto load_row = [&](int r) {
        RowIn in;
        const int pix = (y0 + r) * p.W + xc;
        in.fx = __ldg(cx + pix);
        in.fy = __ldg(cy + pix);
#pragma unroll
        for (int c = 0; c < C; ++c) in.d[c] = __ldg(data + (c * HW + pix));
        in.dm = HAS_BAD ? __ldg(dmask + pix) : 1;
        in.fmk = fmask ? __ldg(fmask + pix) : 1;
        return in;
    };
    RowIn nxt = load_row(0);
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const RowIn cur = nxt;
        if (r + 1 < rows) nxt = load_row(r + 1);
        const int y = y0 + r;
        const float fx = cur.fx, fy = cur.fy;
        float X = fx, Y = fy;
        bool fm = active;
        if (p.relative) {
            X = __fadd_rn(__fmul_rn(p.sign, fx), (float)(x + p.src_x0));     // utils.py:1056-1057
            Y = __fadd_rn(__fmul_rn(p.sign, fy), (float)(y + p.src_y0));
            if (p.occlude_zero) fm = fm && !is_zero_vec(fx, fy);             // utils.py:1191-1192
        }
        fm = fm && (cur.fmk != 0);
        // utils.py:1098-1111: floor, neighbours, weights times the in-bounds test (saturating conversion: far away
        // and NaN positions fail the range tests / the w > 0 test below)
        const float x0f = floorf(X), y0f = floorf(Y);
        const int xi = __float2int_rd(X), yi = __float2int_rd(Y);
        const int ly = yi - p.acc_y0, lx = xi - p.src_x0;                   // position inside the stored accumulator
        const bool bx0 = (unsigned)xi < (unsigned)p.gw && (unsigned)lx < (unsigned)p.W;
        const bool bx1 = (unsigned)(xi + 1) < (unsigned)p.gw && (unsigned)(lx + 1) < (unsigned)p.W;
        const bool by0 = (unsigned)yi < (unsigned)p.gh && (unsigned)ly < (unsigned)p.acc_h;
        const bool by1 = (unsigned)(yi + 1) < (unsigned)p.gh && (unsigned)(ly + 1) < (unsigned)p.acc_h;
        const float ax0 = __fsub_rn(__fadd_rn(x0f, 1.0f), X), ax1 = __fsub_rn(X, x0f);
        const float ay0 = __fsub_rn(__fadd_rn(y0f, 1.0f), Y), ay1 = __fsub_rn(Y, y0f);
        const float w_nw = (fm && bx0 && by0) ? __fmul_rn(ay0, ax0) : 0.0f;   // utils.py:1114, 1123
        const float w_ne = (fm && bx1 && by0) ? __fmul_rn(ay0, ax1) : 0.0f;
        const float w_sw = (fm && bx0 && by1) ? __fmul_rn(ay1, ax0) : 0.0f;
        const float w_se = (fm && bx1 && by1) ? __fmul_rn(ay1, ax1) : 0.0f;
        const bool p_nw = w_nw > 0.0f, p_ne = w_ne > 0.0f, p_sw = w_sw > 0.0f, p_se = w_se > 0.0f;   // adding +0 is a no-op
        const int o = (int)((unsigned)ly * (unsigned)p.W + (unsigned)lx);

        const float *d = cur.d;
        const float badf = (HAS_BAD && cur.dm == 0) ? 1.0f : 0.0f;

        // per-tap accumulator vectors: [w*d_0 .. w*d_{C-1}, w, w*bad]   (utils.py:1133, 1139-1140)
        AccVec<KP> top, bot, ene, ese;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const float m = (j < C) ? d[j < C ? j : 0] : (j == C ? 1.0f : ((BAD_SLOT && j == C + 1) ? badf : 0.0f));
            top.v[j] = __fmul_rn(w_nw, m); ene.v[j] = __fmul_rn(w_ne, m);
            bot.v[j] = __fmul_rn(w_sw, m); ese.v[j] = __fmul_rn(w_se, m);
        }

        if constexpr (SPLIT) {
            if (badf != 0.0f) {                          // few sources carry bad weight: plain scalar REDs, no merging
                if (p_nw) atomicAdd(accb + o, w_nw);
                if (p_ne) atomicAdd(accb + (o + 1), w_ne);
                if (p_sw) atomicAdd(accb + (o + p.W), w_sw);
                if (p_se) atomicAdd(accb + (o + p.W + 1), w_se);
            }
        }

        // ---- horizontal merge with lane-1
        const int prev_o = __shfl_up_sync(0xffffffffu, o, 1);
        const unsigned pe = (p_ne ? 1u : 0u) | (p_se ? 2u : 0u);
        const unsigned prev_pe = __shfl_up_sync(0xffffffffu, pe, 1);
        const bool chain = (lane > 0) && (prev_o + 1 == o);
        const bool merge_n = chain && (prev_pe & 1u) && p_nw;
        const bool merge_s = chain && (prev_pe & 2u) && p_sw;
        const unsigned bn = __ballot_sync(0xffffffffu, merge_n), bs = __ballot_sync(0xffffffffu, merge_s);
        const bool taken_n = (lane < 31) && ((bn >> (lane + 1)) & 1u);
        const bool taken_s = (lane < 31) && ((bs >> (lane + 1)) & 1u);
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            if (j >= C + 1 + (BAD_SLOT ? 1 : 0)) continue;                 // padding slots stay zero
            const float pn = __shfl_up_sync(0xffffffffu, ene.v[j], 1), ps = __shfl_up_sync(0xffffffffu, ese.v[j], 1);
            if (merge_n) top.v[j] += pn;
            if (merge_s) bot.v[j] += ps;
        }
        // ---- vertical merge with the carried south taps of the previous source row
        const bool vmerge = carry_ok && p_nw && (carry_o == o);
        red_vec<KP>(acc, carry_o, carry, carry_ok && !vmerge);
        if (vmerge) {
#pragma unroll
            for (int j = 0; j < KP; ++j) top.v[j] += carry.v[j];
        }
        red_vec<KP>(acc, o, top, p_nw);
        red_vec<KP>(acc, o + 1, ene, p_ne && !taken_n);
        red_vec<KP>(acc, o + p.W + 1, ese, p_se && !taken_s);
        carry = bot;
        carry_o = o + p.W;
        carry_ok = p_sw;
    }
    red_vec<KP>(acc, carry_o, carry, carry_ok);
}

// normalise + masks + un-occlude fill from the interleaved accumulator.  Same results as splat_normalise_kernel for
// every mask bit; the payload is multiplied by the correctly rounded reciprocal of the clamped density instead of
// divided by it (<= 1 ulp apart, against a parity tolerance of 1e-4) and the mask-channel division only runs where
// its outcome is not known beforehand.  One thread walks kNormRows rows of one column so that the per-item pointer
// arithmetic is paid once, not per pixel.
constexpr int kNormRows = 4;       // CTA tile = 64 x 16 (2 warps across x, 4 across y)

template <int C, bool HAS_BAD, int KP, bool SPLIT = false>
__global__ void __launch_bounds__(256) splat_normalise_v1_kernel(const SplatParams p, const int n0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kNormRows) + (warp >> 1) * kNormRows;
    const int n = n0 + blockIdx.z;
    if (x >= p.W || y0 >= p.H) return;
    const int rows = min(kNormRows, p.H - y0);
    const int HW = p.H * p.W;
    const float *ab = p.acc + ((int64_t)blockIdx.z * p.acc_h * p.W + (int64_t)(y0 + p.src_y0 - p.acc_y0) * p.W + x) * KP;
    const float *cx = p.cx + (int64_t)n * p.coord_bs, *cy = p.cy + (int64_t)n * p.coord_bs;
    const float *data = (C > 0) ? p.data + (int64_t)n * p.data_bs : nullptr;
    const uint8_t *dmask = HAS_BAD ? p.data_mask + (int64_t)n * p.data_mask_bs : nullptr;
    const uint8_t *fmask = p.flow_mask ? p.flow_mask + (int64_t)n * p.flow_mask_bs : nullptr;
    float *out = (C > 0) ? p.out + (int64_t)n * C * HW : nullptr;
    float *density = p.density ? p.density + (int64_t)n * HW : nullptr;
    uint8_t *wmask = p.warped_mask ? p.warped_mask + (int64_t)n * HW : nullptr;
    uint8_t *unocc_o = p.unocc ? p.unocc + (int64_t)n * HW : nullptr;
    uint8_t *valid = p.valid ? p.valid + (int64_t)n * HW : nullptr;
    const float *abad = SPLIT ? p.acc_bad + (int64_t)blockIdx.z * p.acc_h * p.W +
                                    (int64_t)(y0 + p.src_y0 - p.acc_y0) * p.W + x : nullptr;
    asm volatile("" : "+l"(ab), "+l"(abad), "+l"(cx), "+l"(cy), "+l"(data), "+l"(dmask), "+l"(fmask));
    asm volatile("" : "+l"(out), "+l"(density), "+l"(wmask), "+l"(unocc_o), "+l"(valid));
    const bool occ = p.occlude_zero && p.relative;
    const bool eq1 = p.valid_mode == OFB_VALID_EQ1;

    float a[kNormRows][KP], sb[kNormRows];
#pragma unroll
    for (int r = 0; r < kNormRows; ++r) {
        if (r >= rows) break;
        const float *ar = ab + r * (p.W * KP);
        if constexpr (SPLIT) sb[r] = __ldcg(abad + r * p.W);
        if constexpr (KP == 1) {
            a[r][0] = __ldcg(ar);
        } else if constexpr (KP == 2) {
            const float2 q = __ldcg(reinterpret_cast<const float2 *>(ar));
            a[r][0] = q.x; a[r][1] = q.y;
        } else {
#pragma unroll
            for (int h = 0; h < KP / 4; ++h) {
                const float4 q = __ldcg(reinterpret_cast<const float4 *>(ar) + h);
                a[r][4 * h] = q.x; a[r][4 * h + 1] = q.y; a[r][4 * h + 2] = q.z; a[r][4 * h + 3] = q.w;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kNormRows; ++r) {
        if (r >= rows) break;
        const int pix = (y0 + r) * p.W + x;
        const float den = a[r][C];
        const float bad = SPLIT ? sb[r] : (HAS_BAD ? a[r][C + 1] : 0.0f);
        const float dc = fmaxf(den, OFB_MIN_DENSITY);
        const bool wm = den > 0.0f;
        bool unocc = false;
        if (occ && !wm) {
            unocc = is_zero_vec(__ldg(cx + pix), __ldg(cy + pix));
            if (fmask) unocc = unocc && (__ldg(fmask + pix) != 0);
        }
        if (C > 0) {
            const float inv = __frcp_rn(dc);
#pragma unroll
            for (int c = 0; c < C; ++c)
                out[c * HW + pix] = unocc ? __ldg(data + (c * HW + pix)) : __fmul_rn(a[r][c], inv);
        }
        if (density) density[pix] = den;
        if (wmask) wmask[pix] = wm ? 1 : 0;
        if (unocc_o) unocc_o[pix] = unocc ? 1 : 0;
        if (valid) {
            // mask channel = (den - bad) / max(den, 1e-3)   (utils.py:1144 applied to the mask plane)
            float mch;
            if (unocc) {
                mch = (HAS_BAD && __ldg(dmask + pix) == 0) ? 0.0f : 1.0f;
            } else if (bad == 0.0f) {
                // den / max(den, 1e-3): exactly 1 from 1e-3 up, exactly 0 at 0, a real quotient only in between
                mch = (den >= OFB_MIN_DENSITY) ? 1.0f : (wm ? __fdiv_rn(den, OFB_MIN_DENSITY) : 0.0f);
            } else {
                mch = __fdiv_rn(__fsub_rn(den, bad), dc);
            }
            valid[pix] = (eq1 ? (mch == 1.0f) : (mch > 0.99999f)) ? 1 : 0;
        }
    }
}

typedef void (*SplatKernel)(const SplatParams, const int);
// kp floats per pixel in the vectors; split != 0: plus one plane for the bad-weight channel (acc_s / norm_s kernels)
struct SplatV1 { SplatKernel acc, norm; int kp; SplatKernel acc_s, norm_s; int kp_s; };

template <int C>
static SplatV1 splat_v1_for(bool bad)
{
    constexpr int KPb = (C + 2 <= 2) ? 2 : ((C + 2 <= 4) ? 4 : 8);
    constexpr int KPn = (C + 1 <= 1) ? 1 : ((C + 1 <= 2) ? 2 : ((C + 1 <= 4) ? 4 : 8));
    if (bad) {
        SplatV1 k{splat_accumulate_v1_kernel<C, true, KPb>, splat_normalise_v1_kernel<C, true, KPb>, KPb,
                  nullptr, nullptr, 0};
        if constexpr (KPn < KPb) {       // the bad channel alone would double the vector: give it its own plane
            k.acc_s = splat_accumulate_v1_kernel<C, true, KPn, true>;
            k.norm_s = splat_normalise_v1_kernel<C, true, KPn, true>;
            k.kp_s = KPn;
        }
        return k;
    }
    return {splat_accumulate_v1_kernel<C, false, KPn>, splat_normalise_v1_kernel<C, false, KPn>, KPn,
            nullptr, nullptr, 0};
}

static bool splat_v1(int C, bool bad, SplatV1 *k)
{
    switch (C) {
        case 0: *k = splat_v1_for<0>(bad); return true;
        case 1: *k = splat_v1_for<1>(bad); return true;
        case 2: *k = splat_v1_for<2>(bad); return true;
        case 3: *k = splat_v1_for<3>(bad); return true;
        case 4: *k = splat_v1_for<4>(bad); return true;
        case 5: *k = splat_v1_for<5>(bad); return true;
        case 6: *k = splat_v1_for<6>(bad); return true;
        default: return false;                          // wider payloads take the planar kernels
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
    int src_y0, src_x0, gh, gw;      // window frame as in SplatParams (v1 kernels only; identity = 0, 0, H, W)
    int acc_y0, acc_h;               // rows of the frame that G stores (window path; padded apply: src_y0, H)
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

// ------------------------------------------------------------------------------------------------
// v1 adjoint: per-target prepass + vector tap gathers
// ------------------------------------------------------------------------------------------------
// The v0 gather evaluates gnum_c = g'_c/dc and gden at every tap: 2C+3 scalar gathers per tap, 4 taps per source.
// Here a prepass writes G[p] = {gnum_0 .. gnum_{C-1}, gden} once per TARGET pixel, interleaved (KG = 2/4/8 floats),
// and the per-source kernel fetches each tap with one 64/128-bit load.  
constexpr int kPreRows = 4;        // rows per thread; CTA tile = 64 x 16 (2 warps across x, 4 across y)

template <int C, int KG>
__global__ void __launch_bounds__(256) splat_bwd_prepass_kernel(const SplatBwdParams p, float *__restrict__ G)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kPreRows) + (warp >> 1) * kPreRows;
    const int n = blockIdx.z;
    if (x >= p.W || y0 >= p.H) return;
    const int rows = min(kPreRows, p.H - y0);
    const int HW = p.H * p.W;
    const float *den = p.density + (int64_t)n * HW;
    const uint8_t *uno = p.unocc ? p.unocc + (int64_t)n * HW : nullptr;
    const float *go = p.gout + (int64_t)n * C * HW, *ou = p.out + (int64_t)n * C * HW;
    const float *gd = p.gdensity ? p.gdensity + (int64_t)n * HW : nullptr;
    float *gp = G + (int64_t)n * HW * KG;
    asm volatile("" : "+l"(den), "+l"(uno), "+l"(go), "+l"(ou), "+l"(gd), "+l"(gp));

    // all loads of the column first (a streaming kernel lives on bytes in flight), arithmetic after
    float dq[kPreRows], g[kPreRows][C > 0 ? C : 1], o[kPreRows][C > 0 ? C : 1], gden[kPreRows];
    bool dead[kPreRows];
#pragma unroll
    for (int r = 0; r < kPreRows; ++r) {
        if (r >= rows) break;
        const int pix = (y0 + r) * p.W + x;
        dq[r] = __ldg(den + pix);
        dead[r] = uno && (__ldg(uno + pix) != 0);                   // utils.py:1203 overwrote it
        gden[r] = gd ? __ldg(gd + pix) : 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            g[r][c] = __ldg(go + (c * HW + pix));
            o[r][c] = __ldg(ou + (c * HW + pix));
        }
    }
#pragma unroll
    for (int r = 0; r < kPreRows; ++r) {
        if (r >= rows) break;
        const int pix = (y0 + r) * p.W + x;
        const float dc = fmaxf(dq[r], OFB_MIN_DENSITY);
        float v[KG];
#pragma unroll
        for (int j = 0; j < KG; ++j) v[j] = 0.0f;
        float s = 0.0f;
        if (!dead[r]) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                v[c] = g[r][c] / dc;
                s += g[r][c] * o[r][c];
            }
        }
        float gdn = gden[r];
        if (!dead[r] && dq[r] >= OFB_MIN_DENSITY) gdn -= s / dc;    // clamp_min passes the gradient where den >= 1e-3
        v[C] = gdn;
        float *gq = gp + (int64_t)pix * KG;
        if constexpr (KG == 2) {
            *reinterpret_cast<float2 *>(gq) = make_float2(v[0], v[1]);
        } else {
#pragma unroll
            for (int h = 0; h < KG / 4; ++h)
                reinterpret_cast<float4 *>(gq)[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
    }
}

template <int KG>
__device__ __forceinline__ void load_g(const float *__restrict__ G, int q, float (&v)[KG])
{
    const float *gp = G + (int64_t)q * KG;
    if constexpr (KG == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(gp));
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int h = 0; h < KG / 4; ++h) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(gp) + h);
            v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
        }
    }
}

template <int C, int KG, bool WIN>
__global__ void __launch_bounds__(256) splat_bwd_gather_kernel(const SplatBwdParams p, const float *__restrict__ G)
{
    // WIN = false (compile time): the tensors ARE the frame, no origin shift and no second range test
    const int fx0 = WIN ? p.src_x0 : 0, fy0 = WIN ? p.src_y0 : 0, gw = WIN ? p.gw : p.W, gh = WIN ? p.gh : p.H;
    const int gy0 = WIN ? p.acc_y0 : 0, grows = WIN ? p.acc_h : p.H;   // rows of the frame held by G
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kPreRows) + (warp >> 1) * kPreRows;
    const int n = blockIdx.z;
    if (x >= p.W || y0 >= p.H) return;
    const int rows = min(kPreRows, p.H - y0);
    const int HW = p.H * p.W;
    const float *cx = p.cx + (int64_t)n * p.coord_bs, *cy = p.cy + (int64_t)n * p.coord_bs;
    const uint8_t *fmask = p.flow_mask ? p.flow_mask + (int64_t)n * p.flow_mask_bs : nullptr;
    const float *data = p.data + (int64_t)n * p.data_bs;
    const float *Gn = G + (int64_t)n * grows * p.W * KG;
    const uint8_t *uno = (p.gdata && p.unocc) ? p.unocc + (int64_t)n * HW : nullptr;
    const float *go = p.gout + (int64_t)n * C * HW;
    float *gdp = p.gdata ? p.gdata + (int64_t)n * C * HW : nullptr;
    float *gcx = p.gcx ? p.gcx + (int64_t)n * p.gcoord_bs : nullptr, *gcy = p.gcx ? p.gcy + (int64_t)n * p.gcoord_bs : nullptr;
    asm volatile("" : "+l"(cx), "+l"(cy), "+l"(fmask), "+l"(data), "+l"(Gn), "+l"(uno), "+l"(go), "+l"(gdp), "+l"(gcx), "+l"(gcy));

    // the streamed inputs of the whole column first, then the dependent tap gathers row by row
    float fxs[kPreRows], fys[kPreRows], dats[kPreRows][C];
    uint8_t fms[kPreRows], uns[kPreRows];
#pragma unroll
    for (int r = 0; r < kPreRows; ++r) {
        if (r >= rows) break;
        const int pix = (y0 + r) * p.W + x;
        fxs[r] = __ldg(cx + pix);
        fys[r] = __ldg(cy + pix);
        fms[r] = fmask ? __ldg(fmask + pix) : 1;
        uns[r] = uno ? __ldg(uno + pix) : 0;
#pragma unroll
        for (int c = 0; c < C; ++c) dats[r][c] = __ldg(data + (c * HW + pix));
    }
#pragma unroll
    for (int r = 0; r < kPreRows; ++r) {
        if (r >= rows) break;
        const int y = y0 + r, pix = y * p.W + x;
        const float fx = fxs[r], fy = fys[r];
        float X = fx, Y = fy;
        bool fm = true;
        if (p.relative) {
            X = __fadd_rn(__fmul_rn(p.sign, fx), (float)(x + fx0));
            Y = __fadd_rn(__fmul_rn(p.sign, fy), (float)(y + fy0));
            if (p.occlude_zero) fm = !is_zero_vec(fx, fy);
        }
        fm = fm && (fms[r] != 0);

        float gd[C], gx = 0.0f, gy = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) gd[c] = 0.0f;
        if (fm) {
            const float *dat = dats[r];
            const float x0f = floorf(X), y0f = floorf(Y);
            const int xi = __float2int_rd(X), yi = __float2int_rd(Y);
            // in-frame tests decide the weights (utils.py:1110-1111); a tap outside the stored window exists in the
            // frame but was cut away with its gradient, so it contributes nothing
            const bool bx[2] = {(unsigned)xi < (unsigned)gw, (unsigned)(xi + 1) < (unsigned)gw};
            const bool by[2] = {(unsigned)yi < (unsigned)gh, (unsigned)(yi + 1) < (unsigned)gh};
            const int lx = xi - fx0, ly = yi - gy0;
            const bool sx[2] = {!WIN || (unsigned)lx < (unsigned)p.W, !WIN || (unsigned)(lx + 1) < (unsigned)p.W};
            const bool sy[2] = {!WIN || (unsigned)ly < (unsigned)grows, !WIN || (unsigned)(ly + 1) < (unsigned)grows};
            const float ax[2] = {__fsub_rn(__fadd_rn(x0f, 1.0f), X), __fsub_rn(X, x0f)};
            const float ay[2] = {__fsub_rn(__fadd_rn(y0f, 1.0f), Y), __fsub_rn(Y, y0f)};
            const float wx[2] = {bx[0] ? ax[0] : 0.0f, bx[1] ? ax[1] : 0.0f};      // utils.py:1110
            const float wy[2] = {by[0] ? ay[0] : 0.0f, by[1] ? ay[1] : 0.0f};      // utils.py:1111
            const int o = ly * p.W + lx;
            float Gv[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    Gv[a][b] = 0.0f;
                    if (!(bx[b] && by[a] && sx[b] && sy[a])) continue;          // contributes to nothing
                    float g[KG];
                    load_g<KG>(Gn, o + a * p.W + b, g);
                    const float w = __fmul_rn(wy[a], wx[b]);
                    float acc = g[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        acc = __fmaf_rn(g[c], dat[c], acc);
                        gd[c] = __fmaf_rn(w, g[c], gd[c]);
                    }
                    Gv[a][b] = acc;
                }
            const float ibx0 = bx[0] ? 1.0f : 0.0f, ibx1 = bx[1] ? 1.0f : 0.0f;
            const float iby0 = by[0] ? 1.0f : 0.0f, iby1 = by[1] ? 1.0f : 0.0f;
            // d w / d x: wx[0] = (x1 - x)*inbx[0] -> -inbx[0];  wx[1] = (x - x0)*inbx[1] -> +inbx[1]
            gx = wy[0] * (ibx1 * Gv[0][1] - ibx0 * Gv[0][0]) + wy[1] * (ibx1 * Gv[1][1] - ibx0 * Gv[1][0]);
            gy = wx[0] * (iby1 * Gv[1][0] - iby0 * Gv[0][0]) + wx[1] * (iby1 * Gv[1][1] - iby0 * Gv[0][1]);
            if (p.relative) {
                gx *= p.sign;
                gy *= p.sign;
            }
        }
        if (gdp) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float v = gd[c];
                if (uns[r]) v += __ldg(go + (c * HW + pix));
                gdp[c * HW + pix] = v;
            }
        }
        if (gcx) {
            gcx[pix] = gx;
            gcy[pix] = gy;
        }
    }
}

typedef void (*SplatBwdPre)(const SplatBwdParams, float *);
typedef void (*SplatBwdGat)(const SplatBwdParams, const float *);
struct SplatBwdV1 { SplatBwdPre pre; SplatBwdGat gat, gat_win; int kg; };

template <int C>
static SplatBwdV1 splat_bwd_v1_for()
{
    constexpr int KG = (C + 1 <= 2) ? 2 : ((C + 1 <= 4) ? 4 : 8);
    return {splat_bwd_prepass_kernel<C, KG>, splat_bwd_gather_kernel<C, KG, false>, splat_bwd_gather_kernel<C, KG, true>, KG};
}

static bool splat_bwd_v1(int C, SplatBwdV1 *k)
{
    switch (C) {
        case 1: *k = splat_bwd_v1_for<1>(); return true;
        case 2: *k = splat_bwd_v1_for<2>(); return true;
        case 3: *k = splat_bwd_v1_for<3>(); return true;
        case 4: *k = splat_bwd_v1_for<4>(); return true;
        case 5: *k = splat_bwd_v1_for<5>(); return true;
        case 6: *k = splat_bwd_v1_for<6>(); return true;
        case 7: *k = splat_bwd_v1_for<7>(); return true;
        default: return false;
    }
}

}  // namespace ofb

using namespace ofb;

// layout of the vector-RED path for this problem: floats per pixel in the vectors and whether the bad-weight channel
// sits in a plane of its own behind them
static void splat_v1_layout(const SplatV1 &k, int *kp, bool *split)
{
    static const bool no_split = getenv("OFB_SPLAT_NOSPLIT") != nullptr;
    *split = k.acc_s != nullptr && !no_split;
    *kp = *split ? k.kp_s : k.kp;
}

static int64_t splat_v1_chunk(int64_t item_floats, int N)
{
    static const long chunk_mb = getenv("OFB_SPLAT_CHUNK_MB") ? atol(getenv("OFB_SPLAT_CHUNK_MB")) : 1024;
    int64_t chunk = ((int64_t)chunk_mb << 20) / (item_floats * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > N) chunk = N;
    return chunk;
}

// floats of workspace ofb_splat_s_fwd wants for this problem: one chunk of interleaved accumulators for the
// vector-RED path (C <= 6), C+2 planes per item for the planar path
extern "C" int64_t ofb_splat_s_workspace(int N, int C, int H, int W, int has_data_mask)
{
    SplatV1 k;
    if (N <= 0 || C < 0 || H <= 0 || W <= 0) return 0;
    if (g_splat_det) return splat_det_floats(N, C, H, W);        // also covers the fast path (it needs less)
    if (getenv("OFB_SPLAT_PLANAR") == nullptr && splat_v1(C, has_data_mask != 0, &k)) {
        int kp;
        bool split;
        splat_v1_layout(k, &kp, &split);
        const int64_t item_floats = (int64_t)H * W * (kp + (split ? 1 : 0));
        return splat_v1_chunk(item_floats, N) * item_floats;
    }
    return (int64_t)N * (C + 2) * H * W;
}

// win = {src_y0, src_x0, gh, gw}: the H x W tensors are the window at (src_y0, src_x0) of a gh x gw frame in which
// the end points are computed and tested; what lands outside the window is dropped
static int splat_fwd_impl(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                          const float *data, int64_t data_bs,
                          const uint8_t *data_mask, int64_t data_mask_bs,
                          const uint8_t *flow_mask, int64_t flow_mask_bs,
                          int occlude_zero, float *acc, int64_t acc_floats,
                          float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                          uint8_t *valid, int valid_mode,
                          int N, int C, int H, int W, const int *win, void *stream)
{
    if (!cx || !cy || !acc || N <= 0 || C < 0 || (C > 0 && (!data || !out))) return OFB_EINVAL;
    const bool windowed = !(win[0] == 0 && win[1] == 0 && win[2] == H && win[3] == W);
    if (win[0] < 0 || win[1] < 0 || win[0] + H > win[2] || win[1] + W > win[3]) return OFB_EINVAL;
    if (windowed && !relative) return OFB_EINVAL;
    if (occlude_zero && !relative) return OFB_EINVAL;
    if (valid && valid_mode != OFB_VALID_GT && valid_mode != OFB_VALID_EQ1) return OFB_EINVAL;
    if (H < 1 || W < 1) return OFB_ESHAPE;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplatParams p{cx, cy, coord_bs, relative, sign, data, data_bs, data_mask, data_mask_bs,
                  flow_mask, flow_mask_bs, occlude_zero, acc, out, density, warped_mask, unocc,
                  valid, valid_mode, N, C, H, W, win[0], win[2], win[0], H, nullptr, win[1], win[3]};
    if (g_splat_det && !windowed && acc_floats >= splat_det_floats(N, C, H, W) && aligned16(acc)) {
        // deterministic merge: max|data| -> fixed-point accumulate -> normalise
        unsigned long long *acc64 = reinterpret_cast<unsigned long long *>(acc);
        unsigned *maxbits = reinterpret_cast<unsigned *>(acc + (int64_t)N * (C + 2) * H * W * 2);
        cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)splat_det_floats(N, C, H, W), st);
        if (e != cudaSuccess) return (int)e;
        if (C > 0) {
            const int64_t count = (int64_t)(data_bs == 0 ? 1 : N) * C * H * W;
            absmax_kernel<<<148 * 8, 256, 0, st>>>(data, count, maxbits);
            OFB_CHECK_LAUNCH();
        }
        const dim3 block(32, 8);
        const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
        splat_accumulate_det_kernel<<<grid, block, 0, st>>>(p, acc64, maxbits);
        OFB_CHECK_LAUNCH();
        splat_normalise_det_kernel<<<grid, block, 0, st>>>(p, acc64, maxbits);
        OFB_CHECK_LAUNCH();
        return 0;
    }
    SplatV1 k;
    static const bool planar_only = getenv("OFB_SPLAT_PLANAR") != nullptr;
    if (!planar_only && splat_v1(C, data_mask != nullptr, &k)) {
        // chunk = as many items as the chunk budget and the workspace the caller provides allow, at least one
        int kp;
        bool split;
        splat_v1_layout(k, &kp, &split);
        const int64_t item_floats = (int64_t)H * W * (kp + (split ? 1 : 0));
        int64_t chunk = splat_v1_chunk(item_floats, N);
        const int64_t fit = acc_floats / item_floats;
        if (chunk > fit) chunk = fit;
        if (fit >= 1) {
            const SplatKernel acc_k = split ? k.acc_s : k.acc, norm_k = split ? k.norm_s : k.norm;
            p.acc_bad = split ? acc + chunk * (int64_t)H * W * kp : nullptr;   // planes behind the chunk's vectors
            const dim3 agrid((W + 63) / 64, (H + 4 * kSplatRows - 1) / (4 * kSplatRows), 1);
            for (int n0 = 0; n0 < N; n0 += (int)chunk) {
                const int nn = (N - n0 < chunk) ? (N - n0) : (int)chunk;
                cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)item_floats * chunk, st);
                if (e != cudaSuccess) return (int)e;
                acc_k<<<dim3(agrid.x, agrid.y, nn), 256, 0, st>>>(p, n0);
                OFB_CHECK_LAUNCH();
                norm_k<<<dim3((W + 63) / 64, (H + 4 * kNormRows - 1) / (4 * kNormRows), nn), 256, 0, st>>>(p, n0);
                OFB_CHECK_LAUNCH();
            }
            return 0;
        }
    }
    if (windowed) return OFB_EINVAL;                                       // the planar kernels know no window
    if (acc_floats < (int64_t)N * (C + 2) * H * W) return OFB_EINVAL;     // planar path needs C+2 planes per item
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)N * (C + 2) * H * W, st);
    if (e != cudaSuccess) return (int)e;
    const dim3 block(32, 8);
    const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    splat_accumulate_kernel<<<grid, block, 0, st>>>(p);
    OFB_CHECK_LAUNCH();
    splat_normalise_kernel<<<grid, block, 0, st>>>(p);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_splat_s_fwd(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                               const float *data, int64_t data_bs,
                               const uint8_t *data_mask, int64_t data_mask_bs,
                               const uint8_t *flow_mask, int64_t flow_mask_bs,
                               int occlude_zero, float *acc, int64_t acc_floats,
                               float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                               uint8_t *valid, int valid_mode,
                               int N, int C, int H, int W, void *stream)
{
    const int win[4] = {0, 0, H, W};
    return splat_fwd_impl(cx, cy, coord_bs, relative, sign, data, data_bs, data_mask, data_mask_bs, flow_mask,
                          flow_mask_bs, occlude_zero, acc, acc_floats, out, density, warped_mask, unocc, valid,
                          valid_mode, N, C, H, W, win, stream);
}

extern "C" int ofb_splat_s_win_fwd(const float *cx, const float *cy, int64_t coord_bs, float sign,
                                   const float *data, int64_t data_bs,
                                   const uint8_t *data_mask, int64_t data_mask_bs,
                                   const uint8_t *flow_mask, int64_t flow_mask_bs,
                                   int occlude_zero, float *acc, int64_t acc_floats,
                                   float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                                   uint8_t *valid, int valid_mode,
                                   int N, int C, int H, int W, const int *window, void *stream)
{
    if (!window) return OFB_EINVAL;
    return splat_fwd_impl(cx, cy, coord_bs, 1, sign, data, data_bs, data_mask, data_mask_bs, flow_mask,
                          flow_mask_bs, occlude_zero, acc, acc_floats, out, density, warped_mask, unocc, valid,
                          valid_mode, N, C, H, W, window, stream);
}

// floats of workspace ofb_splat_s_bwd wants (0: none, the generic gather is used)
extern "C" int64_t ofb_splat_s_bwd_workspace(int N, int C, int H, int W)
{
    SplatBwdV1 k;
    if (N <= 0 || H <= 0 || W <= 0 || !splat_bwd_v1(C, &k)) return 0;
    return (int64_t)N * H * W * k.kg;
}

static int splat_bwd_impl(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                          const float *data, int64_t data_bs,
                          const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                          const float *out, const float *density, const uint8_t *unocc,
                          const float *gout, const float *gdensity,
                          float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                          float *work, int64_t work_floats,
                          int N, int C, int H, int W, const int *win, void *stream)
{
    if (!cx || !cy || !density || N <= 0 || C < 0) return OFB_EINVAL;
    const bool windowed = !(win[0] == 0 && win[1] == 0 && win[2] == H && win[3] == W);
    if (win[0] < 0 || win[1] < 0 || win[0] + H > win[2] || win[1] + W > win[3]) return OFB_EINVAL;
    if (windowed && !relative) return OFB_EINVAL;
    if (C > 0 && (!data || !out || !gout)) return OFB_EINVAL;
    if ((gcx == nullptr) != (gcy == nullptr)) return OFB_EINVAL;
    if (H < 1 || W < 1) return OFB_ESHAPE;
    if ((int64_t)H * W > 0x3fffffff || N > 65535) return OFB_EINVAL;
    if (!gcx && !gdata) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplatBwdParams p{cx, cy, coord_bs, relative, sign, data, data_bs, flow_mask, flow_mask_bs, occlude_zero,
                     out, density, unocc, gout, gdensity, gcx, gcy, gcoord_bs, gdata, N, C, H, W,
                     win[0], win[1], win[2], win[3], win[0], H};
    const dim3 block(32, 8);
    const dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    SplatBwdV1 k;
    if (work && splat_bwd_v1(C, &k) && work_floats >= (int64_t)N * H * W * k.kg) {
        k.pre<<<dim3((W + 63) / 64, (H + 4 * kPreRows - 1) / (4 * kPreRows), N), 256, 0, st>>>(p, work);
        OFB_CHECK_LAUNCH();
        (windowed ? k.gat_win : k.gat)<<<dim3((W + 63) / 64, (H + 4 * kPreRows - 1) / (4 * kPreRows), N), 256, 0, st>>>(p, work);
        OFB_CHECK_LAUNCH();
        return 0;
    }
    if (windowed) return OFB_EINVAL;                  // the generic gather knows no window
    int c0 = 0;
    do {
        const int cn = (C - c0 < kMaxC) ? (C - c0) : kMaxC;
        splat_bwd_kernel<<<grid, block, 0, st>>>(p, c0, cn, c0 == 0);
        OFB_CHECK_LAUNCH();
        c0 += kMaxC;
    } while (c0 < C);
    return 0;
}

extern "C" int ofb_splat_s_bwd(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                               const float *data, int64_t data_bs,
                               const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                               const float *out, const float *density, const uint8_t *unocc,
                               const float *gout, const float *gdensity,
                               float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                               float *work, int64_t work_floats,
                               int N, int C, int H, int W, void *stream)
{
    const int win[4] = {0, 0, H, W};
    return splat_bwd_impl(cx, cy, coord_bs, relative, sign, data, data_bs, flow_mask, flow_mask_bs, occlude_zero, out,
                          density, unocc, gout, gdensity, gcx, gcy, gcoord_bs, gdata, work, work_floats, N, C, H, W,
                          win, stream);
}

extern "C" int ofb_splat_s_win_bwd(const float *cx, const float *cy, int64_t coord_bs, float sign,
                                   const float *data, int64_t data_bs,
                                   const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                                   const float *out, const float *density, const uint8_t *unocc,
                                   const float *gout, const float *gdensity,
                                   float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                                   float *work, int64_t work_floats,
                                   int N, int C, int H, int W, const int *window, void *stream)
{
    if (!window) return OFB_EINVAL;
    return splat_bwd_impl(cx, cy, coord_bs, 1, sign, data, data_bs, flow_mask, flow_mask_bs, occlude_zero, out,
                          density, unocc, gout, gdensity, gcx, gcy, gcoord_bs, gdata, work, work_floats, N, C, H, W,
                          window, stream);
}

// ------------------------------------------------------------------------------------------------
// row-band entry points: accumulate and normalise separately, so that accumulator halos can be exchanged between
// the GPUs that share one oversized field (oflibpytorch_b200/sharding.py)
// ------------------------------------------------------------------------------------------------
extern "C" int ofb_splat_s_kp(int C, int has_data_mask)
{
    SplatV1 k;
    return splat_v1(C, has_data_mask != 0, &k) ? k.kp : 0;
}

// band[4] = {src_y0, global_h, acc_y0, acc_h}; acc = (N, acc_h, W, KP) floats with KP = ofb_splat_s_kp(C, mask),
// zero-filled by the call
extern "C" int ofb_splat_s_band_accumulate(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                                           const float *data, int64_t data_bs,
                                           const uint8_t *data_mask, int64_t data_mask_bs,
                                           const uint8_t *flow_mask, int64_t flow_mask_bs,
                                           int occlude_zero, float *acc, int N, int C, int H, int W, const int *band,
                                           void *stream)
{
    SplatV1 k;
    if (!cx || !cy || !acc || !band || N <= 0 || C < 0 || (C > 0 && !data) || H < 1 || W < 1) return OFB_EINVAL;
    if (occlude_zero && !relative) return OFB_EINVAL;
    if (!splat_v1(C, data_mask != nullptr, &k)) return OFB_EINVAL;
    const int src_y0 = band[0], gh = band[1], acc_y0 = band[2], acc_h = band[3];
    if (src_y0 < 0 || src_y0 + H > gh || acc_y0 < 0 || acc_h < 1 || acc_y0 + acc_h > gh || N > 65535 ||
        (int64_t)acc_h * W > 0x3fffffff)
        return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)N * acc_h * W * k.kp, st);
    if (e != cudaSuccess) return (int)e;
    SplatParams p{cx, cy, coord_bs, relative, sign, data, data_bs, data_mask, data_mask_bs, flow_mask, flow_mask_bs,
                  occlude_zero, acc, nullptr, nullptr, nullptr, nullptr, nullptr, 0, N, C, H, W, src_y0, gh, acc_y0, acc_h,
                  nullptr, 0, W};
    k.acc<<<dim3((W + 63) / 64, (H + 4 * kSplatRows - 1) / (4 * kSplatRows), N), 256, 0, st>>>(p, 0);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ofb_splat_s_band_normalise(const float *cx, const float *cy, int64_t coord_bs, int relative,
                                          const float *data, int64_t data_bs,
                                          const uint8_t *data_mask, int64_t data_mask_bs,
                                          const uint8_t *flow_mask, int64_t flow_mask_bs,
                                          int occlude_zero, const float *acc,
                                          float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                                          uint8_t *valid, int valid_mode,
                                          int N, int C, int H, int W, const int *band, void *stream)
{
    SplatV1 k;
    if (!cx || !cy || !acc || !band || N <= 0 || C < 0 || (C > 0 && (!data || !out)) || H < 1 || W < 1) return OFB_EINVAL;
    if (valid && valid_mode != OFB_VALID_GT && valid_mode != OFB_VALID_EQ1) return OFB_EINVAL;
    if (!splat_v1(C, data_mask != nullptr, &k)) return OFB_EINVAL;
    const int src_y0 = band[0], gh = band[1], acc_y0 = band[2], acc_h = band[3];
    if (src_y0 < acc_y0 || src_y0 + H > acc_y0 + acc_h || acc_y0 + acc_h > gh || N > 65535) return OFB_EINVAL;
    SplatParams p{cx, cy, coord_bs, relative, 1.0f, data, data_bs, data_mask, data_mask_bs, flow_mask, flow_mask_bs,
                  occlude_zero, const_cast<float *>(acc), out, density, warped_mask, unocc, valid, valid_mode,
                  N, C, H, W, src_y0, gh, acc_y0, acc_h, nullptr, 0, W};
    k.norm<<<dim3((W + 63) / 64, (H + 4 * kNormRows - 1) / (4 * kNormRows), N), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, 0);
    OFB_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// row-band adjoint: prepass on the rank's own target rows, then (after the neighbours' halo rows of G have been
// copied in) the gather for the rank's own sources against the extended G
// ------------------------------------------------------------------------------------------------
extern "C" int ofb_splat_s_kg(int C)
{
    SplatBwdV1 k;
    return splat_bwd_v1(C, &k) ? k.kg : 0;
}

// G = (N, H, W, KG) floats, KG = ofb_splat_s_kg(C): per target pixel {gout_c / max(den, 1e-3), d loss / d den}
extern "C" int ofb_splat_s_band_bwd_prepass(const float *out, const float *density, const uint8_t *unocc,
                                            const float *gout, const float *gdensity, float *G,
                                            int N, int C, int H, int W, void *stream)
{
    SplatBwdV1 k;
    if (!density || !G || N <= 0 || C < 0 || H < 1 || W < 1 || N > 65535 || (C > 0 && (!out || !gout))) return OFB_EINVAL;
    if (!splat_bwd_v1(C, &k)) return OFB_EINVAL;
    SplatBwdParams p{nullptr, nullptr, 0, 1, 1.0f, nullptr, 0, nullptr, 0, 0, out, density, unocc, gout, gdensity,
                     nullptr, nullptr, 0, nullptr, N, C, H, W, 0, 0, H, W, 0, H};
    k.pre<<<dim3((W + 63) / 64, (H + 4 * kPreRows - 1) / (4 * kPreRows), N), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, G);
    OFB_CHECK_LAUNCH();
    return 0;
}

// band[4] = {src_y0, global_h, acc_y0, acc_h}: the H x W source tensors are rows [src_y0, src_y0 + H) of a field of
// global_h rows, G stores rows [acc_y0, acc_y0 + acc_h) as (N, acc_h, W, KG).  unocc / gout are the rank's own rows.
extern "C" int ofb_splat_s_band_bwd_gather(const float *cx, const float *cy, int64_t coord_bs, float sign,
                                           const float *data, int64_t data_bs,
                                           const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                                           const float *G, const uint8_t *unocc, const float *gout,
                                           float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                                           int N, int C, int H, int W, const int *band, void *stream)
{
    SplatBwdV1 k;
    if (!cx || !cy || !G || !band || N <= 0 || C < 0 || H < 1 || W < 1 || N > 65535 || (C > 0 && !data)) return OFB_EINVAL;
    if ((gcx == nullptr) != (gcy == nullptr)) return OFB_EINVAL;
    if (gdata && unocc && !gout) return OFB_EINVAL;
    if (!splat_bwd_v1(C, &k)) return OFB_EINVAL;
    const int src_y0 = band[0], gh = band[1], acc_y0 = band[2], acc_h = band[3];
    if (src_y0 < 0 || src_y0 + H > gh || acc_y0 < 0 || acc_h < 1 || acc_y0 + acc_h > gh ||
        (int64_t)acc_h * W > 0x3fffffff)
        return OFB_EINVAL;
    if (!gcx && !gdata) return 0;
    SplatBwdParams p{cx, cy, coord_bs, 1, sign, data, data_bs, flow_mask, flow_mask_bs, occlude_zero,
                     nullptr, nullptr, unocc, gout, nullptr, gcx, gcy, gcoord_bs, gdata, N, C, H, W,
                     src_y0, 0, gh, W, acc_y0, acc_h};
    k.gat_win<<<dim3((W + 63) / 64, (H + 4 * kPreRows - 1) / (4 * kPreRows), N), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, G);
    OFB_CHECK_LAUNCH();
    return 0;
}
