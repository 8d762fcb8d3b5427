// W: backward warp ('t' flows) with fused validity, and its adjoint.
//
// Replaces reference utils.py:541-555 (apply_flow ref 't': grid - flow, normalise_coords, grid_sample(bilinear,
// zeros, align_corners=True)) and the mask tracking of Flow.apply (flow_class.py:896-898,922-934).
//
// Forward, three kernels (measured on B200, profiles/):
//   lean    (default for the plain warp, 1..4 channels): one thread per output pixel, 4 rows per thread; a warp whose
//           taps all lie inside the image takes a path without tap predicates; bit-identical to the direct kernel.
//   direct  (the general path: windows for padded / row-band warps, more than 4 channels, validity only): taps
//           through the read-only path with explicit bounds predicates.
//   staged  (opt-in, OFB_TMA=1; OFB_TMA=2: persistent, double-buffered tile loop): a CTA streams its tile's flow, reduces the bounding box of the sample positions and
//           has one thread issue TMA box copies (cp.async.bulk.tensor.3d) of that source region into shared memory,
//           zero-filled outside the image by the TMA unit; taps are then LDS [reg + imm].  Parity-green but slower
//           (0.52 of the HBM peak against 0.63 on the benchmark's flow field with 32 x 32 tiles and 64 x 48 boxes, 5 CTAs
//           per SM; 0.43 with the first geometry, 64 x 32 tiles / 80 x 48 boxes / 3 CTAs): the box over-fetches 3x from
//           L2 and the load -> TMA -> compute phases of a CTA do not overlap.
// All three are bounded by the L1 data pipe, not by HBM: a tap row of 32 lanes is an unaligned 128-byte span that
// drifts over several image rows (7.3 sectors and 2.3 wavefronts per load on the benchmark's flow field), and
// l1tex__data_pipe_lsu_wavefronts sits at 70 % of its peak whatever the instruction count (the lean kernel issues
// 128 instead of 225 instructions per 32 pixels and takes the same time).
//
// Adjoint: warp_t_tile.cu (default), the register-merge kernel below (windows, fall-back), warp_t_adj.cu (ring, opt-in).
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "warp_t.cuh"

namespace ofb {

int g_fwd_impl = 0;    // tuning knob (ofb_set_option key 5): 0 lean forward kernel, 1 direct kernel

struct WarpFwdParams {
    const float *flow; int64_t flow_bs; float sign;
    const float *src; int64_t src_bs;
    const uint8_t *src_mask; int64_t src_mask_bs;
    const uint8_t *flow_mask; int64_t flow_mask_bs;
    float add_flow, thresh;
    float *out; uint8_t *valid;
    int N, C, H, W;
    Window win;
    // device-side early-out decision (ofb_early_decide) and the flags of the result (combine mode 3 chains):
    // out_flags = int32[N + 1] flags, followed by the per-CTA partial maxima the finishing kernel reduces
    const int32_t *sel; int32_t *out_flags;
    float *dtaps;     // (N, 2C, H, W): d out[c] / d ix, d out[c] / d iy for the adjoint (ofb_warp_t_fwd_d), or nullptr
};

// Per-thread base pointers of one batch item (computed once; everything below is 32-bit offsets from them).
struct ItemPtrs {
    const float *src;         // payload planes of this item, channel group already applied
    float *out;               // output planes of this item, channel group already applied
    const uint8_t *smask;     // payload mask of this item or nullptr
    const uint8_t *fmask;     // flow mask of this item or nullptr
    uint8_t *valid;           // validity plane of this item
};

// One output pixel through global memory: channels [0, CN) (CN = 0: cn_rt channels, runtime loop).
// pix = y*W + x of the output pixel; c_first = absolute index of channel 0 (for add_flow).
struct OutStats { unsigned all, masked; };   // largest |component| of the result as a bit pattern (stats.cu)

template <int CN, bool WANT_VALID, bool ALL_IN, bool STATS = false>
__device__ __forceinline__ void warp_pixel_direct(const WarpFwdParams &p, const Window &wn, const ItemPtrs &ip,
                                                  const PixGeom &g, int pix, float fx, float fy, int c_first, int cn_rt,
                                                  OutStats *os = nullptr)
{
    unsigned okbit = 1u;
    const int HW = p.H * p.W;                       // output / flow plane
    const int HWs = wn.sh * wn.sw, W = wn.sw;       // stored source plane and its row pitch
    if (WANT_VALID) {
        float a, b, c, d;
        if (ip.smask) {
            uint8_t ua, ub, uc, ud;
            load_taps<ALL_IN, uint8_t>(ip.smask + g.o, ip.smask + (g.o + W), g, ua, ub, uc, ud);
            a = mask_to_float(ua); b = mask_to_float(ub); c = mask_to_float(uc); d = mask_to_float(ud);
        } else if (ALL_IN) {
            a = b = c = d = 1.0f;
        } else {
            a = g.pnw ? 1.0f : 0.0f; b = g.pne ? 1.0f : 0.0f; c = g.psw ? 1.0f : 0.0f; d = g.pse ? 1.0f : 0.0f;
        }
        bool ok = blend4(g, a, b, c, d) > p.thresh;                                  // flow_class.py:898,922
        if (ip.fmask) ok = ok && (__ldg(ip.fmask + pix) != 0);                       // flow_class.py:934
        ip.valid[pix] = ok ? 1 : 0;
        okbit = ok ? 1u : 0u;
    }
    if (CN > 0) {
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            const int e = c * HWs + g.o;                                             // one int add per plane
            float a, b, cc, d;
            load_taps<ALL_IN, float>(ip.src + e, ip.src + (e + W), g, a, b, cc, d);
            float r = blend4(g, a, b, cc, d);
            if (p.add_flow != 0.0f && c_first + c < 2)                               // flow_class.py:1804,1808
                r = __fadd_rn(__fmul_rn(p.add_flow, (c_first + c) == 0 ? fx : fy), r);
            ip.out[c * HW + pix] = r;
            if (STATS) {
                const unsigned u = __float_as_uint(r) & 0x7fffffffu;
                os->all = max(os->all, u);
                os->masked = max(os->masked, u * okbit);
            }
        }
    } else {
#pragma unroll 1
        for (int c = 0; c < cn_rt; ++c) {
            const int64_t e = (int64_t)c * HWs + g.o;
            float a, b, cc, d;
            load_taps<ALL_IN, float>(ip.src + e, ip.src + (e + W), g, a, b, cc, d);
            float r = blend4(g, a, b, cc, d);
            if (p.add_flow != 0.0f && c_first + c < 2)
                r = __fadd_rn(__fmul_rn(p.add_flow, (c_first + c) == 0 ? fx : fy), r);
            ip.out[(int64_t)c * HW + pix] = r;
        }
    }
}

// dispatch on the warp-uniform "every tap of every lane is inside the image" predicate (true for all but the
// warps that sample near the image border)
template <int CN, bool WANT_VALID, bool STATS = false>
__device__ __forceinline__ void warp_pixel(const WarpFwdParams &p, const Window &wn, const ItemPtrs &ip, const PixGeom &g,
                                           int pix, float fx, float fy, int c_first, int cn_rt, OutStats *os = nullptr)
{
    warp_pixel_direct<CN, WANT_VALID, false, STATS>(p, wn, ip, g, pix, fx, fy, c_first, cn_rt, os);
}

// Device-side early-outs.  The reference decides on the host whether a flow is "all zero" and then returns one of its
// inputs unchanged (utils.py:497-498, flow_class.py:1729-1744); deciding here, from the flags ofb_flow_stats_batch left
// on the device, gives the same results without a host synchronisation.  `rules` holds up to three rules, one per byte,
// evaluated in order: bit 7 set = rule present, bit 5 = which flags (0: warping flow, 1: payload flow), bits 2-4 =
// log2 of the OFB_STAT_* bit that must be CLEAR in the batch summary word flags[N], bits 0-1 = what to do:
//   1  out = add_flow * flow + src,  valid = src_mask & flow_mask      (apply with a zero flow; A zero in combine)
//   2  out = src,                    valid = src_mask                  ("return the payload flow")
//   3  out = add_flow * flow,        valid = flow_mask                 ("return the warping flow")
// One thread evaluates the rules ONCE per call and leaves the decision in *sel; the warp kernels and their adjoints
// read that single word (every CTA testing the flags itself cost 6-18 % of the kernel, measured).
__global__ void early_decide_kernel(const int32_t *fflags, const int32_t *pflags, int rules, int N, int32_t *sel)
{
    int r = 0;
    for (int k = 0; k < 3 && r == 0; ++k) {
        const int rule = (rules >> (8 * k)) & 0xff;
        if (!(rule & 0x80)) break;
        const int32_t *fl = (rule & 0x20) ? pflags : fflags;
        if (fl && !(fl[N] & (1 << ((rule >> 2) & 7)))) r = rule & 3;
    }
    *sel = r;
}

// per-lane variant for the cold early-out kernel (no warp collective: lanes may sit on different items)
__device__ __forceinline__ void publish_stats_lane(int32_t *flags, int n, int N, unsigned all, unsigned masked)
{
    if (all == 0) return;
    const unsigned thr = __float_as_uint(OFB_THRESHOLD);
    int fl = OFB_STAT_NONZERO;
    if (all >= 0x7f800000u) fl |= OFB_STAT_NONFINITE;
    if (all >= thr) fl |= OFB_STAT_NONZERO_T;
    if (masked != 0) fl |= OFB_STAT_NONZERO_M;
    if (masked >= thr) fl |= OFB_STAT_NONZERO_TM;
    if ((*reinterpret_cast<volatile int32_t *>(flags + n) & fl) != fl) atomicOr(flags + n, fl);
    if ((*reinterpret_cast<volatile int32_t *>(flags + N) & fl) != fl) atomicOr(flags + N, fl);
}

// flags of a result: every CTA leaves its two maxima in the workspace behind the flags with plain stores (no atomics,
// no same-address traffic, nothing to wait for); stats_finish_kernel folds them.  Reached by every thread of the CTA.
__device__ __forceinline__ void publish_stats(int32_t *flags, int N, unsigned all, unsigned masked)
{
    __shared__ unsigned s_all[8], s_masked[8];
    all = __reduce_max_sync(0xffffffffu, all);
    masked = __reduce_max_sync(0xffffffffu, masked);
    if ((threadIdx.x & 31) == 0) { s_all[threadIdx.x >> 5] = all; s_masked[threadIdx.x >> 5] = masked; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { all = max(all, s_all[k]); masked = max(masked, s_masked[k]); }
        const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        unsigned *ws = reinterpret_cast<unsigned *>(flags + (N + 1));
        ws[2 * cta] = all;
        ws[2 * cta + 1] = masked;
    }
}

// one CTA per batch item: fold the item's partial maxima into its flags and into the batch summary
__global__ void __launch_bounds__(256) stats_finish_kernel(int32_t *flags, int N, int ctas_per_item)
{
    const int n = blockIdx.x;
    const unsigned *ws = reinterpret_cast<const unsigned *>(flags + (N + 1)) + (int64_t)n * ctas_per_item * 2;
    unsigned all = 0, masked = 0;
    for (int i = threadIdx.x; i < ctas_per_item; i += 256) { all = max(all, ws[2 * i]); masked = max(masked, ws[2 * i + 1]); }
    __shared__ unsigned s_all[8], s_masked[8];
    all = __reduce_max_sync(0xffffffffu, all);
    masked = __reduce_max_sync(0xffffffffu, masked);
    if ((threadIdx.x & 31) == 0) { s_all[threadIdx.x >> 5] = all; s_masked[threadIdx.x >> 5] = masked; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { all = max(all, s_all[k]); masked = max(masked, s_masked[k]); }
        publish_stats_lane(flags, n, N, all, masked);
    }
}

__device__ __forceinline__ ItemPtrs item_ptrs(const WarpFwdParams &p, const Window &wn, int n, int c_off)
{
    const int64_t HW = (int64_t)p.H * p.W, HWs = (int64_t)wn.sh * wn.sw;
    ItemPtrs ip;
    ip.src = p.src + (int64_t)n * p.src_bs + c_off * HWs;
    ip.out = p.out + ((int64_t)n * p.C + c_off) * HW;
    ip.smask = p.src_mask ? p.src_mask + (int64_t)n * p.src_mask_bs : nullptr;
    ip.fmask = p.flow_mask ? p.flow_mask + (int64_t)n * p.flow_mask_bs : nullptr;
    ip.valid = p.valid + (int64_t)n * HW;
    // Opaque to the optimiser from here on: otherwise ptxas re-derives base + n*stride in 64-bit arithmetic at every
    // use (5 instructions per tap pair, measured) instead of keeping five pointers in registers.
    asm volatile("" : "+l"(ip.src), "+l"(ip.out), "+l"(ip.smask), "+l"(ip.fmask), "+l"(ip.valid));
    return ip;
}

#ifndef OFB_FWD_ROWS
#define OFB_FWD_ROWS 4
#endif
#ifndef OFB_FWD_MINB
#define OFB_FWD_MINB 8
#endif
constexpr int kDirRows = OFB_FWD_ROWS;    // rows per thread; CTA tile = 64 x 4*rows (2 warps across x, 4 across y)

// The early-out results are plain copies.  They live in a kernel of their own (launched behind the warp whenever rules
// are given; it returns at once unless a rule fired) so that the warp kernel keeps its 32 registers.
__global__ void __launch_bounds__(256) warp_fwd_early_kernel(const WarpFwdParams p)
{
    const int sel = *p.sel;
    if (!sel) return;
    const int HW = p.H * p.W;
    if (p.out_flags) {      // the warp's own partial maxima describe a result that is being replaced
        const int64_t words = 2 * (int64_t)p.N * (((p.W + 63) / 64) * ((p.H + 4 * kDirRows - 1) / (4 * kDirRows)));
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
            p.out_flags[p.N + 1 + i] = 0;
    }
    unsigned all = 0, masked = 0;
    int last_n = -1;
    const int64_t total = (int64_t)p.N * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / HW), pix = (int)(i - (int64_t)n * HW);
        if (n != last_n) {
            if (p.out_flags && last_n >= 0) publish_stats_lane(p.out_flags, last_n, p.N, all, masked);
            all = masked = 0;
            last_n = n;
        }
        const float *fl = p.flow + (int64_t)n * p.flow_bs;
        unsigned ok = 1u;
        if (p.valid) {
            if (sel != 3 && p.src_mask) ok &= (p.src_mask[(int64_t)n * p.src_mask_bs + pix] != 0);
            if (sel != 2 && p.flow_mask) ok &= (p.flow_mask[(int64_t)n * p.flow_mask_bs + pix] != 0);
            p.valid[(int64_t)n * HW + pix] = (uint8_t)ok;
        }
        for (int c = 0; c < p.C; ++c) {
            float v = (sel == 3) ? 0.0f : p.src[(int64_t)n * p.src_bs + (int64_t)c * HW + pix];
            // add_flow * f is exact for add_flow = 1 (the only value in use); f + 0.0 = f
            if (sel != 2 && p.add_flow != 0.0f && c < 2) v = __fadd_rn(__fmul_rn(p.add_flow, fl[c * HW + pix]), v);
            p.out[((int64_t)n * p.C + c) * HW + pix] = v;
            if (c < 2) {
                const unsigned u = __float_as_uint(v) & 0x7fffffffu;
                all = max(all, u);
                masked = max(masked, u * ok);
            }
        }
    }
    if (p.out_flags && last_n >= 0) publish_stats_lane(p.out_flags, last_n, p.N, all, masked);
}

#ifndef OFB_LEAN_PF
#define OFB_LEAN_PF 128
#endif
// L2 prefetch of the lines the CTA launched `dist` slots later will stream (flow, payload rows, masks of its 64 x 16
// pixel block): one 128-byte line per thread of a 256-thread CTA, fire-and-forget (no register lands, no scoreboard).
// Used by the lean forward only: the same prologue made the direct kernel with result flags (combine_flows) 4 % slower
// and left the tile adjoint unchanged (both measured at distance 16..256).
template <int CN>
__device__ __forceinline__ void prefetch_ahead(const WarpFwdParams &p, const int dist, const int cta_rows)
{
    const int W = p.W, HW = p.H * p.W;
    const long long lin = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + dist;
    const long long total = (long long)gridDim.x * gridDim.y * gridDim.z;
    if (lin >= total) return;
    const int pz = (int)(lin / ((long long)gridDim.x * gridDim.y));
    const int rem = (int)(lin - (long long)pz * gridDim.x * gridDim.y);
    const int py = rem / (int)gridDim.x, px = rem - py * (int)gridDim.x;
    const int t = threadIdx.x;
    const int row = min(py * cta_rows + (t & 15), p.H - 1), half = (t >> 4) & 1, what = t >> 5;   // 8 groups of 32
    const int col = min(px * 64 + half * 32, W - 1);
    const void *a = nullptr;
    if (what < 2) a = p.flow + (int64_t)pz * p.flow_bs + (int64_t)what * HW + row * W + col;
    else if (what < 2 + CN) a = p.src + (int64_t)pz * p.src_bs + (int64_t)(what - 2) * HW + row * W + col;
    else if (what == 2 + CN && p.src_mask && half == 0) a = p.src_mask + (int64_t)pz * p.src_mask_bs + row * W + col;
    else if (what == 3 + CN && p.flow_mask && half == 0) a = p.flow_mask + (int64_t)pz * p.flow_mask_bs + row * W + col;
    if (a) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
}

template <int CN, bool WANT_VALID, bool WIN, bool STATS = false>
__global__ void __launch_bounds__(256, OFB_FWD_MINB) warp_t_fwd_direct_kernel(const WarpFwdParams p, const CoordCtx cc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kDirRows) + (warp >> 1) * kDirRows;
    const int n = blockIdx.z;
    // (an early-out decision is NOT tested here: the warp runs regardless and warp_fwd_early_kernel, launched behind
    // it, overwrites the result in the rare case a rule fired -- testing per CTA cost 5-18 % of this kernel, measured)
    const bool inside = x < p.W;
    if (!STATS && !inside) return;
    const int HW = p.H * p.W;
    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    float fx[kDirRows], fy[kDirRows];
#pragma unroll
    for (int r = 0; r < kDirRows; ++r) {
        const int q = min(y0 + r, p.H - 1) * p.W + min(x, p.W - 1);
        fx[r] = __ldg(fl + q);
        fy[r] = __ldg(fl + (HW + q));
    }
    // identity window: built from H, W so that everything folds to the plain-warp code at compile time
    const Window wn = WIN ? p.win : Window{0, 0, p.H, p.W, 0, 0, p.H, p.W};
    const float xf = (float)(x + wn.x0);
    const ItemPtrs ip = item_ptrs(p, wn, n, 0);
    const int rows = (STATS && !inside) ? 0 : min(kDirRows, p.H - y0);
    OutStats os{0u, 0u};
#pragma unroll
    for (int r = 0; r < kDirRows; ++r) {
        if (r >= rows) break;
        float ix, iy;
        sample_pos(cc, fx[r], fy[r], xf, (float)(y0 + r + wn.y0), ix, iy);
        const PixGeom g = pix_geom<WIN>(ix, iy, wn);
        warp_pixel<CN, WANT_VALID, STATS>(p, wn, ip, g, (y0 + r) * p.W + x, fx[r], fy[r], 0, p.C, &os);
    }
    if (STATS) publish_stats(p.out_flags, p.N, os.all, os.masked);
}

// ------------------------------------------------------------------------------------------------
// lean forward (the default for the plain warp: identity window, 1..4 channels)
// ------------------------------------------------------------------------------------------------
// Same mapping and arithmetic as the direct kernel, organised around what the SASS of that kernel showed it paid for
// (225 warp-instructions per 32 pixels at 32 registers, 73 % issue-active):
//   * a warp whose 32 x 4 taps all lie inside the image -- all but the warps sampling at the border -- takes a path
//     with no tap predicates and no zero initialisation: one 64-bit address per tap ROW and plane, derived from
//     per-item bases that stay in UNIFORM registers (they depend on blockIdx.z only);
//   * everything else falls back to warp_pixel_direct, per warp and per row.
#ifndef OFB_LEAN_ROWS
#define OFB_LEAN_ROWS 4
#endif
#ifndef OFB_LEAN_MINB
#define OFB_LEAN_MINB 8
#endif
constexpr int kLeanRows = OFB_LEAN_ROWS;

//   * WANT_D: the kernel also writes d out[c] / d ix and d out[c] / d iy (2C planes).  The adjoint's flow gradient is
//     then a stream (sum_c gout[c] * d[c]) instead of a second gather of the 4C payload taps, which is what bounds
//     both kernels (L1 wavefronts of unaligned tap rows): 24 B/pixel more traffic for C = 3, about 30 % less adjoint.
//   * WIN: windowed frame (struct Window: padded apply with cut, row bands); the taps live in stored-region coordinates.
template <int CN, bool WANT_VALID, bool STATS, bool WANT_D, bool WIN = false>
__global__ void __launch_bounds__(256, OFB_LEAN_MINB) warp_t_fwd_lean_kernel(const WarpFwdParams p, const CoordCtx cc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kLeanRows) + (warp >> 1) * kLeanRows;
    const int n = blockIdx.z;
    const int W = p.W, HW = p.H * p.W;
    const bool inside = x < W;
    const unsigned live = __ballot_sync(0xffffffffu, inside);
    if (!STATS && !inside) return;
    const int rows = (STATS && !inside) ? 0 : min(kLeanRows, p.H - y0);
    // per-item bases: functions of blockIdx.z only
    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    const Window wn = WIN ? p.win : Window{0, 0, p.H, W, 0, 0, p.H, W};
    const int SW = wn.sw, SH = wn.sh, HWs = SH * SW;      // stored source region (the field itself without a window)
    const float *src = p.src + (int64_t)n * p.src_bs;
    float *out = p.out + (int64_t)n * CN * HW;
    const uint8_t *smask = p.src_mask ? p.src_mask + (int64_t)n * p.src_mask_bs : nullptr;
    const uint8_t *fmask = p.flow_mask ? p.flow_mask + (int64_t)n * p.flow_mask_bs : nullptr;
    uint8_t *valid = WANT_VALID ? p.valid + (int64_t)n * HW : nullptr;
    float *dt = WANT_D ? p.dtaps + (int64_t)n * (2 * CN) * HW : nullptr;
    asm volatile("" : "+l"(src), "+l"(out), "+l"(smask), "+l"(fmask), "+l"(valid), "+l"(dt));    // see item_ptrs()

    // L2 prefetch of the lines the CTA launched OFB_LEAN_PF slots later will stream (its flow, payload rows and masks): one
    // 128-byte line per thread, fire-and-forget (no register lands, no scoreboard).  The kernel waits on memory with its
    // register file full of loads in flight (DESIGN 3 "What bounds W"); turning that CTA's first-touch misses into L2
    // hits shortens every wait: -11 % at C = 3, -7 % at C = 2, -13 % at C = 4 (batch 32 x 1080p; distances 16..450 within
    // 1 %, 900: -6 %, 2400: +28 %), +12 % at C = 1, which is therefore left out.
    if (OFB_LEAN_PF > 0 && !WIN && CN >= 2) prefetch_ahead<CN>(p, OFB_LEAN_PF, 4 * kLeanRows);
    float fx[kLeanRows], fy[kLeanRows];
    {
        const int q0 = y0 * W + min(x, W - 1);
#pragma unroll
        for (int r = 0; r < kLeanRows; ++r) {
            const int q = (y0 + r < p.H) ? q0 + r * W : q0;
            fx[r] = __ldg(fl + q);
            fy[r] = __ldg(fl + (HW + q));
        }
    }
    const float xf = (float)(x + wn.x0);
    OutStats os{0u, 0u};
#pragma unroll
    for (int r = 0; r < kLeanRows; ++r) {
        if (r >= rows) break;
        const int pix = (y0 + r) * W + x;
        float ix, iy;
        ix = round_trip_unchecked(__fmaf_rn(-cc.sign, fx[r], xf), cc.dw);
        iy = round_trip_unchecked(__fmaf_rn(-cc.sign, fy[r], (float)(y0 + r + wn.y0)), cc.dh);
        const int xi = (int)((unsigned)__float2int_rd(ix) - (unsigned)wn.sx0);       // stored-region coordinates
        const int yi = (int)((unsigned)__float2int_rd(iy) - (unsigned)wn.sy0);
        const bool all_in = (unsigned)xi < (unsigned)(SW - 1) && (unsigned)yi < (unsigned)(SH - 1);
        if (__all_sync(live, all_in)) {   // lanes right of the field never get here
            const float x0 = floorf(ix), y0f = floorf(iy);
            const float ex = __fsub_rn(__fadd_rn(x0, 1.0f), ix), wx = __fsub_rn(ix, x0);
            const float sy = __fsub_rn(__fadd_rn(y0f, 1.0f), iy), ny = __fsub_rn(iy, y0f);
            const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
            const int o = yi * SW + xi;
            unsigned okbit = 1u;
            if (WANT_VALID) {
                float m;
                if (smask) {
                    const uint8_t *mn = smask + o, *ms = smask + (o + SW);
                    const float a = mask_to_float(__ldg(mn)), b = mask_to_float(__ldg(mn + 1));
                    const float c = mask_to_float(__ldg(ms)), d = mask_to_float(__ldg(ms + 1));
                    m = __fmaf_rn(d, wse, __fmaf_rn(c, wsw, __fmaf_rn(b, wne, __fmul_rn(a, wnw))));
                } else {
                    m = __fadd_rn(__fadd_rn(__fadd_rn(wnw, wne), wsw), wse);     // 1*w == w, fma(1, w, r) == w + r
                }
                bool ok = m > p.thresh;                                          // flow_class.py:898,922
                if (fmask) ok = ok && (__ldg(fmask + pix) != 0);                 // flow_class.py:934
                valid[pix] = ok ? 1 : 0;
                okbit = ok ? 1u : 0u;
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const float *tn = src + (c * HWs + o), *ts = src + (c * HWs + o + SW);
                const float a = __ldg(tn), b = __ldg(tn + 1), cq = __ldg(ts), d = __ldg(ts + 1);
                float v = __fmaf_rn(d, wse, __fmaf_rn(cq, wsw, __fmaf_rn(b, wne, __fmul_rn(a, wnw))));
                if (p.add_flow != 0.0f && c < 2)                                 // flow_class.py:1804,1808
                    v = __fadd_rn(__fmul_rn(p.add_flow, c == 0 ? fx[r] : fy[r]), v);
                out[c * HW + pix] = v;
                if (WANT_D) {           // same expressions as the adjoint kernels
                    dt[(2 * c) * HW + pix] = __fmaf_rn(b - a, sy, (d - cq) * ny);
                    dt[(2 * c + 1) * HW + pix] = __fmaf_rn(cq - a, ex, (d - b) * wx);
                }
                if (STATS) {
                    const unsigned u = __float_as_uint(v) & 0x7fffffffu;
                    os.all = max(os.all, u);
                    os.masked = max(os.masked, u * okbit);
                }
            }
        } else {
            ItemPtrs ip{src, out, smask, fmask, valid};
            const PixGeom g = pix_geom<WIN>(ix, iy, wn);
            warp_pixel_direct<CN, WANT_VALID, false, STATS>(p, wn, ip, g, pix, fx[r], fy[r], 0, CN, &os);
            if (WANT_D) {
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    float a, b, cq, d;
                    load_taps<false, float>(src + (c * HWs + g.o), src + (c * HWs + g.o + SW), g, a, b, cq, d);
                    dt[(2 * c) * HW + pix] = __fmaf_rn(b - a, g.sy, (d - cq) * g.ny);
                    dt[(2 * c + 1) * HW + pix] = __fmaf_rn(cq - a, g.ex, (d - b) * g.wx);
                }
            }
        }
    }
    if (STATS) publish_stats(p.out_flags, p.N, os.all, os.masked);
}

// ------------------------------------------------------------------------------------------------
// staged forward
// ------------------------------------------------------------------------------------------------
#ifndef OFB_ST_TW
#define OFB_ST_TW 32
#endif
#ifndef OFB_ST_ROWS
#define OFB_ST_ROWS 4
#endif
#ifndef OFB_ST_BOXW
#define OFB_ST_BOXW 64
#endif
#ifndef OFB_ST_BOXH
#define OFB_ST_BOXH 48
#endif
#ifndef OFB_ST_MINB
#define OFB_ST_MINB 5
#endif
constexpr int kTileW = OFB_ST_TW, kStWX = kTileW / 32, kRows = OFB_ST_ROWS;   // 8 warps: kStWX across x, 8 / kStWX across y
constexpr int kTileH = (8 / kStWX) * kRows;
constexpr int kBoxW = OFB_ST_BOXW, kBoxH = OFB_ST_BOXH;   // staged source region (floats): tile + slack
constexpr int kPlane = kBoxW * kBoxH;                     // elements per staged channel
constexpr int kMBoxW = kBoxW + 16;                        // byte-mask box is wider: its origin is 16-pixel aligned
constexpr int kMaxCN = 4;                                 // payload channels staged per launch
// TMA needs the box origin 16-byte aligned along x: multiples of 4 floats / 16 mask bytes (measured: anything else
// raises an illegal-instruction fault).  Rounding the origin down costs at most 3 (15) columns of the box.

struct StageCtl {
    unsigned long long bar;   // mbarrier
    int minx, maxx, miny, maxy;
    int mode;                 // 0 staged, 1 direct (footprint too large), 2 nothing in bounds
};

constexpr int staged_smem_bytes(int cn, int vmode)
{
    return kPlane * 4 * cn + (vmode == 2 ? kMBoxW * kBoxH : 0) + 64;
}

// CN    payload channels handled by this launch (1..4), planes [c_off, c_off + CN) of C_total
// VMODE 0: no validity output; 1: validity with an all-true payload mask; 2: validity with a staged payload mask
template <int CN, int VMODE>
__global__ void __launch_bounds__(256, OFB_ST_MINB) warp_t_fwd_staged_kernel(const WarpFwdParams p,
                                                                   const __grid_constant__ CUtensorMap src_map,
                                                                   const __grid_constant__ CUtensorMap mask_map,
                                                                   const int c_off, const CoordCtx cc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    uint8_t *mtile = smem + kPlane * 4 * CN;
    StageCtl &ctl = *reinterpret_cast<StageCtl *>(smem + kPlane * 4 * CN + (VMODE == 2 ? kMBoxW * kBoxH : 0));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * kTileW + (warp % kStWX) * 32 + lane;
    const int y0 = blockIdx.y * kTileH + (warp / kStWX) * kRows;
    const int n = blockIdx.z;
    const int HW = p.H * p.W;
    const int xc = min(x, p.W - 1);                       // clamped copies for loads; stores are predicated
    const float xf = (float)x;

    if (threadIdx.x == 0) {
        ctl.minx = INT_MAX; ctl.miny = INT_MAX; ctl.maxx = INT_MIN; ctl.maxy = INT_MIN;
        mbar_init(reinterpret_cast<uint64_t *>(&ctl.bar), 1);
    }
    __syncthreads();

    // ---- phase 1: flow -> sample positions, footprint of the tile
    float ix[kRows], iy[kRows];
    {
        const float *fl = p.flow + (int64_t)n * p.flow_bs + xc;
        float fx[kRows], fy[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int yc = min(y0 + r, p.H - 1);
            fx[r] = __ldg(fl + yc * p.W);
            fy[r] = __ldg(fl + HW + yc * p.W);
        }
        const float wm1 = cc.dw.d, hm1 = cc.dh.d;
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            sample_pos(cc, fx[r], fy[r], xf, (float)(y0 + r), ix[r], iy[r]);
            const float fx0 = floorf(ix[r]), fy0 = floorf(iy[r]);
            // pixels with at least one tap inside the image: x0 in [-1, W-1], y0 in [-1, H-1]
            const bool any = (x < p.W) && (y0 + r < p.H) && fx0 >= -1.0f && fx0 <= wm1 && fy0 >= -1.0f && fy0 <= hm1;
            const int xi = any ? (int)fx0 : INT_MAX, yi = any ? (int)fy0 : INT_MAX;
            minx = min(minx, xi); miny = min(miny, yi);
            maxx = max(maxx, any ? xi : INT_MIN); maxy = max(maxy, any ? yi : INT_MIN);
        }
        minx = __reduce_min_sync(0xffffffffu, minx); maxx = __reduce_max_sync(0xffffffffu, maxx);
        miny = __reduce_min_sync(0xffffffffu, miny); maxy = __reduce_max_sync(0xffffffffu, maxy);
        if (lane == 0 && minx != INT_MAX) {
            atomicMin(&ctl.minx, minx); atomicMax(&ctl.maxx, maxx);
            atomicMin(&ctl.miny, miny); atomicMax(&ctl.maxy, maxy);
        }
    }
    __syncthreads();

    // ---- phase 2: one thread decides and issues the box copies
    if (threadIdx.x == 0) {
        int mode = 0;
        if (ctl.minx == INT_MAX) mode = 2;
        else if (ctl.maxx + 2 - (ctl.minx & ~3) > kBoxW || ctl.maxy - ctl.miny + 2 > kBoxH) mode = 1;
        ctl.mode = mode;
        if (mode == 0) {
            mbar_expect_tx(reinterpret_cast<uint64_t *>(&ctl.bar),
                           (uint32_t)(kPlane * 4 * CN + (VMODE == 2 ? kMBoxW * kBoxH : 0)));
            tma_load_3d(tile, &src_map, ctl.minx & ~3, ctl.miny, ((p.src_bs == 0) ? 0 : n * p.C) + c_off,
                        reinterpret_cast<uint64_t *>(&ctl.bar));
            if (VMODE == 2)
                tma_load_3d(mtile, &mask_map, ctl.minx & ~15, ctl.miny, (p.src_mask_bs == 0) ? 0 : n,
                            reinterpret_cast<uint64_t *>(&ctl.bar));
        }
    }
    __syncthreads();
    const int mode = ctl.mode;
    const int ox = ctl.minx & ~3, oy = ctl.miny;      // origin of the payload box
    const int mdx = ox - (ctl.minx & ~15);            // payload-box column -> mask-box column

    if (mode == 1) {   // footprint larger than the box: this tile goes through global memory
        if (x >= p.W) return;
        const float *fl = p.flow + (int64_t)n * p.flow_bs + x;
        const Window wn{0, 0, p.H, p.W, 0, 0, p.H, p.W};
        ItemPtrs ip = item_ptrs(p, wn, n, c_off);
        if (VMODE != 2) ip.smask = nullptr;
#pragma unroll 1
        for (int r = 0; r < kRows; ++r) {            // positions are recomputed: keeps this cold path small
            const int y = y0 + r;
            if (y >= p.H) break;
            const float fx = __ldg(fl + y * p.W), fy = __ldg(fl + HW + y * p.W);
            float jx, jy;
            sample_pos(cc, fx, fy, xf, (float)y, jx, jy);
            const PixGeom g = pix_geom<false>(jx, jy, wn);
            warp_pixel_direct<CN, VMODE != 0, false>(p, wn, ip, g, y * p.W + x, fx, fy, c_off, CN);
        }
        return;
    }
    if (mode == 0) mbar_wait(reinterpret_cast<uint64_t *>(&ctl.bar), 0);

    // ---- phase 3: taps from shared memory, no bounds logic (the TMA unit zero-filled what lies outside)
    float *out = p.out + ((int64_t)n * p.C + c_off) * HW + x;
    const float wm1 = cc.dw.d, hm1 = cc.dh.d;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const float fx0 = floorf(ix[r]), fy0 = floorf(iy[r]);
        const bool any = mode == 0 && fx0 >= -1.0f && fx0 <= wm1 && fy0 >= -1.0f && fy0 <= hm1;
        const bool st = (x < p.W) && (y0 + r < p.H);
        const int xi = any ? (int)fx0 - ox : 0, yi = any ? (int)fy0 - oy : 0;
        const int off = yi * kBoxW + xi;
        const float ex = __fsub_rn(__fadd_rn(fx0, 1.0f), ix[r]), wx = __fsub_rn(ix[r], fx0);
        const float sy = __fsub_rn(__fadd_rn(fy0, 1.0f), iy[r]), ny = __fsub_rn(iy[r], fy0);
        const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
        const int pix = (y0 + r) * p.W;

        if (VMODE != 0) {   // the payload mask rides through the same weights (flow_class.py:898,922,934)
            float a, b, c, d;
            if (VMODE == 2) {
                const uint8_t *m = mtile + yi * kMBoxW + xi + mdx;
                a = m[0] ? 1.0f : 0.0f; b = m[1] ? 1.0f : 0.0f;
                c = m[kMBoxW] ? 1.0f : 0.0f; d = m[kMBoxW + 1] ? 1.0f : 0.0f;
            } else {
                const int gx = xi + ox, gy = yi + oy;
                const bool bx0 = gx >= 0, bx1 = gx + 1 <= p.W - 1, by0 = gy >= 0, by1 = gy + 1 <= p.H - 1;
                a = (bx0 && by0) ? 1.0f : 0.0f; b = (bx1 && by0) ? 1.0f : 0.0f;
                c = (bx0 && by1) ? 1.0f : 0.0f; d = (bx1 && by1) ? 1.0f : 0.0f;
            }
            float m = __fmul_rn(a, wnw);
            m = __fmaf_rn(b, wne, m);
            m = __fmaf_rn(c, wsw, m);
            m = __fmaf_rn(d, wse, m);
            bool ok = any && (m > p.thresh);
            if (st) {
                if (p.flow_mask) ok = ok && (__ldg(p.flow_mask + (int64_t)n * p.flow_mask_bs + pix + x) != 0);
                p.valid[(int64_t)n * HW + pix + x] = ok ? 1 : 0;
            }
        }
        const float *t = tile + off;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            float v = __fmul_rn(t[c * kPlane], wnw);
            v = __fmaf_rn(t[c * kPlane + 1], wne, v);
            v = __fmaf_rn(t[c * kPlane + kBoxW], wsw, v);
            v = __fmaf_rn(t[c * kPlane + kBoxW + 1], wse, v);
            v = any ? v : 0.0f;
            if (p.add_flow != 0.0f && c_off + c < 2 && st) {                          // flow_class.py:1804,1808
                const float f = __ldg(p.flow + (int64_t)n * p.flow_bs + (c_off + c) * HW + pix + x);
                v = __fadd_rn(__fmul_rn(p.add_flow, f), v);
            }
            if (st) out[(int64_t)c * HW + pix] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// staged forward, persistent and double-buffered (opt-in, OFB_TMA=2)
// ------------------------------------------------------------------------------------------------
// The same staging as above with the phases of consecutive tiles overlapped: a CTA walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...; while it evaluates tile k from stage k & 1, the TMA boxes of tile k + 1 are in flight into
// the other stage and the flow of tile k + 2 is in flight into registers.  Needs 1..4 channels in one group.
struct PipeCtl {
    unsigned long long bar[2];      // one mbarrier per stage
    int ox[2], oy[2], mode[2];      // per stage: box origin and 0 staged / 1 direct (footprint too large) / 2 nothing in bounds
    int minx, maxx, miny, maxy;     // footprint reduction scratch
};

template <int CN, int VMODE>
__global__ void __launch_bounds__(256, 2) warp_t_fwd_pipe_kernel(const WarpFwdParams p,
                                                                 const __grid_constant__ CUtensorMap src_map,
                                                                 const __grid_constant__ CUtensorMap mask_map,
                                                                 const CoordCtx cc, const int tiles_x, const int tiles_y)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int kStage = kPlane * 4 * CN + (VMODE == 2 ? kMBoxW * kBoxH : 0);
    constexpr int kStageAl = (kStage + 1023) & ~1023;
    PipeCtl &ctl = *reinterpret_cast<PipeCtl *>(smem + 2 * kStageAl);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int HW = p.H * p.W;
    const int ntiles = tiles_x * tiles_y * p.N;
    const float wm1 = cc.dw.d, hm1 = cc.dh.d;

    if (threadIdx.x == 0) {
        mbar_init(reinterpret_cast<uint64_t *>(&ctl.bar[0]), 1);
        mbar_init(reinterpret_cast<uint64_t *>(&ctl.bar[1]), 1);
        ctl.minx = INT_MAX; ctl.miny = INT_MAX; ctl.maxx = INT_MIN; ctl.maxy = INT_MIN;
    }
    __syncthreads();

    auto tile_xyn = [&](int t, int &x, int &y0, int &n) {
        n = t / (tiles_x * tiles_y);
        const int rem = t - n * (tiles_x * tiles_y);
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        x = tx * kTileW + (warp % kStWX) * 32 + lane;
        y0 = ty * kTileH + (warp / kStWX) * kRows;
    };
    auto load_flow = [&](int t, float (&fx)[kRows], float (&fy)[kRows]) {
        int x, y0, n;
        tile_xyn(t, x, y0, n);
        const float *fl = p.flow + (int64_t)n * p.flow_bs + min(x, p.W - 1);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int yc = min(y0 + r, p.H - 1);
            fx[r] = __ldg(fl + yc * p.W);
            fy[r] = __ldg(fl + HW + yc * p.W);
        }
    };
    // sample positions of tile t, footprint reduction, box copies into stage s (issued by thread 0)
    auto stage_tile = [&](int t, int s, const float (&fx)[kRows], const float (&fy)[kRows], float (&ix)[kRows],
                          float (&iy)[kRows]) {
        int x, y0, n;
        tile_xyn(t, x, y0, n);
        const float xf = (float)x;
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            sample_pos(cc, fx[r], fy[r], xf, (float)(y0 + r), ix[r], iy[r]);
            const float fx0 = floorf(ix[r]), fy0 = floorf(iy[r]);
            const bool any = (x < p.W) && (y0 + r < p.H) && fx0 >= -1.0f && fx0 <= wm1 && fy0 >= -1.0f && fy0 <= hm1;
            if (any) {
                const int xi = (int)fx0, yi = (int)fy0;
                minx = min(minx, xi); maxx = max(maxx, xi); miny = min(miny, yi); maxy = max(maxy, yi);
            }
        }
        minx = __reduce_min_sync(0xffffffffu, minx); maxx = __reduce_max_sync(0xffffffffu, maxx);
        miny = __reduce_min_sync(0xffffffffu, miny); maxy = __reduce_max_sync(0xffffffffu, maxy);
        if (lane == 0 && minx != INT_MAX) {
            atomicMin(&ctl.minx, minx); atomicMax(&ctl.maxx, maxx);
            atomicMin(&ctl.miny, miny); atomicMax(&ctl.maxy, maxy);
        }
        __syncthreads();                       // footprint complete; every thread is done with stage s (tile t - 2)
        if (threadIdx.x == 0) {
            int mode = 0;
            if (ctl.minx == INT_MAX) mode = 2;
            else if (ctl.maxx + 2 - (ctl.minx & ~3) > kBoxW || ctl.maxy - ctl.miny + 2 > kBoxH) mode = 1;
            ctl.mode[s] = mode; ctl.ox[s] = ctl.minx & ~3; ctl.oy[s] = ctl.miny;
            if (mode == 0) {
                unsigned char *st = smem + s * kStageAl;
                fence_proxy_async();           // the stage was last read through the generic proxy
                mbar_expect_tx(reinterpret_cast<uint64_t *>(&ctl.bar[s]), (uint32_t)kStage);
                tma_load_3d(st, &src_map, ctl.minx & ~3, ctl.miny, (p.src_bs == 0) ? 0 : n * p.C,
                            reinterpret_cast<uint64_t *>(&ctl.bar[s]));
                if (VMODE == 2)
                    tma_load_3d(st + kPlane * 4 * CN, &mask_map, ctl.minx & ~15, ctl.miny, (p.src_mask_bs == 0) ? 0 : n,
                                reinterpret_cast<uint64_t *>(&ctl.bar[s]));
            }
            ctl.minx = INT_MAX; ctl.miny = INT_MAX; ctl.maxx = INT_MIN; ctl.maxy = INT_MIN;
        }
        __syncthreads();                       // stage info published, scratch reset
    };

    const int t0 = blockIdx.x;
    if (t0 >= ntiles) return;
    float fxn[kRows], fyn[kRows], ix[kRows], iy[kRows], ixn[kRows], iyn[kRows];
    load_flow(t0, fxn, fyn);
    stage_tile(t0, 0, fxn, fyn, ix, iy);
    if (t0 + (int)gridDim.x < ntiles) load_flow(t0 + gridDim.x, fxn, fyn);

    unsigned phase = 0u;                        // bit s: parity of stage s's next barrier completion
    for (int k = 0;; ++k) {
        const int s = k & 1;
        const int t = t0 + k * (int)gridDim.x, tn = t + (int)gridDim.x, tnn = tn + (int)gridDim.x;
        if (tn < ntiles) stage_tile(tn, s ^ 1, fxn, fyn, ixn, iyn);
        if (tnn < ntiles) load_flow(tnn, fxn, fyn);

        // ---- evaluate tile t from stage s
        int x, y0, n;
        tile_xyn(t, x, y0, n);
        const int mode = ctl.mode[s], ox = ctl.ox[s], oy = ctl.oy[s];
        if (mode == 0) {                        // (CTA-uniform) the stage's barrier completes once per staged tile
            mbar_wait(reinterpret_cast<uint64_t *>(&ctl.bar[s]), (phase >> s) & 1u);
            phase ^= 1u << s;
        }
        if (mode == 1) {                        // footprint larger than the box: through global memory
            if (x < p.W) {
                const Window wn{0, 0, p.H, p.W, 0, 0, p.H, p.W};
                ItemPtrs ip = item_ptrs(p, wn, n, 0);
                if (VMODE != 2) ip.smask = nullptr;
                const float *fl = p.flow + (int64_t)n * p.flow_bs + x;
#pragma unroll 1
                for (int r = 0; r < kRows; ++r) {
                    const int y = y0 + r;
                    if (y >= p.H) break;
                    const PixGeom g = pix_geom<false>(ix[r], iy[r], wn);
                    warp_pixel_direct<CN, VMODE != 0, false>(p, wn, ip, g, y * p.W + x, __ldg(fl + y * p.W),
                                                            __ldg(fl + HW + y * p.W), 0, CN);
                }
            }
        } else {
            const float *tile = reinterpret_cast<const float *>(smem + s * kStageAl);
            const uint8_t *mtile = smem + s * kStageAl + kPlane * 4 * CN;
            const int mdx = ox - ((ox + 0) & ~15) ;          // payload-box column -> mask-box column (mask origin = minx & ~15)
            float *out = p.out + (int64_t)n * p.C * HW + x;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float fx0 = floorf(ix[r]), fy0 = floorf(iy[r]);
                const bool any = mode == 0 && fx0 >= -1.0f && fx0 <= wm1 && fy0 >= -1.0f && fy0 <= hm1;
                const bool st = (x < p.W) && (y0 + r < p.H);
                const int xi = any ? (int)fx0 - ox : 0, yi = any ? (int)fy0 - oy : 0;
                const int off = yi * kBoxW + xi;
                const float ex = __fsub_rn(__fadd_rn(fx0, 1.0f), ix[r]), wx = __fsub_rn(ix[r], fx0);
                const float sy = __fsub_rn(__fadd_rn(fy0, 1.0f), iy[r]), ny = __fsub_rn(iy[r], fy0);
                const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
                const int pix = (y0 + r) * p.W;
                if (VMODE != 0) {
                    float a, b, c, d;
                    if (VMODE == 2) {
                        const uint8_t *m = mtile + yi * kMBoxW + xi + mdx;
                        a = m[0] ? 1.0f : 0.0f; b = m[1] ? 1.0f : 0.0f;
                        c = m[kMBoxW] ? 1.0f : 0.0f; d = m[kMBoxW + 1] ? 1.0f : 0.0f;
                    } else {
                        const int gx = xi + ox, gy = yi + oy;
                        const bool bx0 = gx >= 0, bx1 = gx + 1 <= p.W - 1, by0 = gy >= 0, by1 = gy + 1 <= p.H - 1;
                        a = (bx0 && by0) ? 1.0f : 0.0f; b = (bx1 && by0) ? 1.0f : 0.0f;
                        c = (bx0 && by1) ? 1.0f : 0.0f; d = (bx1 && by1) ? 1.0f : 0.0f;
                    }
                    float m = __fmul_rn(a, wnw);
                    m = __fmaf_rn(b, wne, m);
                    m = __fmaf_rn(c, wsw, m);
                    m = __fmaf_rn(d, wse, m);
                    bool ok = any && (m > p.thresh);
                    if (st) {
                        if (p.flow_mask) ok = ok && (__ldg(p.flow_mask + (int64_t)n * p.flow_mask_bs + pix + x) != 0);
                        p.valid[(int64_t)n * HW + pix + x] = ok ? 1 : 0;
                    }
                }
                const float *tp = tile + off;
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    float v = __fmul_rn(tp[c * kPlane], wnw);
                    v = __fmaf_rn(tp[c * kPlane + 1], wne, v);
                    v = __fmaf_rn(tp[c * kPlane + kBoxW], wsw, v);
                    v = __fmaf_rn(tp[c * kPlane + kBoxW + 1], wse, v);
                    v = any ? v : 0.0f;
                    if (p.add_flow != 0.0f && c < 2 && st) {
                        const float f = __ldg(p.flow + (int64_t)n * p.flow_bs + c * HW + pix + x);
                        v = __fadd_rn(__fmul_rn(p.add_flow, f), v);
                    }
                    if (st) out[(int64_t)c * HW + pix] = v;
                }
            }
        }
        if (tn >= ntiles) break;
#pragma unroll
        for (int r = 0; r < kRows; ++r) { ix[r] = ixn[r]; iy[r] = iyn[r]; }
    }
}

typedef void (*PipeKernel)(const WarpFwdParams, const CUtensorMap, const CUtensorMap, const CoordCtx, const int, const int);

template <int CN>
static PipeKernel pipe_kernel_for(int vmode)
{
    if (vmode == 0) return warp_t_fwd_pipe_kernel<CN, 0>;
    if (vmode == 1) return warp_t_fwd_pipe_kernel<CN, 1>;
    return warp_t_fwd_pipe_kernel<CN, 2>;
}

static PipeKernel pipe_kernel(int cn, int vmode)
{
    switch (cn) {
        case 1: return pipe_kernel_for<1>(vmode);
        case 2: return pipe_kernel_for<2>(vmode);
        case 3: return pipe_kernel_for<3>(vmode);
        default: return pipe_kernel_for<4>(vmode);
    }
}

typedef void (*StagedKernel)(const WarpFwdParams, const CUtensorMap, const CUtensorMap, const int, const CoordCtx);

template <int CN>
static StagedKernel staged_kernel_for(int vmode)
{
    if (vmode == 0) return warp_t_fwd_staged_kernel<CN, 0>;
    if (vmode == 1) return warp_t_fwd_staged_kernel<CN, 1>;
    return warp_t_fwd_staged_kernel<CN, 2>;
}

static StagedKernel staged_kernel(int cn, int vmode)
{
    switch (cn) {
        case 1: return staged_kernel_for<1>(vmode);
        case 2: return staged_kernel_for<2>(vmode);
        case 3: return staged_kernel_for<3>(vmode);
        default: return staged_kernel_for<4>(vmode);
    }
}

// ------------------------------------------------------------------------------------------------
// adjoint
// ------------------------------------------------------------------------------------------------
// gflow is a gather (payload taps x gout).  gsrc is the un-normalised splat of w*gout; issued naively that is 4*C
// red.global.add per pixel and the L2 atomic unit becomes the limit (measured: ~0.5-0.7 T lane-ops/s).  Neighbouring
// output pixels of a smooth flow hit neighbouring source pixels, so contributions are merged in registers first:
//   horizontally  lane l's east taps land on lane l+1's west taps when o(l+1) == o(l)+1  -> one shuffle per value;
//   vertically    a thread walks kBwdRows rows of one column and carries its south taps, which are the next row's
//                 north taps when o(next) == o + W.
// What is left is ~1 coalesced RED per pixel and channel plus sparse leftovers where the chain breaks.  Shared-memory
// accumulation was rejected: float atomicAdd on shared memory is a CAS loop on sm_100 (ATOMS.CAST.SPIN, ~8 clk per
// warp-op, tools/microbench.cu).
#ifndef OFB_BWD_ROWS
#define OFB_BWD_ROWS 4
#endif
#ifndef OFB_BWD_MINB
#define OFB_BWD_MINB 4
#endif
constexpr int kBwdRows = OFB_BWD_ROWS;     // rows per thread; CTA tile = 64 x 4*rows (2 warps across x, 4 across y)

// adjoints of the early-out results (warp_fwd_early_kernel): 1 out = add_flow*flow + src, 2 out = src, 3 out = add_flow*flow.
// Launched behind the adjoint kernel; when a rule fired it OVERWRITES gflow and gsrc (a broadcast payload's gradient is
// the sum over the batch).
__global__ void __launch_bounds__(256) warp_bwd_early_kernel(const WarpBwdParams p)
{
    const int sel = *p.sel;
    if (!sel) return;
    const int HW = p.H * p.W;
    const int64_t total = (int64_t)p.N * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / HW), pix = (int)(i - (int64_t)n * HW);
        const float *go = p.gout + (int64_t)n * p.C * HW;
        float rx = 0.0f, ry = 0.0f;
        for (int c = 0; c < p.C; ++c) {
            const float g = p.gout_bcast ? p.gout[0] : go[(int64_t)c * HW + pix];
            if (p.gsrc && p.gsrc_bs != 0) p.gsrc[(int64_t)n * p.gsrc_bs + (int64_t)c * HW + pix] = (sel != 3) ? g : 0.0f;
            if (sel != 2 && c == 0) rx = p.add_flow * g;
            if (sel != 2 && c == 1) ry = p.add_flow * g;
        }
        if (p.gsrc && p.gsrc_bs == 0 && n == 0) {
            for (int c = 0; c < p.C; ++c) {
                float sum = 0.0f;
                if (sel != 3)
                    for (int m = 0; m < p.N; ++m) sum += p.gout_bcast ? p.gout[0] : p.gout[((int64_t)m * p.C + c) * HW + pix];
                p.gsrc[(int64_t)c * HW + pix] = sum;
            }
        }
        if (p.gflow) {
            float *gf = p.gflow + (int64_t)n * 2 * HW;
            gf[pix] = rx;
            gf[HW + pix] = ry;
        }
    }
}

template <int CN, bool GFLOW, bool GSRC, bool WIN>
__global__ void __launch_bounds__(256, OFB_BWD_MINB) warp_t_bwd_kernel(const WarpBwdParams p, const CoordCtx cc, const int c_off,
                                                         const int accumulate_gflow)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 64 + (warp & 1) * 32 + lane;
    const int y0 = blockIdx.y * (4 * kBwdRows) + (warp >> 1) * kBwdRows;
    const int n = blockIdx.z;
    const int HW = p.H * p.W;
    const bool active = x < p.W;                       // inactive lanes still take part in the shuffles
    const int xc = min(x, p.W - 1);
    const int rows = min(kBwdRows, p.H - y0);          // warp-uniform
    if (rows <= 0) return;                             // whole warp below the field (the row prefetch must not run)

    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    const Window wn = WIN ? p.win : Window{0, 0, p.H, p.W, 0, 0, p.H, p.W};
    const int HWs = wn.sh * wn.sw, SW = wn.sw;   // stored source plane and its row pitch
    const float *src = p.src ? p.src + (int64_t)n * p.src_bs + (int64_t)c_off * HWs : nullptr;
    const float *go = p.gout + ((int64_t)n * p.C + c_off) * HW;
    float *gsrc = GSRC ? p.gsrc + (int64_t)n * p.gsrc_bs + (int64_t)c_off * HWs : nullptr;
    float *gf = GFLOW ? p.gflow + (int64_t)n * 2 * HW : nullptr;
    asm volatile("" : "+l"(fl), "+l"(src), "+l"(go), "+l"(gsrc), "+l"(gf));

    const float xf = (float)(x + wn.x0);
    float carry[CN];                                   // south-tap sums waiting for the next row
    int carry_o = 0;
    bool carry_ok = false;
#pragma unroll
    for (int c = 0; c < CN; ++c) carry[c] = 0.0f;

    // software pipeline: flow and upstream gradient of row r+1 are requested before row r is processed
    struct RowIn { float fx, fy, g[CN]; };
    auto load_row = [&](int r) {
        RowIn in;
        const int pix = (y0 + r) * p.W + xc;
        in.fx = __ldg(fl + pix);
        in.fy = __ldg(fl + (HW + pix));
#pragma unroll
        for (int c = 0; c < CN; ++c) in.g[c] = __ldg(go + (c * HW + pix));
        return in;
    };
    RowIn nxt = load_row(0);
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const RowIn cur = nxt;
        if (r + 1 < rows) nxt = load_row(r + 1);
        const int y = y0 + r;
        const int pix = y * p.W + xc;
        const float fx = cur.fx, fy = cur.fy;
        float ix, iy;
        sample_pos(cc, fx, fy, xf, (float)(y + wn.y0), ix, iy);
        PixGeom g = pix_geom<WIN>(ix, iy, wn);
        g.pnw = g.pnw && active; g.pne = g.pne && active; g.psw = g.psw && active; g.pse = g.pse && active;

        float gv[CN];
#pragma unroll
        for (int c = 0; c < CN; ++c) gv[c] = active ? cur.g[c] : 0.0f;

        if (GFLOW) {
            float gix = 0.0f, giy = 0.0f;
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const int e = c * HWs + g.o;
                float a, b, cq, d;
                load_taps<false, float>(src + e, src + (e + SW), g, a, b, cq, d);
                // d out / d ix and d out / d iy  (SURVEY.md Appendix A.1); gradients carry a tolerance, so FMAs
                gix = __fmaf_rn(gv[c], __fmaf_rn(b - a, g.sy, (d - cq) * g.ny), gix);
                giy = __fmaf_rn(gv[c], __fmaf_rn(cq - a, g.ex, (d - b) * g.wx), giy);
            }
            // d(ix)/d(flow_x) = -sign: the (W-1)/2 and 2/(W-1) factors of the chain cancel
            float rx = -p.sign * gix, ry = -p.sign * giy;
            if (p.add_flow != 0.0f) {                  // out[c] = add_flow*flow[c] + ...  (combine mode 3), c < 2
                if (c_off == 0) rx += p.add_flow * gv[0];
                if (c_off == 0 && CN > 1) ry += p.add_flow * gv[CN > 1 ? 1 : 0];
                if (c_off == 1) ry += p.add_flow * gv[0];
            }
            if (active) {
                if (accumulate_gflow) { rx += __ldcg(gf + pix); ry += __ldcg(gf + (HW + pix)); }
                __stcg(gf + pix, rx);
                __stcg(gf + (HW + pix), ry);
            }
        }

        if (GSRC) {
            // ---- horizontal merge with lane-1: its east taps are my west taps when the offsets chain
            const int prev_o = __shfl_up_sync(0xffffffffu, g.o, 1);
            const unsigned pe = (g.pne ? 1u : 0u) | (g.pse ? 2u : 0u);
            const unsigned prev_pe = __shfl_up_sync(0xffffffffu, pe, 1);
            const bool chain = (lane > 0) && (prev_o + 1 == g.o);
            const bool merge_n = chain && (prev_pe & 1u) && g.pnw;
            const bool merge_s = chain && (prev_pe & 2u) && g.psw;
            const unsigned bn = __ballot_sync(0xffffffffu, merge_n), bs = __ballot_sync(0xffffffffu, merge_s);
            const bool taken_n = (lane < 31) && ((bn >> (lane + 1)) & 1u);   // my NE was absorbed by lane+1
            const bool taken_s = (lane < 31) && ((bs >> (lane + 1)) & 1u);
            // ---- vertical merge: the carried south taps are this row's north taps when the offsets chain
            const bool vmerge = carry_ok && g.pnw && (carry_o == g.o);
            const bool flush = carry_ok && !vmerge;
            const bool east_n = g.pne && !taken_n, east_s = g.pse && !taken_s;
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const float v_ne = g.pne ? g.wne * gv[c] : 0.0f, v_se = g.pse ? g.wse * gv[c] : 0.0f;
                const float p_ne = __shfl_up_sync(0xffffffffu, v_ne, 1), p_se = __shfl_up_sync(0xffffffffu, v_se, 1);
                float top = g.pnw ? g.wnw * gv[c] : 0.0f, bot = g.psw ? g.wsw * gv[c] : 0.0f;
                if (merge_n) top += p_ne;
                if (merge_s) bot += p_se;
                const float *gs = gsrc + c * HWs;
                red_add<0>(gs + carry_o, carry[c], flush);
                if (vmerge) top += carry[c];
                const float *t = gs + g.o;
                red_add<0>(t, top, g.pnw);
                red_add<4>(t, v_ne, east_n);
                red_add<4>(t + SW, v_se, east_s);
                carry[c] = bot;
            }
            carry_o = g.o + SW;
            carry_ok = g.psw;
        }
    }
    if (GSRC && carry_ok) {
#pragma unroll
        for (int c = 0; c < CN; ++c) red_add<0>(gsrc + (c * HWs + carry_o), carry[c], true);
    }
}

typedef void (*BwdKernel)(const WarpBwdParams, const CoordCtx, const int, const int);

template <int CN, bool WIN>
static BwdKernel bwd_kernel_for(bool gflow, bool gsrc)
{
    if (gflow && gsrc) return warp_t_bwd_kernel<CN, true, true, WIN>;
    if (gflow) return warp_t_bwd_kernel<CN, true, false, WIN>;
    return warp_t_bwd_kernel<CN, false, true, WIN>;
}

static BwdKernel bwd_kernel(int cn, bool gflow, bool gsrc, bool win)
{
    switch (cn) {
        case 1: return win ? bwd_kernel_for<1, true>(gflow, gsrc) : bwd_kernel_for<1, false>(gflow, gsrc);
        case 2: return win ? bwd_kernel_for<2, true>(gflow, gsrc) : bwd_kernel_for<2, false>(gflow, gsrc);
        case 3: return win ? bwd_kernel_for<3, true>(gflow, gsrc) : bwd_kernel_for<3, false>(gflow, gsrc);
        default: return win ? bwd_kernel_for<4, true>(gflow, gsrc) : bwd_kernel_for<4, false>(gflow, gsrc);
    }
}

}  // namespace ofb

using namespace ofb;

static bool window_ok(const Window &w, int H, int W)
{
    return H >= 1 && W >= 1 && w.gh >= 2 && w.gw >= 2 && w.sh >= 1 && w.sw >= 1 && w.y0 >= 0 && w.x0 >= 0 &&
           w.y0 + H <= w.gh && w.x0 + W <= w.gw && w.sy0 >= 0 && w.sx0 >= 0 && w.sy0 + w.sh <= w.gh &&
           w.sx0 + w.sw <= w.gw && (int64_t)w.sh * w.sw <= 0x1fffffff && (int64_t)H * W <= 0x1fffffff;   // 4 planes of 32-bit offsets
}

static bool window_is_identity(const Window &w, int H, int W)
{
    return w.y0 == 0 && w.x0 == 0 && w.gh == H && w.gw == W && w.sy0 == 0 && w.sx0 == 0 && w.sh == H && w.sw == W;
}

static int warp_t_fwd_impl(const float *flow, int64_t flow_bs, float flow_sign,
                           const float *src, int64_t src_bs,
                           const uint8_t *src_mask, int64_t src_mask_bs,
                           const uint8_t *flow_mask, int64_t flow_mask_bs,
                           float add_flow, float valid_thresh,
                           float *out, uint8_t *valid,
                           int N, int C, int H, int W, const Window &win, void *stream,
                           const int32_t *sel = nullptr, int32_t *out_flags = nullptr, float *dtaps = nullptr)
{
    if (!flow || N <= 0 || C < 0 || (C > 0 && (!src || !out)) || (C == 0 && !valid)) return OFB_EINVAL;
    if (out_flags && (C != 2 || !valid)) return OFB_EINVAL;       // flags of a result: flow-valued results only
    if (add_flow != 0.0f && C != 2) return OFB_EINVAL;
    if (win.gh < 2 || win.gw < 2) return OFB_ESHAPE;
    if (!window_ok(win, H, W) || N > 65535) return OFB_EINVAL;
    WarpFwdParams p{flow, flow_bs, flow_sign, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs,
                    add_flow, valid_thresh, out, valid, N, C, H, W, win, sel, out_flags, dtaps};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CoordCtx cc{{(float)(win.gw - 1), 1.0f / (float)(win.gw - 1)}, {(float)(win.gh - 1), 1.0f / (float)(win.gh - 1)},
                      flow_sign};

    // Measured on B200 (profiles/): the L1-cached direct kernel beats the staged one on every flow tried (bounding-box
    // staging over-fetches (1+g)^2 for a flow gradient g, and its load->TMA->compute phases do not overlap inside a
    // CTA), so the staged kernel is opt-in (OFB_TMA=1) until it is software-pipelined across tiles.
    static const bool use_tma = getenv("OFB_TMA") != nullptr;
    bool staged = use_tma && !dtaps && C > 0 && W >= kTileW && H >= kTileH && window_is_identity(win, H, W) && !sel && !out_flags;
    const int vmode = !valid ? 0 : (src_mask ? 2 : 1);
    alignas(64) CUtensorMap mask_map;
    alignas(64) CUtensorMap src_maps[2];          // [0]: groups of kMaxCN channels, [1]: the remainder group
    const int groups = (C + kMaxCN - 1) / kMaxCN;
    const int last_cn = C - (groups - 1) * kMaxCN;
    if (staged) {
        const uint64_t planes = (uint64_t)((src_bs == 0) ? 1 : N) * C;
        const int first_cn = groups > 1 ? kMaxCN : C;
        staged = make_map_3d(&src_maps[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, W, H, planes, kBoxW, kBoxH, first_cn);
        if (staged && groups > 1 && last_cn != kMaxCN)
            staged = make_map_3d(&src_maps[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, W, H, planes, kBoxW, kBoxH,
                                 last_cn);
        else
            src_maps[1] = src_maps[0];
        if (staged && vmode == 2)
            staged = make_map_3d(&mask_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, src_mask, W, H,
                                 (uint64_t)((src_mask_bs == 0) ? 1 : N), kMBoxW, kBoxH, 1);
        else
            mask_map = src_maps[0];
    }
    static const bool use_pipe = getenv("OFB_TMA") != nullptr && atoi(getenv("OFB_TMA")) == 2;
    if (staged && use_pipe && groups == 1) {
        const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
        const int64_t ntiles = (int64_t)tiles_x * tiles_y * N;
        if (ntiles <= 0x7fffffff) {
            PipeKernel k = pipe_kernel(C, vmode);
            const int stage = kPlane * 4 * C + (vmode == 2 ? kMBoxW * kBoxH : 0);
            const int smem = 2 * ((stage + 1023) & ~1023) + (int)sizeof(PipeCtl) + 64;
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            const int ctas = (int)(ntiles < 148 * 2 ? ntiles : 148 * 2);
            k<<<ctas, 256, smem, st>>>(p, src_maps[0], mask_map, cc, tiles_x, tiles_y);
            OFB_CHECK_LAUNCH();
            return 0;
        }
    }
    if (staged) {
        const dim3 grid((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH, N);
        for (int g = 0; g < groups; ++g) {
            const int cn = (g == groups - 1) ? last_cn : kMaxCN;
            const int vm = (g == 0) ? vmode : 0;              // the validity bit is written once
            StagedKernel k = staged_kernel(cn, vm);
            const int smem = staged_smem_bytes(cn, vm);
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            k<<<grid, 256, smem, st>>>(p, (g == groups - 1 && groups > 1) ? src_maps[1] : src_maps[0], mask_map,
                                       g * kMaxCN, cc);
            OFB_CHECK_LAUNCH();
        }
        return 0;
    }
    {
        const dim3 grid((W + 63) / 64, (H + 4 * kDirRows - 1) / (4 * kDirRows), N);
        const bool wv = valid != nullptr;
        typedef void (*DirectKernel)(const WarpFwdParams, const CoordCtx);
        DirectKernel k;
        const int cn = (C >= 1 && C <= 4) ? C : 0;     // 0 = runtime channel loop (also C == 0: validity only)
#define OFB_PICK(CN_) (wv || C == 0 ? (win_id ? warp_t_fwd_direct_kernel<CN_, true, false> : warp_t_fwd_direct_kernel<CN_, true, true>) \
                                    : (win_id ? warp_t_fwd_direct_kernel<CN_, false, false> : warp_t_fwd_direct_kernel<CN_, false, true>))
        const bool win_id = window_is_identity(win, H, W);
        switch (cn) {
            case 1: k = OFB_PICK(1); break;
            case 2: k = OFB_PICK(2); break;
            case 3: k = OFB_PICK(3); break;
            case 4: k = OFB_PICK(4); break;
            default: k = OFB_PICK(0); break;
        }
#undef OFB_PICK
        if (out_flags) {
            if (!win_id) return OFB_EINVAL;
            k = warp_t_fwd_direct_kernel<2, true, false, true>;
        }
        dim3 lgrid = grid;
        // (with out_flags -- the result's flags, combine mode 3 -- the direct kernel is the faster one: the lean kernel
        // spills at 32 registers once it carries the two maxima, measured 1.31 against 1.42 ms at 16 x 4K)
        const bool lean = (g_fwd_impl == 0 || dtaps) && (!out_flags || dtaps) && (win_id || !dtaps) && cn >= 1 &&
                          kLeanRows == kDirRows && win.gw <= 65536 && win.gh <= 65536;
        if (dtaps && !lean) return OFB_EINVAL;
        if (lean) {     // the plain warp: lean kernel
#define OFB_LEAN(CN_) (dtaps ? (wv ? warp_t_fwd_lean_kernel<CN_, true, false, true> : warp_t_fwd_lean_kernel<CN_, false, false, true>) \
                       : !win_id ? (wv ? warp_t_fwd_lean_kernel<CN_, true, false, false, true> : warp_t_fwd_lean_kernel<CN_, false, false, false, true>) \
                             : (wv ? warp_t_fwd_lean_kernel<CN_, true, false, false> : warp_t_fwd_lean_kernel<CN_, false, false, false>))
            switch (cn) {
                case 1: k = OFB_LEAN(1); break;
                case 2: k = OFB_LEAN(2); break;
                case 3: k = OFB_LEAN(3); break;
                default: k = OFB_LEAN(4); break;
            }
#undef OFB_LEAN
            if (out_flags) k = warp_t_fwd_lean_kernel<2, true, true, true>;     // only with saved derivatives
        }
        k<<<lgrid, 256, 0, st>>>(p, cc);
    }
    OFB_CHECK_LAUNCH();
    if (sel) {
        warp_fwd_early_kernel<<<148 * 8, 256, 0, st>>>(p);
        OFB_CHECK_LAUNCH();
    }
    if (out_flags) {
        stats_finish_kernel<<<N, 256, 0, st>>>(out_flags, N, ((W + 63) / 64) * ((H + 4 * kDirRows - 1) / (4 * kDirRows)));
        OFB_CHECK_LAUNCH();
    }
    return 0;
}

static int warp_t_bwd_impl(const float *flow, int64_t flow_bs, float flow_sign,
                           const float *src, int64_t src_bs,
                           const float *gout, float add_flow,
                           float *gflow, float *gsrc, int64_t gsrc_bs,
                           int N, int C, int H, int W, const Window &win, void *stream, const int32_t *sel = nullptr,
                           const float *dtaps = nullptr, int gout_bcast = 0)
{
    if (!flow || !gout || N <= 0 || C <= 0 || (gflow && !src && !dtaps)) return OFB_EINVAL;
    if (dtaps && C > 4) return OFB_EINVAL;
    if (add_flow != 0.0f && C != 2) return OFB_EINVAL;
    if (win.gh < 2 || win.gw < 2) return OFB_ESHAPE;
    if (!window_ok(win, H, W) || N > 65535) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!gflow && !gsrc) return 0;
    const CoordCtx cc{{(float)(win.gw - 1), 1.0f / (float)(win.gw - 1)}, {(float)(win.gh - 1), 1.0f / (float)(win.gh - 1)},
                      flow_sign};
    bool done = false;
    if (gsrc) {
        const WarpBwdParams rp{flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W, win,
                               nullptr, dtaps, gout_bcast};
        // tile-owned accumulation (warp_t_tile.cu), the default; 1 = outside its domain
        int rs = warp_t_bwd_tile_launch(rp, cc, st);
        if (rs == 0) done = true;
        else if (rs != 1) return rs;
        if (!done && !sel && window_is_identity(win, H, W)) {      // ring adjoint (warp_t_adj.cu), opt-in experiment
            rs = warp_t_bwd_ring_launch(rp, cc, st);
            if (rs != 1) return rs;
        }
    }
    if (!done && ((gflow && !src) || gout_bcast)) return OFB_EINVAL;      // the other kernels gather the payload taps and read a dense gout
    if (!done) {
        // Optional L2-resident accumulation (OFB_BWD_CHUNK_MB > 0): walk the batch in chunks whose gradient image fits the
        // L2, zero-filling each chunk right before its kernel so the REDs hit L2-resident lines (-24 B/pixel of DRAM
        // traffic).  Measured on B200: a LOSS as separate launches (1.47 -> 1.58/1.75/1.88 ms at 96/48/24 MB chunks,
        // tails and launch gaps dominate), hence off by default; it needs the zero-fill fused into one persistent kernel.
        static const long chunk_mb = getenv("OFB_BWD_CHUNK_MB") ? atol(getenv("OFB_BWD_CHUNK_MB")) : 0;
        const int64_t src_plane = (int64_t)win.sh * win.sw, out_plane = (int64_t)H * W;
        const int64_t item_bytes = (int64_t)C * src_plane * 4;
        int chunk = N;
        if (gsrc && gsrc_bs != 0 && chunk_mb > 0) {
            chunk = (int)((chunk_mb << 20) / item_bytes);
            if (chunk < 1) chunk = 1;
            if (chunk > N) chunk = N;
        }
        if (gsrc && gsrc_bs == 0) {                          // broadcast payload: one image accumulates the whole batch
            cudaError_t e = cudaMemsetAsync(gsrc, 0, (size_t)item_bytes, st);
            if (e != cudaSuccess) return (int)e;
        }
        for (int n0 = 0; n0 < N; n0 += chunk) {
            const int nn = (N - n0 < chunk) ? (N - n0) : chunk;
            if (gsrc && gsrc_bs != 0) {
                cudaError_t e = cudaMemsetAsync(gsrc + (int64_t)n0 * gsrc_bs, 0, (size_t)item_bytes * nn, st);
                if (e != cudaSuccess) return (int)e;
            }
            WarpBwdParams p{flow + (int64_t)n0 * flow_bs, flow_bs, flow_sign, src ? src + (int64_t)n0 * src_bs : nullptr, src_bs,
                            gout + (int64_t)n0 * C * out_plane, add_flow, gflow ? gflow + (int64_t)n0 * 2 * out_plane : nullptr,
                            gsrc ? gsrc + (int64_t)n0 * gsrc_bs : nullptr, gsrc_bs, nn, C, H, W, win, sel};
            const dim3 grid((W + 63) / 64, (H + 4 * kBwdRows - 1) / (4 * kBwdRows), nn);
            for (int c0 = 0; c0 < C; c0 += kMaxCN) {       // channel groups of <= 4; gflow accumulates across groups
                const int cn = (C - c0 < kMaxCN) ? (C - c0) : kMaxCN;
                bwd_kernel(cn, gflow != nullptr, gsrc != nullptr, !window_is_identity(win, H, W))<<<grid, 256, 0, st>>>(
                    p, cc, c0, c0 > 0 ? 1 : 0);
                OFB_CHECK_LAUNCH();
            }
        }
    }
    if (sel) {
        const WarpBwdParams p{flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W, win,
                              sel, nullptr, gout_bcast};
        warp_bwd_early_kernel<<<148 * 8, 256, 0, st>>>(p);
        OFB_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int ofb_warp_t_fwd(const float *flow, int64_t flow_bs, float flow_sign,
                              const float *src, int64_t src_bs,
                              const uint8_t *src_mask, int64_t src_mask_bs,
                              const uint8_t *flow_mask, int64_t flow_mask_bs,
                              float add_flow, float valid_thresh,
                              float *out, uint8_t *valid,
                              int N, int C, int H, int W, void *stream)
{
    if (H < 2 || W < 2) return OFB_ESHAPE;
    return warp_t_fwd_impl(flow, flow_bs, flow_sign, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs,
                           add_flow, valid_thresh, out, valid, N, C, H, W, Window{0, 0, H, W, 0, 0, H, W}, stream);
}

extern "C" int ofb_warp_t_bwd(const float *flow, int64_t flow_bs, float flow_sign,
                              const float *src, int64_t src_bs,
                              const float *gout, float add_flow,
                              float *gflow, float *gsrc, int64_t gsrc_bs,
                              int N, int C, int H, int W, void *stream)
{
    if (H < 2 || W < 2) return OFB_ESHAPE;
    return warp_t_bwd_impl(flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W,
                           Window{0, 0, H, W, 0, 0, H, W}, stream);
}

// device-side early-out decision and the _ex variants that honour it (see include/oflib_b200.h)
extern "C" int ofb_early_decide(const int32_t *flow_flags, const int32_t *payload_flags, int rules, int N, int32_t *sel,
                                void *stream)
{
    if (!sel || N <= 0) return OFB_EINVAL;
    early_decide_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(flow_flags, payload_flags, rules, N, sel);
    OFB_CHECK_LAUNCH();
    return 0;
}

extern "C" int64_t ofb_warp_t_flags_ints(int N, int H, int W)
{
    const int64_t ctas = (int64_t)((W + 63) / 64) * ((H + 4 * kDirRows - 1) / (4 * kDirRows));
    return (int64_t)N + 1 + 2 * ctas * N;
}

extern "C" int ofb_warp_t_fwd_ex(const float *flow, int64_t flow_bs, float flow_sign,
                                 const float *src, int64_t src_bs,
                                 const uint8_t *src_mask, int64_t src_mask_bs,
                                 const uint8_t *flow_mask, int64_t flow_mask_bs,
                                 float add_flow, float valid_thresh,
                                 float *out, uint8_t *valid,
                                 int N, int C, int H, int W,
                                 const int32_t *sel, int32_t *out_flags, void *stream)
{
    if (H < 2 || W < 2) return OFB_ESHAPE;
    return warp_t_fwd_impl(flow, flow_bs, flow_sign, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs,
                           add_flow, valid_thresh, out, valid, N, C, H, W, Window{0, 0, H, W, 0, 0, H, W}, stream,
                           sel, out_flags);
}

extern "C" int ofb_warp_t_bwd_ex(const float *flow, int64_t flow_bs, float flow_sign,
                                 const float *src, int64_t src_bs,
                                 const float *gout, float add_flow,
                                 float *gflow, float *gsrc, int64_t gsrc_bs,
                                 int N, int C, int H, int W,
                                 const int32_t *sel, void *stream)
{
    if (H < 2 || W < 2) return OFB_ESHAPE;
    return warp_t_bwd_impl(flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W,
                           Window{0, 0, H, W, 0, 0, H, W}, stream, sel);
}

// ofb_warp_t_fwd_ex that also saves d out / d(ix, iy) for the adjoint, and the adjoint that consumes it
extern "C" int ofb_warp_t_fwd_d(const float *flow, int64_t flow_bs, float flow_sign,
                                const float *src, int64_t src_bs,
                                const uint8_t *src_mask, int64_t src_mask_bs,
                                const uint8_t *flow_mask, int64_t flow_mask_bs,
                                float add_flow, float valid_thresh,
                                float *out, uint8_t *valid,
                                int N, int C, int H, int W,
                                const int32_t *sel, int32_t *out_flags, float *dtaps, void *stream)
{
    if (H < 2 || W < 2) return OFB_ESHAPE;
    if (dtaps && (C < 1 || C > 4)) return OFB_EINVAL;
    return warp_t_fwd_impl(flow, flow_bs, flow_sign, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs,
                           add_flow, valid_thresh, out, valid, N, C, H, W, Window{0, 0, H, W, 0, 0, H, W}, stream,
                           sel, out_flags, dtaps);
}

extern "C" int ofb_warp_t_bwd_d(const float *flow, int64_t flow_bs, float flow_sign,
                                const float *src, int64_t src_bs,
                                const float *gout, float add_flow,
                                float *gflow, float *gsrc, int64_t gsrc_bs,
                                int N, int C, int H, int W,
                                const int32_t *sel, const float *dtaps, int gout_broadcast, void *stream)
{
    if (H < 2 || W < 2) return OFB_ESHAPE;
    return warp_t_bwd_impl(flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W,
                           Window{0, 0, H, W, 0, 0, H, W}, stream, sel, dtaps, gout_broadcast ? 1 : 0);
}

// Windowed variants: window[8] = {y0, x0, global_h, global_w, src_y0, src_x0, src_h, src_w} (see struct Window).
extern "C" int ofb_warp_t_win_fwd(const float *flow, int64_t flow_bs, float flow_sign,
                                  const float *src, int64_t src_bs,
                                  const uint8_t *src_mask, int64_t src_mask_bs,
                                  const uint8_t *flow_mask, int64_t flow_mask_bs,
                                  float add_flow, float valid_thresh,
                                  float *out, uint8_t *valid,
                                  int N, int C, int H, int W, const int *window, void *stream)
{
    if (!window) return OFB_EINVAL;
    const Window w{window[0], window[1], window[2], window[3], window[4], window[5], window[6], window[7]};
    return warp_t_fwd_impl(flow, flow_bs, flow_sign, src, src_bs, src_mask, src_mask_bs, flow_mask, flow_mask_bs,
                           add_flow, valid_thresh, out, valid, N, C, H, W, w, stream);
}

extern "C" int ofb_warp_t_win_bwd(const float *flow, int64_t flow_bs, float flow_sign,
                                  const float *src, int64_t src_bs,
                                  const float *gout, float add_flow,
                                  float *gflow, float *gsrc, int64_t gsrc_bs,
                                  int N, int C, int H, int W, const int *window, void *stream)
{
    if (!window) return OFB_EINVAL;
    const Window w{window[0], window[1], window[2], window[3], window[4], window[5], window[6], window[7]};
    return warp_t_bwd_impl(flow, flow_bs, flow_sign, src, src_bs, gout, add_flow, gflow, gsrc, gsrc_bs, N, C, H, W, w,
                           stream);
}
