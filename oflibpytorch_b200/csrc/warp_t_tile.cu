// W adjoint, tile-owned accumulation (the default, for the plain warp and for windowed frames).
//
// Adjoint of reference utils.py:541-555 (grid_sample bilinear / zeros / align_corners=True, i.e. ATen's
// grid_sampler_2d_backward): gflow is a gather, gsrc the un-normalised splat of w * gout.
//
// What bounded the register-merge kernel (warp_t.cu) was measured with ncu: every red.global costs the SM's L1 -> crossbar
// port ONE request per 32-byte sector it touches, however few lanes are in it (385 M sectors = 13.4 GB of crossbar writes
// for 1.6 GB of gradient, l1tex__m_l1tex2xbar_req_cycles_active 70 %), and merging taps in registers reduces lanes, not
// sectors.  Here a CTA owns a 32 x 32 tile of OUTPUT pixels and accumulates all its 4*C contributions per pixel into a
// shared-memory window (64 x 52 per channel) that covers the tile's footprint in the source image, then flushes the
// window ONCE with dense 16-byte REDs, skipping empty quads: 17 RED sectors per 32 pixels instead of 93 (ncu).
//
// Float atomics on shared memory are CAS loops on sm_100 (ATOMS.CAST.SPIN); 32-bit integer adds are native
// (ATOMS.ADD, about 2 clk per warp-op, tools/microbench.cu).  The window therefore accumulates in FIXED POINT with a
// per-tile power-of-two scale: amax = max |gout| over the tile, amax in [2^E, 2^(E+1)), scale 2^(19-E).  A source pixel
// receives at most one tap of each of the tile's 1024 pixels, each |w*g| < 2^(E+1), so the sum stays below 2^31 for
// every input; one contribution is rounded to 2^(E-19) <= 1.9e-6 * amax (the fp32 REDs it replaces round at 6e-8 of the
// running sum; the parity bar for gradients is 1e-4 of the gradient scale).  Integer addition is associative: the
// tile's part of the sum does not depend on the order of arrival.  Tiles whose footprint does not fit the window (violent
// flows) and tiles with a non-finite upstream gradient issue float REDs to global memory directly, per tap.
#include <climits>

#include "common.cuh"
#include "warp_t.cuh"

namespace ofb {

extern int g_bwd_impl;
extern int g_bwd_za;

constexpr int kTT = 32;                    // tile edge (output pixels): 8 warps x 4 rows, lane = column
constexpr int kTRows = 4;                  // rows per thread
#ifndef OFB_TILE_MINB
#define OFB_TILE_MINB 4
#endif
#ifndef OFB_TILE_WINH
#define OFB_TILE_WINH 52
#endif
// Window geometry.  32 x 32 tiles of the benchmark's flow field map to about 37 x 37 +- 8 source pixels.
//   pitch 64: a multiple of 32, so lanes on neighbouring source rows never share a bank (their columns differ by less
//             than 32); a 56-wide window had 105 M bank-conflict wavefronts per launch, 25 per 32 pixels.
//   52 rows:  covers 99 % of the tiles in 40 KB for three channels.  What matters is that four resident CTAs leave the
//             SM enough L1 for the tap gathers of the flow gradient -- measured per 32 x 3 x 1080p launch: 64 x 64
//             (48 KB) 1.28 ms, 56 x 56 1.06, 64 x 56 1.06, 64 x 52 1.03, 64 x 48 1.04, 64 x 44 1.06 (fall-backs grow);
//             a window of fixed capacity whose pitch follows the footprint: 1.10.
constexpr int kTWin = 64;                  // window row pitch (ints)
constexpr int kTWinH = OFB_TILE_WINH;      // window rows
constexpr int kTPlane = kTWin * kTWinH;    // ints per channel

#ifndef OFB_TILE_PRE
#define OFB_TILE_PRE (-1)   // rows whose payload taps are loaded ahead of the barriers; -1: 2 for C <= 2, 1 for C = 3, 0 for C = 4
#endif
#ifndef OFB_TILE_LD
#define OFB_TILE_LD __ldg
#endif

struct TileCtl {
    int minx, maxx, miny, maxy;
    unsigned amax;
};

__device__ __forceinline__ void red_add_f32(float *a, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(v));
}

__device__ __forceinline__ void red_add_v4(float *a, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w));
}

// GFLOW: 0 no flow gradient, 1 from the payload taps (global gathers), 2 from the derivative planes the forward saved
// WIN:   the flow / gout / gflow tensors are a window of a larger frame and src / gsrc a stored region of it (struct Window,
//        warp_t.cuh: padded apply with cut, row bands); everything about the taps then lives in stored-region coordinates
#ifndef OFB_TILE_ZROWS
#define OFB_TILE_ZROWS 4
#endif
constexpr int kZRows = OFB_TILE_ZROWS;     // zero-ahead distance in tile rows

__device__ __forceinline__ int ld_acquire(const int *a)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}

// ZA ("zero ahead", opt-in: ofb_set_option(6, 1)): the kernel zero-fills the payload gradient itself.  Tile rows are
// numbered through the batch (R = n * tiles_y + ty); the CTA of tile (R, tx) first stores zeros over the region of tile
// (R + kZRows, tx) and, behind its first barrier, one thread fences and bumps za[R + kZRows]; before a CTA issues REDs it
// waits until every tile row its footprint touches is complete (za[R'] == tiles_x).  CTAs are dispatched in launch
// order, so the rows waited for belong to CTAs that are already running and whose zero-fill does not wait for anything
// (the forward-progress argument of a decoupled look-back scan).  The host zero-fills the first kZRows tile rows; a tile
// whose footprint reaches kZRows or more tile rows DOWN cannot wait (those rows belong to later CTAs): it leaves its
// payload-gradient part to warp_t_bwd_tile_deferred_kernel.
// Measured (ncu, 32 x 3 x 1080p): it does what it was built for -- the zeros are still in L2 when the REDs arrive, DRAM
// reads 2.91 -> 2.13 GB (no read for the read-modify-write), writes unchanged, no 0.8 GB memset -- and it is SLOWER:
// 1.08 ms against 0.93 + 0.108 (memset) = 1.03 ms; without the waits (wrong results) 1.03 ms.  The kernel is not bound by
// DRAM bytes, and the zero stores cost inside it what they cost the memset.  Hence opt-in, not the default.
template <int CN, int GFLOW, bool WIN, bool ZA = false>
__global__ void __launch_bounds__(256, OFB_TILE_MINB) warp_t_bwd_tile_kernel(const WarpBwdParams p, const CoordCtx cc, const int c_off,
                                                                 const int accumulate_gflow, const int vec_ok)
{
    extern __shared__ __align__(16) int win[];                 // [CN][kTWin][kTWin]
    __shared__ TileCtl ctl;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * kTT + lane;
    const int y0 = blockIdx.y * kTT + warp * kTRows;
    const int n = blockIdx.z;
    const int W = p.W, H = p.H, HW = H * W;                                     // flow / output window
    const int SW = WIN ? p.win.sw : W, SH = WIN ? p.win.sh : H, HWs = SH * SW;  // stored source region
    const int X0 = WIN ? p.win.x0 : 0, Y0 = WIN ? p.win.y0 : 0;                 // window origin in the global frame
    const int SX0 = WIN ? p.win.sx0 : 0, SY0 = WIN ? p.win.sy0 : 0;             // stored region origin
    const bool xin = x < W;
    const int xc = min(x, W - 1);

    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    const float *src = GFLOW == 1 ? p.src + (int64_t)n * p.src_bs + (int64_t)c_off * HWs : nullptr;
    const float *dt = GFLOW == 2 ? p.dtaps + ((int64_t)n * p.C + c_off) * 2 * HW : nullptr;
    const float *go = p.gout_bcast ? p.gout : p.gout + ((int64_t)n * p.C + c_off) * HW;
    float *gsrc = p.gsrc + (int64_t)n * p.gsrc_bs + (int64_t)c_off * HWs;
    float *gf = GFLOW ? p.gflow + (int64_t)n * 2 * HW : nullptr;
    asm volatile("" : "+l"(src), "+l"(dt), "+l"(go), "+l"(gsrc), "+l"(gf));

    // ---- phase 0: clear the window (under the latency of phase 1's loads; clearing only the footprint once it is known
    // was measured 3 % slower: one more barrier, and the stores no longer overlap anything), initialise the control block
    {
        int4 *w4 = reinterpret_cast<int4 *>(win);
#pragma unroll
        for (int i = threadIdx.x; i < CN * kTPlane / 4; i += 256) w4[i] = make_int4(0, 0, 0, 0);
        if (threadIdx.x == 0) {
            ctl.minx = INT_MAX; ctl.miny = INT_MAX; ctl.maxx = INT_MIN; ctl.maxy = INT_MIN; ctl.amax = 0u;
        }
    }
    const int Rg = ZA ? n * (int)gridDim.y + (int)blockIdx.y : 0;              // this tile's row, counted through the batch
    bool za_signal = false;
    if (ZA) {
        const int Rt = Rg + kZRows;
        if (Rt < (int)(gridDim.y * gridDim.z)) {
            const int nz = Rt / (int)gridDim.y, tyz = Rt - nz * (int)gridDim.y;
            const int zy = tyz * kTT + (threadIdx.x >> 3), zx = blockIdx.x * kTT + (threadIdx.x & 7) * 4;
            if (zy < H && zx < W) {
                float *z = p.gsrc + (int64_t)nz * p.gsrc_bs + (zy * W + zx);
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    if (vec_ok) {
                        *reinterpret_cast<float4 *>(z + c * HW) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    } else {
                        for (int k = 0; k < 4; ++k)
                            if (zx + k < W) z[c * HW + k] = 0.0f;
                    }
                }
            }
            za_signal = true;
        }
    }

    // ---- phase 1: inputs of the thread's rows, sample positions, footprint and largest |gout| of the tile
    float ix[kTRows], iy[kTRows], gv[kTRows][CN];
    // Payload taps of the thread's first row(s), loaded AHEAD of the two barriers of the footprint reduction so that their
    // latency overlaps the barrier waits (C = 3: 1.033 -> 1.017 ms, C = 2: 0.811 -> 0.773 ms at 32 x 1080p; a second row
    // spills at C = 3: 1.097 ms).
    constexpr int kPre = (GFLOW != 1 || WIN || ZA) ? 0 : (OFB_TILE_PRE >= 0 ? OFB_TILE_PRE : (CN <= 2 ? 2 : (CN == 3 ? 1 : 0)));
    float pre[kPre > 0 ? kPre : 1][CN][4];
    {
        float fx[kTRows], fy[kTRows];
#pragma unroll
        for (int r = 0; r < kTRows; ++r) {
            const int q = min(y0 + r, H - 1) * W + xc;
            fx[r] = OFB_TILE_LD(fl + q);
            fy[r] = OFB_TILE_LD(fl + (HW + q));
#pragma unroll
            for (int c = 0; c < CN; ++c) gv[r][c] = OFB_TILE_LD(go + (p.gout_bcast ? 0 : c * HW + q));
        }
        const float xf = (float)(x + X0);
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
        unsigned amax = 0u;
#pragma unroll
        for (int r = 0; r < kTRows; ++r) {
            const bool act = xin && (y0 + r < H);
            ix[r] = round_trip_unchecked(__fmaf_rn(-cc.sign, fx[r], xf), cc.dw);
            iy[r] = round_trip_unchecked(__fmaf_rn(-cc.sign, fy[r], (float)(y0 + r + Y0)), cc.dh);
            // tap position in the stored source region (which lies inside the global frame: only its bounds matter)
            const int xi = (int)((unsigned)__float2int_rd(ix[r]) - (unsigned)SX0);
            const int yi = (int)((unsigned)__float2int_rd(iy[r]) - (unsigned)SY0);
            // pixels with at least one tap inside the region stretch the footprint, by those taps only
            const bool any = act && xi >= -1 && xi <= SW - 1 && yi >= -1 && yi <= SH - 1;
            if (any) {
                minx = min(minx, max(xi, 0)); maxx = max(maxx, min(xi + 1, SW - 1));
                miny = min(miny, max(yi, 0)); maxy = max(maxy, min(yi + 1, SH - 1));
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                if (!act) gv[r][c] = 0.0f;
                amax = max(amax, __float_as_uint(gv[r][c]) & 0x7fffffffu);
            }
        }
        minx = __reduce_min_sync(0xffffffffu, minx); maxx = __reduce_max_sync(0xffffffffu, maxx);
        miny = __reduce_min_sync(0xffffffffu, miny); maxy = __reduce_max_sync(0xffffffffu, maxy);
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (kPre > 0) {
#pragma unroll
            for (int r = 0; r < kPre; ++r) {
                const int xi0 = (int)((unsigned)__float2int_rd(ix[r]) - (unsigned)SX0);
                const int yi0 = (int)((unsigned)__float2int_rd(iy[r]) - (unsigned)SY0);
                const bool in0 = (unsigned)xi0 < (unsigned)(SW - 1) && (unsigned)yi0 < (unsigned)(SH - 1);
                if (__all_sync(0xffffffffu, in0 && xin) && y0 + r < H) {       // the predicate of phase 2's fast path
                    const int o0 = yi0 * SW + xi0;
#pragma unroll
                    for (int c = 0; c < CN; ++c) {
                        const float *tn = src + (c * HWs + o0);
                        pre[r][c][0] = __ldg(tn); pre[r][c][1] = __ldg(tn + 1);
                        pre[r][c][2] = __ldg(tn + SW); pre[r][c][3] = __ldg(tn + SW + 1);
                    }
                }
            }
        }
        __syncthreads();                                   // the control block is initialised
        // release by ONE thread behind the barrier (the pattern of a cooperative-groups grid sync: the fence is cumulative
        // over what the barrier ordered before it); a fence in every thread showed up as 7 % membar stalls
        if (ZA && za_signal && threadIdx.x == 0) { __threadfence(); atomicAdd(p.za + (Rg + kZRows), 1); }
        if (lane == 0) {
            if (minx != INT_MAX) {
                atomicMin(&ctl.minx, minx); atomicMax(&ctl.maxx, maxx);
                atomicMin(&ctl.miny, miny); atomicMax(&ctl.maxy, maxy);
            }
            atomicMax(&ctl.amax, amax);
        }
        __syncthreads();
    }
    const int ox = ctl.minx & ~3, oy = ctl.miny;           // window origin (x a multiple of 4: 16-byte REDs at the flush)
    const int ex_w = ctl.maxx - ox + 1, ex_h = ctl.maxy - oy + 1;
    const unsigned amax = ctl.amax;
    const bool nothing = ctl.minx == INT_MAX || amax == 0u;                 // no tap in the image, or a zero gradient
    // fixed point needs the footprint inside the window and a finite gradient; otherwise float REDs per tap
    bool tiled = !nothing && ex_w <= kTWin && ex_h <= kTWinH && amax < 0x7f800000u;
    bool za_defer = false;
    if (ZA && !nothing) {
        const int r0 = n * (int)gridDim.y + (ctl.miny >> 5), r1 = n * (int)gridDim.y + (ctl.maxy >> 5);
        za_defer = r1 >= Rg + kZRows;                       // CTA-uniform
        if (za_defer) {
            if (threadIdx.x == 0) {
                const int slot = atomicAdd(p.za + (gridDim.y * gridDim.z), 1);
                p.za[gridDim.y * gridDim.z + 1 + slot] = (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
            }
        } else {
            for (int R = max(r0, kZRows) + (int)threadIdx.x; R <= r1; R += 256)
                while (ld_acquire(p.za + R) < (int)gridDim.x) {}
            if (!tiled) __syncthreads();                    // the float REDs of phase 2 come before the next barrier
        }
    }
    if (za_defer) tiled = false;
    const bool skip_red = nothing || za_defer;              // no contribution to the payload gradient from this CTA
    // amax in [2^E, 2^(E+1)), E = e - 127 clamped to >= -100: scale 2^(19-E), inverse 2^(E-19)
    const int e = max((int)(amax >> 23), 27);
    const float scale = __uint_as_float((unsigned)(127 + 19 + 127 - e) << 23);
    const float inv_scale = __uint_as_float((unsigned)(e - 19) << 23);

    const int fq = threadIdx.x & 15, fr0 = threadIdx.x >> 4;           // flush: 16 quads per window row, 16 rows per pass
    const int nq = (ex_w + 3) >> 2;

    // ---- phase 2: flow gradient (gather) and the contributions to the payload gradient
    const int rows = min(kTRows, H - y0);                   // warp-uniform; <= 0 below the field
#pragma unroll
    for (int r = 0; r < kTRows; ++r) {
        if (r >= rows) break;
        const int pix = (y0 + r) * W + xc;
        const int xi = (int)((unsigned)__float2int_rd(ix[r]) - (unsigned)SX0);      // stored-region coordinates
        const int yi = (int)((unsigned)__float2int_rd(iy[r]) - (unsigned)SY0);
        const float x0 = floorf(ix[r]), y0f = floorf(iy[r]);
        const float ex = __fsub_rn(__fadd_rn(x0, 1.0f), ix[r]), wx = __fsub_rn(ix[r], x0);
        const float sy = __fsub_rn(__fadd_rn(y0f, 1.0f), iy[r]), ny = __fsub_rn(iy[r], y0f);
        const int o = yi * SW + xi;
        const bool all_in = (unsigned)xi < (unsigned)(SW - 1) && (unsigned)yi < (unsigned)(SH - 1);
        if (__all_sync(0xffffffffu, all_in && xin)) {
            // ---- every tap of every lane is inside the image: no predicates
            if (GFLOW) {
                float gix = 0.0f, giy = 0.0f;
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    float dx, dy;
                    if (GFLOW == 2) {
                        dx = __ldg(dt + ((2 * c) * HW + pix));
                        dy = __ldg(dt + ((2 * c + 1) * HW + pix));
                    } else {
                        const float *tn = src + (c * HWs + o), *ts = src + (c * HWs + o + SW);
                        float a, b, cq, d;
                        if (r < kPre) {                    // same predicate as the preload: the values are there
                            a = pre[r < kPre ? r : 0][c][0]; b = pre[r < kPre ? r : 0][c][1];
                            cq = pre[r < kPre ? r : 0][c][2]; d = pre[r < kPre ? r : 0][c][3];
                        } else {
                            a = __ldg(tn); b = __ldg(tn + 1); cq = __ldg(ts); d = __ldg(ts + 1);
                        }
                        // d out / d ix and d out / d iy
                        dx = __fmaf_rn(b - a, sy, (d - cq) * ny);
                        dy = __fmaf_rn(cq - a, ex, (d - b) * wx);
                    }
                    gix = __fmaf_rn(gv[r][c], dx, gix);
                    giy = __fmaf_rn(gv[r][c], dy, giy);
                }
                float rx = -p.sign * gix, ry = -p.sign * giy;
                if (p.add_flow != 0.0f) {                  // out[c] = add_flow*flow[c] + ...  (combine mode 3), c < 2
                    if (c_off == 0) rx += p.add_flow * gv[r][0];
                    if (c_off == 0 && CN > 1) ry += p.add_flow * gv[r][CN > 1 ? 1 : 0];
                    if (c_off == 1) ry += p.add_flow * gv[r][0];
                }
                if (accumulate_gflow) { rx += __ldcg(gf + pix); ry += __ldcg(gf + (HW + pix)); }
                __stcg(gf + pix, rx);
                __stcg(gf + (HW + pix), ry);
            }
            if (tiled) {
                int *wp = win + ((yi - oy) * kTWin + (xi - ox));
                const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    const float gs = gv[r][c] * scale;
                    atomicAdd(wp + c * kTPlane, __float2int_rn(wnw * gs));
                    atomicAdd(wp + c * kTPlane + 1, __float2int_rn(wne * gs));
                    atomicAdd(wp + c * kTPlane + kTWin, __float2int_rn(wsw * gs));
                    atomicAdd(wp + c * kTPlane + kTWin + 1, __float2int_rn(wse * gs));
                }
            } else if (!skip_red) {
                const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    float *t = gsrc + (c * HWs + o);
                    red_add_f32(t, wnw * gv[r][c]);
                    red_add_f32(t + 1, wne * gv[r][c]);
                    red_add_f32(t + SW, wsw * gv[r][c]);
                    red_add_f32(t + SW + 1, wse * gv[r][c]);
                }
            }
        } else {
            // ---- border warps: per-tap predicates
            const bool bx0 = (unsigned)xi < (unsigned)SW, bx1 = (unsigned)(xi + 1) < (unsigned)SW;
            const bool by0 = (unsigned)yi < (unsigned)SH, by1 = (unsigned)(yi + 1) < (unsigned)SH;
            const bool pnw = xin && bx0 && by0, pne = xin && bx1 && by0, psw = xin && bx0 && by1, pse = xin && bx1 && by1;
            if (GFLOW) {
                float gix = 0.0f, giy = 0.0f;
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    float dx, dy;
                    if (GFLOW == 2) {
                        dx = __ldg(dt + ((2 * c) * HW + pix));
                        dy = __ldg(dt + ((2 * c + 1) * HW + pix));
                    } else {
                        const float *tn = src + (c * HWs + o), *ts = src + (c * HWs + o + SW);
                        float a = 0.0f, b = 0.0f, cq = 0.0f, d = 0.0f;
                        if (pnw) a = __ldg(tn);
                        if (pne) b = __ldg(tn + 1);
                        if (psw) cq = __ldg(ts);
                        if (pse) d = __ldg(ts + 1);
                        dx = __fmaf_rn(b - a, sy, (d - cq) * ny);
                        dy = __fmaf_rn(cq - a, ex, (d - b) * wx);
                    }
                    gix = __fmaf_rn(gv[r][c], dx, gix);
                    giy = __fmaf_rn(gv[r][c], dy, giy);
                }
                float rx = -p.sign * gix, ry = -p.sign * giy;
                if (p.add_flow != 0.0f) {
                    if (c_off == 0) rx += p.add_flow * gv[r][0];
                    if (c_off == 0 && CN > 1) ry += p.add_flow * gv[r][CN > 1 ? 1 : 0];
                    if (c_off == 1) ry += p.add_flow * gv[r][0];
                }
                if (xin) {
                    if (accumulate_gflow) { rx += __ldcg(gf + pix); ry += __ldcg(gf + (HW + pix)); }
                    __stcg(gf + pix, rx);
                    __stcg(gf + (HW + pix), ry);
                }
            }
            if (!skip_red) {
                const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
                int *wp = win + ((yi - oy) * kTWin + (xi - ox));       // only dereferenced for in-image taps
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    if (tiled) {
                        const float gs = gv[r][c] * scale;
                        if (pnw) atomicAdd(wp + c * kTPlane, __float2int_rn(wnw * gs));
                        if (pne) atomicAdd(wp + c * kTPlane + 1, __float2int_rn(wne * gs));
                        if (psw) atomicAdd(wp + c * kTPlane + kTWin, __float2int_rn(wsw * gs));
                        if (pse) atomicAdd(wp + c * kTPlane + kTWin + 1, __float2int_rn(wse * gs));
                    } else {
                        float *t = gsrc + (c * HWs + o);
                        if (pnw) red_add_f32(t, wnw * gv[r][c]);
                        if (pne) red_add_f32(t + 1, wne * gv[r][c]);
                        if (psw) red_add_f32(t + SW, wsw * gv[r][c]);
                        if (pse) red_add_f32(t + SW + 1, wse * gv[r][c]);
                    }
                }
            }
        }
    }
    if (!tiled) return;                                     // CTA-uniform
    __syncthreads();

    // ---- phase 3: flush the footprint, 4 columns per thread, empty quads skipped
    {
        const int q = fq, rr0 = fr0;
        if (q < nq) {
            const int gx = ox + 4 * q;
#pragma unroll 1
            for (int c = 0; c < CN; ++c) {
                float *gp = gsrc + (c * HWs + gx);
                const int *wc = win + c * kTPlane + 4 * q;
#pragma unroll 2
                for (int rr = rr0; rr < ex_h; rr += 16) {
                    const int4 v = *reinterpret_cast<const int4 *>(wc + rr * kTWin);
                    if ((v.x | v.y | v.z | v.w) == 0) continue;
                    float *g = gp + (oy + rr) * SW;
                    const float a = (float)v.x * inv_scale, b = (float)v.y * inv_scale;
                    const float cq = (float)v.z * inv_scale, d = (float)v.w * inv_scale;
                    if (vec_ok) {
                        red_add_v4(g, a, b, cq, d);
                    } else {                                 // nonzero entries are inside the image by construction
                        if (v.x) red_add_f32(g, a);
                        if (v.y) red_add_f32(g + 1, b);
                        if (v.z) red_add_f32(g + 2, cq);
                        if (v.w) red_add_f32(g + 3, d);
                    }
                }
            }
        }
    }
}

typedef void (*TileKernel)(const WarpBwdParams, const CoordCtx, const int, const int, const int);

// Deferred pass of the zero-ahead mode: the payload-gradient part of the tiles listed in za[ntr + 1 ...] (footprints that
// reach kZRows or more tile rows down), as plain float REDs per tap.  Launched behind the tile kernel, when everything is
// zero-filled; the list is almost always empty.
__global__ void __launch_bounds__(256) warp_t_bwd_tile_deferred_kernel(const WarpBwdParams p, const CoordCtx cc, const int tiles_x,
                                                                       const int tiles_y)
{
    const int ntr = tiles_y * p.N;
    const int count = p.za[ntr];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = p.W, H = p.H, HW = H * W;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int tile = p.za[ntr + 1 + e];
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int x = tx * kTT + lane;
        if (x >= W) continue;
        const float *fl = p.flow + (int64_t)n * p.flow_bs;
        const float *go = p.gout_bcast ? p.gout : p.gout + (int64_t)n * p.C * HW;
        float *gsrc = p.gsrc + (int64_t)n * p.gsrc_bs;
        for (int r = 0; r < kTRows; ++r) {
            const int y = ty * kTT + warp * kTRows + r;
            if (y >= H) break;
            const int q = y * W + x;
            const float ix = round_trip_unchecked(__fmaf_rn(-cc.sign, __ldg(fl + q), (float)x), cc.dw);
            const float iy = round_trip_unchecked(__fmaf_rn(-cc.sign, __ldg(fl + (HW + q)), (float)y), cc.dh);
            const int xi = __float2int_rd(ix), yi = __float2int_rd(iy);
            const float x0 = floorf(ix), y0f = floorf(iy);
            const float ex = __fsub_rn(__fadd_rn(x0, 1.0f), ix), wx = __fsub_rn(ix, x0);
            const float sy = __fsub_rn(__fadd_rn(y0f, 1.0f), iy), ny = __fsub_rn(iy, y0f);
            const float wnw = __fmul_rn(ex, sy), wne = __fmul_rn(wx, sy), wsw = __fmul_rn(ex, ny), wse = __fmul_rn(wx, ny);
            const bool bx0 = (unsigned)xi < (unsigned)W, bx1 = (unsigned)(xi + 1) < (unsigned)W;
            const bool by0 = (unsigned)yi < (unsigned)H, by1 = (unsigned)(yi + 1) < (unsigned)H;
            const int o = yi * W + xi;
            for (int c = 0; c < p.C; ++c) {
                const float g = p.gout_bcast ? go[0] : __ldg(go + ((int64_t)c * HW + q));
                float *t = gsrc + ((int64_t)c * HW + o);
                if (bx0 && by0) red_add_f32(t, wnw * g);
                if (bx1 && by0) red_add_f32(t + 1, wne * g);
                if (bx0 && by1) red_add_f32(t + W, wsw * g);
                if (bx1 && by1) red_add_f32(t + W + 1, wse * g);
            }
        }
    }
}

template <int CN>
static TileKernel tile_kernel_za(int gflow)
{
    return gflow ? warp_t_bwd_tile_kernel<CN, 1, false, true> : warp_t_bwd_tile_kernel<CN, 0, false, true>;
}

template <int CN>
static TileKernel tile_kernel_for(int gflow, bool win)
{
    if (win) return gflow ? warp_t_bwd_tile_kernel<CN, 1, true> : warp_t_bwd_tile_kernel<CN, 0, true>;
    if (gflow == 2) return warp_t_bwd_tile_kernel<CN, 2, false>;
    return gflow ? warp_t_bwd_tile_kernel<CN, 1, false> : warp_t_bwd_tile_kernel<CN, 0, false>;
}

static TileKernel tile_kernel(int cn, int gflow, bool win)
{
    switch (cn) {
        case 1: return tile_kernel_for<1>(gflow, win);
        case 2: return tile_kernel_for<2>(gflow, win);
        case 3: return tile_kernel_for<3>(gflow, win);
        default: return tile_kernel_for<4>(gflow, win);
    }
}

// Workspace of the zero-ahead mode: one buffer per device, grown on demand, never freed.  Calls on different streams of
// one device would share it, which is why the mode is opt-in and documented as single-stream.
static cudaError_t za_workspace(size_t bytes, int32_t **out)
{
    static int32_t *buf[64] = {nullptr};
    static size_t cap[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (cap[dev] < bytes) {
        cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;      // an allocation is legal inside a relaxed capture
        cudaThreadExchangeStreamCaptureMode(&mode);
        if (buf[dev]) cudaFree(buf[dev]);
        e = cudaMalloc(reinterpret_cast<void **>(&buf[dev]), bytes);
        cudaThreadExchangeStreamCaptureMode(&mode);
        if (e != cudaSuccess) { buf[dev] = nullptr; cap[dev] = 0; return e; }
        cap[dev] = bytes;
    }
    *out = buf[dev];
    return cudaSuccess;
}

// Returns 0 when it handled the call, 1 when the caller must use the register-merge kernel, or a CUDA error.
int warp_t_bwd_tile_launch(const WarpBwdParams &p, const CoordCtx &cc, cudaStream_t st)
{
    if (!(g_bwd_impl < 0 || g_bwd_impl == 7) || !p.gsrc) return 1;
    const Window &wn = p.win;
    const bool win = !(wn.y0 == 0 && wn.x0 == 0 && wn.gh == p.H && wn.gw == p.W && wn.sy0 == 0 && wn.sx0 == 0 &&
                       wn.sh == p.H && wn.sw == p.W);
    if (wn.gw > 65536 || wn.gh > 65536 || wn.gw < 2 || wn.gh < 2) return 1;      // round_trip_unchecked's domain
    if (win && (p.dtaps || p.gout_bcast)) return 1;
    const int64_t HWs = (int64_t)wn.sh * wn.sw;
    const int64_t item_bytes = (int64_t)p.C * HWs * 4;
    // zero-ahead mode (see the kernel): plain frame, one channel group, a payload gradient per item, payload taps
    const int tiles_x = (p.W + kTT - 1) / kTT, tiles_y = (p.H + kTT - 1) / kTT;
    if (g_bwd_za && !win && p.C <= 4 && p.gsrc_bs != 0 && !p.dtaps && tiles_y >= kZRows &&
        (int64_t)tiles_x * tiles_y * p.N < (1 << 30)) {
        const int ntr = tiles_y * p.N;
        const size_t ws_bytes = sizeof(int32_t) * ((size_t)ntr + 1 + (size_t)tiles_x * tiles_y * p.N);
        int32_t *ws = nullptr;
        cudaError_t e = za_workspace(ws_bytes, &ws);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemsetAsync(ws, 0, sizeof(int32_t) * ((size_t)ntr + 1), st);
        // the first kZRows tile rows (all in item 0) have no CTA ahead of them: rows [0, 32 kZRows) of every plane
        if (e == cudaSuccess)
            e = cudaMemset2DAsync(p.gsrc, (size_t)HWs * 4, 0, (size_t)min(kZRows * kTT, p.H) * p.W * 4, (size_t)p.C, st);
        if (e != cudaSuccess) return (int)e;
        WarpBwdParams q = p;
        q.za = ws;
        const int vec_ok = (p.W % 4 == 0) && aligned16(p.gsrc) && (p.gsrc_bs % 4 == 0);
        TileKernel k;
        switch (p.C) {
            case 1: k = tile_kernel_za<1>(p.gflow != nullptr); break;
            case 2: k = tile_kernel_za<2>(p.gflow != nullptr); break;
            case 3: k = tile_kernel_za<3>(p.gflow != nullptr); break;
            default: k = tile_kernel_za<4>(p.gflow != nullptr); break;
        }
        const int smem = p.C * kTPlane * 4;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) {
            k<<<dim3(tiles_x, tiles_y, p.N), 256, smem, st>>>(q, cc, 0, 0, vec_ok);
            ++g_launches;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            warp_t_bwd_tile_deferred_kernel<<<148 * 4, 256, 0, st>>>(q, cc, tiles_x, tiles_y);
            ++g_launches;
            e = cudaGetLastError();
        }
        return e == cudaSuccess ? 0 : (int)e;
    }
    {
        cudaError_t e = cudaMemsetAsync(p.gsrc, 0, (size_t)(p.gsrc_bs == 0 ? item_bytes : item_bytes * p.N), st);
        if (e != cudaSuccess) return (int)e;
    }
    // 16-byte REDs at the flush: every plane row of the stored region must start 16-byte aligned
    const int vec_ok = (wn.sw % 4 == 0) && aligned16(p.gsrc) && (p.gsrc_bs % 4 == 0);
    const dim3 grid((p.W + kTT - 1) / kTT, (p.H + kTT - 1) / kTT, p.N);
    for (int c0 = 0; c0 < p.C; c0 += 4) {          // channel groups of <= 4; gflow accumulates across groups
        const int cn = (p.C - c0 < 4) ? (p.C - c0) : 4;
        TileKernel k = tile_kernel(cn, !p.gflow ? 0 : (p.dtaps ? 2 : 1), win);
        const int smem = cn * kTPlane * 4;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, 256, smem, st>>>(p, cc, c0, c0 > 0 ? 1 : 0, vec_ok);
        ++g_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace ofb
