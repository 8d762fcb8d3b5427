// W adjoint, ring variant: the gradient with respect to the payload accumulates in shared memory, not in L2.
//
// Replaces grid_sampler_2d_backward behind reference utils.py:555 (same arithmetic as warp_t.cu's register-merge
// kernel).  gflow is a gather and is computed exactly as there.  gsrc -- the un-normalised
// splat of w * gout -- used to be ~12 red.global.add instructions per warp-row (377 M L2 RED sectors for 50 M
// destination sectors, profiles/r01_ncu_full_n64_summary.txt): the L1/L2 atomic pipes, not HBM, bounded the kernel.
//
// Here a WARP owns a strip of 32 output columns and walks it top to bottom.  Its contributions go into a PRIVATE
// shared-memory window of kRingWS source columns x kRingRB source rows (a ring over rows, planar per channel) with
// plain LDS / FFMA / STS -- no atomics, because
//   * the window is private to the warp (no other warp touches it), and
//   * within one warp instruction the 32 target addresses are distinct: lanes are ranked among the lanes that hit the
//     same source pixel (strictly increasing x -> all rank 0; otherwise adjacent duplicates or match.any) and each
//     rank is accumulated in its own pass.
// Source rows the walk has left behind are retired in chunks of kRingCR rows: one TMA tensor reduce-add
// (cp.reduce.async.bulk.tensor, SASS UTMAREDG) moves the chunk's box (columns x rows x channels) into gsrc -- the TMA
// unit clips the box against the image, so there is no border logic -- or, when the tensor does not meet the TMA
// constraints, coalesced red.global.add of the non-zero columns.  Taps that fall outside the window (violent flows)
// and pixels whose four taps are not all inside the image take direct global REDs: always correct, only slower.
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"
#include "warp_t.cuh"

namespace ofb {

int g_bwd_za = 0;            // tile adjoint zero-fills the payload gradient itself (warp_t_tile.cu), key 6
int g_bwd_impl = -1;        // tuning knob (ofb_set_option): -1 auto (= 7), 0 register-merge kernel, 1 ring + RED flush, 2 ring + TMA,
                            // 7 tile-owned fixed-point accumulation (warp_t_tile.cu)
int g_bwd_band_rows = 24;   // output rows per warp task
extern int g_splat_det;     // splat_s.cu: deterministic (fixed-point) splat merge
extern int g_fwd_impl;      // warp_t.cu: W forward implementation

constexpr int kRingWS = 64;                       // window columns
constexpr int kRingRB = 16;                       // ring rows
constexpr int kRingCR = 2, kRingLCR = 1;          // rows per retired chunk (and its log2)
constexpr int kRingNCH = kRingRB / kRingCR;
constexpr int kRingWarps = 8;
constexpr int kRingMargin = (kRingWS - 34) / 2;   // the strip's footprint (about 34 columns) sits centred in the window

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int CN, bool GFLOW, bool TMA>
__global__ void __launch_bounds__(kRingWarps * 32, 2)
warp_t_bwd_ring_kernel(const WarpBwdParams p, const CoordCtx cc, const int c_off, const int accumulate_gflow,
                       const int band_rows, const int strips, const int tasks, const int plane_bs,
                       const __grid_constant__ CUtensorMap gmap)
{
    constexpr int WS = kRingWS, RB = kRingRB, CR = kRingCR, LCR = kRingLCR, NCH = kRingNCH;
    constexpr int CHUNK = CN * CR * WS, FLOATS = (NCH + 1) * CHUNK;   // + one scratch chunk
    constexpr int RANGE = TMA ? RB - CR : RB;     // TMA: the chunk being read by the TMA unit is out of range
    constexpr unsigned full = 0xffffffffu;
    extern __shared__ __align__(128) float ring_smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int task = blockIdx.x * kRingWarps + warp;
    if (task >= tasks) return;
    const int band = task / strips, strip = task - band * strips;
    const int x = strip * 32 + lane, y0 = band * band_rows, n = blockIdx.y;
    const int W = p.W, HW = p.H * p.W;
    const bool active = x < W;
    const int xc = min(x, W - 1);
    const int rows = min(band_rows, p.H - y0);

    const float *fl = p.flow + (int64_t)n * p.flow_bs;
    const float *src = GFLOW ? p.src + (int64_t)n * p.src_bs + (int64_t)c_off * HW : nullptr;
    const float *go = p.gout + ((int64_t)n * p.C + c_off) * HW;
    float *gsrc = p.gsrc + (int64_t)n * p.gsrc_bs + (int64_t)c_off * HW;
    float *gf = GFLOW ? p.gflow + (int64_t)n * 2 * HW : nullptr;
    asm volatile("" : "+l"(fl), "+l"(src), "+l"(go), "+l"(gsrc), "+l"(gf));
    const int plane0 = n * plane_bs + c_off;      // first plane of this launch in the (W, H, planes) tensor map

    float *ring = ring_smem + warp * FLOATS;
    for (int i = lane * 4; i < FLOATS; i += 128) *reinterpret_cast<float4 *>(ring + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const float xf = (float)x;
    struct Stage {          // one output row in flight: geometry, payload taps (for gflow), upstream gradient
        float ex, wx, sy, ny;   // x1-ix, ix-x0, y1-iy, iy-y0
        int xi, yi;             // NW tap
        unsigned pm;            // taps inside the image: 1 nw, 2 ne, 4 sw, 8 se
        float gv[CN];
        float a[CN], b[CN], c[CN], d[CN];
    };
    // geometry of row y from its flow, then the loads that depend on it (payload taps) and the upstream gradient
    auto prepare = [&](Stage &s, float fx, float fy, int y) {
        float ix, iy;
        sample_pos(cc, fx, fy, xf, (float)y, ix, iy);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        s.ex = __fsub_rn(__fadd_rn(fx0, 1.0f), ix); s.wx = __fsub_rn(ix, fx0);
        s.sy = __fsub_rn(__fadd_rn(fy0, 1.0f), iy); s.ny = __fsub_rn(iy, fy0);
        s.xi = __float2int_rd(ix); s.yi = __float2int_rd(iy);          // saturating: far-away positions fail the tests
        const bool bx0 = (unsigned)s.xi < (unsigned)W, bx1 = (unsigned)(s.xi + 1) < (unsigned)W;
        const bool by0 = (unsigned)s.yi < (unsigned)p.H, by1 = (unsigned)(s.yi + 1) < (unsigned)p.H;
        s.pm = active ? ((bx0 && by0 ? 1u : 0u) | (bx1 && by0 ? 2u : 0u) | (bx0 && by1 ? 4u : 0u) | (bx1 && by1 ? 8u : 0u)) : 0u;
        const int pix = y * W + xc;
#pragma unroll
        for (int c = 0; c < CN; ++c) s.gv[c] = active ? __ldg(go + (c * HW + pix)) : 0.0f;
        if (GFLOW) {
            const int o = (int)((unsigned)s.yi * (unsigned)W + (unsigned)s.xi);
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const float *north = src + (c * HW + o);
                s.a[c] = s.b[c] = s.c[c] = s.d[c] = 0.0f;
                if (s.pm & 1u) s.a[c] = __ldg(north);
                if (s.pm & 2u) s.b[c] = __ldg(north + 1);
                if (s.pm & 4u) s.c[c] = __ldg(north + W);
                if (s.pm & 8u) s.d[c] = __ldg(north + W + 1);
            }
        }
    };

    // ---- ring state (warp-uniform)
    int ox = 0, ylo = 0, ytop = INT_MIN, pend_slot = 0;
    bool have = false, pending = false;

    auto zero_chunk = [&](int slot) {
        float *cb = ring + slot * CHUNK;
        for (int i = lane * 4; i < CHUNK; i += 128) *reinterpret_cast<float4 *>(cb + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto resolve = [&]() {          // the chunk handed to the TMA unit earlier: wait for its read, then clear it
        if (TMA && pending) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            zero_chunk(pend_slot);
            __syncwarp();
            pending = false;
        }
    };
    auto flush_chunk = [&](int y) {   // y: first source row of the chunk (multiple of CR)
        const int slot = (y >> LCR) & (NCH - 1);
        float *cb = ring + slot * CHUNK;
        if (TMA) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_reduce_add_3d(&gmap, cb, ox, y, plane0);
                bulk_commit();
            }
            pending = true;
            pend_slot = slot;
        } else {
#pragma unroll
            for (int c = 0; c < CN; ++c)
#pragma unroll
                for (int r = 0; r < CR; ++r)
#pragma unroll
                    for (int h = 0; h < WS / 32; ++h) {
                        const int i = (c * CR + r) * WS + h * 32 + lane;
                        const float v = cb[i];
                        // columns / rows outside the image never received anything: v == 0 there
                        red_add<0>(gsrc + (c * HW + (y + r) * W + ox + h * 32 + lane), v, v != 0.0f);
                        cb[i] = 0.0f;
                    }
            __syncwarp();
        }
    };
    auto advance = [&]() {            // retire the chunk at ylo
        resolve();
        if (ylo <= ytop) flush_chunk(ylo);
        ylo += CR;
    };

    float fx1 = 0.0f, fy1 = 0.0f;          // flow of the row after the current one
    Stage SA, SB;
    prepare(SA, __ldg(fl + (y0 * W + xc)), __ldg(fl + (HW + y0 * W + xc)), y0);
    if (rows > 1) { fx1 = __ldg(fl + ((y0 + 1) * W + xc)); fy1 = __ldg(fl + (HW + (y0 + 1) * W + xc)); }

    // one row: request the flow two rows ahead, prepare the next row (its loads fly during this row's work), consume
    auto step = [&](const int r, const Stage &g, Stage &nxt) {
        float fx2 = 0.0f, fy2 = 0.0f;
        if (r + 2 < rows) { fx2 = __ldg(fl + ((y0 + r + 2) * W + xc)); fy2 = __ldg(fl + (HW + (y0 + r + 2) * W + xc)); }
        if (r + 1 < rows) prepare(nxt, fx1, fy1, y0 + r + 1);
        fx1 = fx2; fy1 = fy2;
        const int pix = (y0 + r) * W + xc;
        const float (&gv)[CN] = g.gv;

        if (GFLOW) {
            float gix = 0.0f, giy = 0.0f;
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                // d out / d ix and d out / d iy  (SURVEY.md Appendix A.1); gradients carry a tolerance, so FMAs
                gix = __fmaf_rn(gv[c], __fmaf_rn(g.b[c] - g.a[c], g.sy, (g.d[c] - g.c[c]) * g.ny), gix);
                giy = __fmaf_rn(gv[c], __fmaf_rn(g.c[c] - g.a[c], g.ex, (g.d[c] - g.b[c]) * g.wx), giy);
            }
            float rx = -p.sign * gix, ry = -p.sign * giy;   // d(ix)/d(flow_x) = -sign
            if (p.add_flow != 0.0f) {                       // out[c] = add_flow*flow[c] + ...  (combine mode 3), c < 2
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

        // ---- retire the source rows this row's taps have moved past
        const bool all4 = g.pm == 15u;
        const int m = __reduce_min_sync(full, all4 ? g.yi : INT_MAX);
        if (m != INT_MAX) {
            if (!have) {
                have = true;
                ylo = m & ~(CR - 1);
                // (a box with a negative x origin is an illegal instruction for the TMA reduce: tools/tma_red_test.cu)
                ox = max(0, (__reduce_min_sync(full, all4 ? g.xi : INT_MAX) - kRingMargin) & ~3);
            } else if (m >= ylo + RB) {               // a jump past the whole ring: flush what is dirty, restart
                for (int k = 0; k < NCH; ++k) advance();
                resolve();
                ylo = m & ~(CR - 1);
            } else {
                while (m >= ylo + CR) advance();
            }
        }
        const int lx = g.xi - ox;
        const bool rg = all4 && (unsigned)lx <= (unsigned)(WS - 2) && (unsigned)(g.yi - ylo) <= (unsigned)(RANGE - 2);
        const bool fb = (g.pm != 0u) && !rg;
        ytop = max(ytop, __reduce_max_sync(full, rg ? g.yi + 1 : INT_MIN));

        const float wnw = g.ex * g.sy, wne = g.wx * g.sy, wsw = g.ex * g.ny, wse = g.wx * g.ny;
        const int o = (int)((unsigned)g.yi * (unsigned)W + (unsigned)g.xi);

        // ---- rank lanes that hit the same source pixel; each rank accumulates in its own pass
        int rank = 0, maxrank = 0;
        {
            const int px = __shfl_up_sync(full, g.xi, 1);
            const bool strict = (lane == 0) || (g.xi > px);
            if (!__all_sync(full, strict)) {
                const int px2 = __shfl_up_sync(full, g.xi, 2);
                const int po = __shfl_up_sync(full, o, 1);
                const bool tame = ((lane == 0) || (g.xi >= px)) && ((lane < 2) || (g.xi > px2));
                if (__all_sync(full, tame)) {         // non-decreasing x, runs of at most two: duplicates are adjacent
                    rank = (lane > 0 && o == po) ? 1 : 0;
                    maxrank = __any_sync(full, rank && rg) ? 1 : 0;
                } else {
                    const unsigned peers = __match_any_sync(full, rg ? o : -1 - lane);
                    rank = __popc(peers & ((1u << lane) - 1u));
                    maxrank = __reduce_max_sync(full, rg ? rank : 0);
                }
            }
        }
        {
            // Branch-free passes: lanes that do not take part (other rank, outside the window, border pixel) run the
            // same LDS/FFMA/STS against a private scratch chunk, so the warp never diverges here.
            const int yn = g.yi & (RB - 1), ys = (g.yi + 1) & (RB - 1);
            float *wn_ = ring + (((yn >> LCR) * (CN * CR) + (yn & (CR - 1))) * WS + lx);
            float *ws_ = ring + (((ys >> LCR) * (CN * CR) + (ys & (CR - 1))) * WS + lx);
            float *dummy = ring + (NCH * CHUNK + lane);   // scratch: overlapping garbage is fine
            for (int pass = 0; pass <= maxrank; ++pass) {
                const bool on = rg && rank == pass;
                float *pn = on ? wn_ : dummy, *ps = on ? ws_ : dummy;
#pragma unroll
                for (int c = 0; c < CN; ++c) pn[c * CR * WS] = __fmaf_rn(wnw, gv[c], pn[c * CR * WS]);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CN; ++c) pn[c * CR * WS + 1] = __fmaf_rn(wne, gv[c], pn[c * CR * WS + 1]);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CN; ++c) ps[c * CR * WS] = __fmaf_rn(wsw, gv[c], ps[c * CR * WS]);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CN; ++c) ps[c * CR * WS + 1] = __fmaf_rn(wse, gv[c], ps[c * CR * WS + 1]);
                __syncwarp();
            }
        }
        // ---- taps outside the window, and pixels at the image border: straight to global memory
        if (__any_sync(full, fb)) {
            const float *t = gsrc + o;
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                red_add<0>(t + c * HW, wnw * gv[c], fb && (g.pm & 1u));
                red_add<4>(t + c * HW, wne * gv[c], fb && (g.pm & 2u));
                red_add<0>(t + (c * HW + W), wsw * gv[c], fb && (g.pm & 4u));
                red_add<4>(t + (c * HW + W), wse * gv[c], fb && (g.pm & 8u));
            }
        }
    };
#pragma unroll 1
    for (int r = 0; r < rows; r += 2) {
        step(r, SA, SB);
        if (r + 1 < rows) step(r + 1, SB, SA);
    }
    // ---- end of the strip: everything still in the ring goes out
    if (have) {
        for (int k = 0; k < NCH && ylo <= ytop; ++k) advance();
        if (TMA && pending) {
            if (lane == 0) bulk_wait_read0();     // the TMA unit must have read the chunk before the CTA's smem goes away
            __syncwarp();
        }
    }
}


typedef void (*RingKernel)(const WarpBwdParams, const CoordCtx, const int, const int, const int, const int, const int,
                           const int, const CUtensorMap);

template <int CN>
static RingKernel ring_kernel_for(bool gflow, bool tma)
{
    if (gflow) return tma ? warp_t_bwd_ring_kernel<CN, true, true> : warp_t_bwd_ring_kernel<CN, true, false>;
    return tma ? warp_t_bwd_ring_kernel<CN, false, true> : warp_t_bwd_ring_kernel<CN, false, false>;
}

static RingKernel ring_kernel(int cn, bool gflow, bool tma)
{
    switch (cn) {
        case 1: return ring_kernel_for<1>(gflow, tma);
        case 2: return ring_kernel_for<2>(gflow, tma);
        case 3: return ring_kernel_for<3>(gflow, tma);
        default: return ring_kernel_for<4>(gflow, tma);
    }
}

int warp_t_bwd_ring_launch(const WarpBwdParams &p, const CoordCtx &cc, cudaStream_t st)
{
    int impl = g_bwd_impl;
    if (impl < 0) impl = 0;
    if ((impl != 1 && impl != 2) || !p.gsrc || p.W < 64 || p.H < 2) return 1;
    const int64_t HW = (int64_t)p.H * p.W;
    const int64_t item_bytes = (int64_t)p.C * HW * 4;
    {
        cudaError_t e = cudaMemsetAsync(p.gsrc, 0, (size_t)(p.gsrc_bs == 0 ? item_bytes : item_bytes * p.N), st);
        if (e != cudaSuccess) return (int)e;
    }
    const int band_rows = g_bwd_band_rows < 4 ? 4 : g_bwd_band_rows;
    const int strips = (p.W + 31) / 32, bands = (p.H + band_rows - 1) / band_rows;
    const int tasks = strips * bands;
    const dim3 grid((tasks + kRingWarps - 1) / kRingWarps, p.N);
    // TMA flush: gsrc as a (W, H, planes) fp32 tensor; needs dense planes, a 16 B aligned base and W % 4 == 0
    bool tma = impl == 2 && (p.gsrc_bs == 0 || p.gsrc_bs == (int64_t)p.C * HW);
    const int plane_bs = p.gsrc_bs == 0 ? 0 : p.C;
    const uint64_t planes = (uint64_t)(p.gsrc_bs == 0 ? 1 : p.N) * p.C;
    for (int c0 = 0; c0 < p.C; c0 += 4) {          // channel groups of <= 4; gflow accumulates across groups
        const int cn = (p.C - c0 < 4) ? (p.C - c0) : 4;
        alignas(64) CUtensorMap gmap;
        bool use_tma = tma && make_map_3d(&gmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.gsrc, p.W, p.H, planes, kRingWS,
                                          kRingCR, cn);
        if (!use_tma) memset(&gmap, 0, sizeof(gmap));
        RingKernel k = ring_kernel(cn, p.gflow != nullptr, use_tma);
        const int smem = kRingWarps * (kRingNCH + 1) * cn * kRingCR * kRingWS * 4;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, kRingWarps * 32, smem, st>>>(p, cc, c0, c0 > 0 ? 1 : 0, band_rows, strips, tasks, plane_bs, gmap);
        ++g_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace ofb

// Tuning knobs (profiling sweeps; the defaults are the measured best).  key 1: W adjoint implementation
// (-1 auto, 0 register-merge, 1 ring + RED flush, 2 ring + TMA reduce flush); key 2: output rows per warp task;
// key 4: deterministic splat merge (64-bit fixed point) on/off; key 5: W forward (0 lean kernel, 1 direct kernel);
// key 6: zero-ahead mode of the tile adjoint on/off.
extern "C" int ofb_set_option(int key, int value)
{
    if (key == 1) { ofb::g_bwd_impl = value; return 0; }
    if (key == 2) { ofb::g_bwd_band_rows = value; return 0; }
    if (key == 4) { ofb::g_splat_det = value ? 1 : 0; return 0; }
    if (key == 5) { ofb::g_fwd_impl = value; return 0; }
    if (key == 6) { ofb::g_bwd_za = value ? 1 : 0; return 0; }
    return OFB_EINVAL;
}
