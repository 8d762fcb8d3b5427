// Shared pieces of the W kernels (warp_t.cu: forward and the register-merge adjoint; warp_t_adj.cu: the ring adjoint).
#pragma once
#include "common.cuh"

namespace ofb {

// Windowed coordinate frame (identity for the plain warp).  The flow / output tensors cover a WINDOW of H x W pixels
// whose top-left corner sits at (y0, x0) of a GLOBAL field of gh x gw pixels; coordinates, the round trip and the
// zero-padding test live in the global frame.  The stored source covers rows [sy0, sy0+sh) x cols [sx0, sx0+sw) of
// the global field (taps outside the stored region read as zero).  Two uses: Flow.apply(padding=..., cut=True)
// without materialising a padded flow (reference flow_class.py:905-918), and row-band spatial sharding of one
// oversized field with halo rows.
struct Window {
    int y0, x0, gh, gw, sy0, sx0, sh, sw;
};

// ------------------------------------------------------------------------------------------------
// direct forward: taps through the read-only path
// ------------------------------------------------------------------------------------------------
struct CoordCtx {     // loop-invariant pieces of the coordinate round trip
    ExactDiv dw, dh;  // (W-1, 1/(W-1)), (H-1, 1/(H-1))
    float sign;
};

// sample position of output pixel (xf, yf) in fp32, exactly as the reference computes it:
// fma(-sign, f, x) == x - sign*f with a single rounding because sign*f is exact (utils.py:548-550)
__device__ __forceinline__ void sample_pos(const CoordCtx &cc, float fx, float fy, float xf, float yf, float &ix, float &iy)
{
    ix = round_trip_fast(__fmaf_rn(-cc.sign, fx, xf), cc.dw);
    iy = round_trip_fast(__fmaf_rn(-cc.sign, fy, yf), cc.dh);
}

struct PixGeom {                 // tap geometry of one output pixel
    float wnw, wne, wsw, wse;    // bilinear weights
    float ex, wx, sy, ny;        // x1-ix, ix-x0, y1-iy, iy-y0 (used by the adjoint)
    int o;                       // yi*W + xi of the NW tap; may be negative (taps left of / above the image), so it
                                 // is SIGN-extended when added to a pointer (only dereferenced under the predicates)
    bool pnw, pne, psw, pse;     // tap inside the image (zero padding otherwise)
};

template <bool WIN>
__device__ __forceinline__ PixGeom pix_geom(float ix, float iy, const Window &wn)
{
    PixGeom g;
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float ex = __fsub_rn(__fadd_rn(x0, 1.0f), ix), wx = __fsub_rn(ix, x0);
    const float sy = __fsub_rn(__fadd_rn(y0, 1.0f), iy), ny = __fsub_rn(iy, y0);
    g.ex = ex; g.wx = wx; g.sy = sy; g.ny = ny;
    g.wnw = __fmul_rn(ex, sy); g.wne = __fmul_rn(wx, sy);
    g.wsw = __fmul_rn(ex, ny); g.wse = __fmul_rn(wx, ny);
    // cvt.rmi.s32.f32 saturates (and maps NaN to 0), so far-away positions become INT_MIN/INT_MAX and fail the
    // unsigned range tests below exactly like the reference's float comparisons x0 >= 0 && x0 <= W-1 do
    const int xi = __float2int_rd(ix), yi = __float2int_rd(iy);
    // WIN = false (compile time): the stored source IS the global field, no second range test, no origin shift
    const int lx = WIN ? xi - wn.sx0 : xi, ly = WIN ? yi - wn.sy0 : yi;   // position inside the stored source region
    const bool bx0 = (unsigned)xi < (unsigned)wn.gw && (!WIN || (unsigned)lx < (unsigned)wn.sw);
    const bool bx1 = (unsigned)(xi + 1) < (unsigned)wn.gw && (!WIN || (unsigned)(lx + 1) < (unsigned)wn.sw);
    const bool by0 = (unsigned)yi < (unsigned)wn.gh && (!WIN || (unsigned)ly < (unsigned)wn.sh);
    const bool by1 = (unsigned)(yi + 1) < (unsigned)wn.gh && (!WIN || (unsigned)(ly + 1) < (unsigned)wn.sh);
    g.pnw = bx0 && by0; g.pne = bx1 && by0; g.psw = bx0 && by1; g.pse = bx1 && by1;
    g.o = (int)((unsigned)ly * (unsigned)wn.sw + (unsigned)lx);   // wraps harmlessly for saturated (far away) positions
    return g;
}

// The four taps of one plane.  `north` points at the NW tap, `south` one image row below; each is ONE 64-bit
// address, the east taps are immediate offsets.  Out-of-image taps are not loaded and read as zero.
// ALL_IN: warp-uniform fast path, every tap of every lane is inside the image -> no predication, no zero-init.
template <bool ALL_IN, typename T>
__device__ __forceinline__ void load_taps(const T *__restrict__ north, const T *__restrict__ south, const PixGeom &g,
                                          T &a, T &b, T &c, T &d)
{
    if (ALL_IN) {
        a = __ldg(north); b = __ldg(north + 1); c = __ldg(south); d = __ldg(south + 1);
    } else {
        a = b = c = d = T(0);
        if (g.pnw) a = __ldg(north);
        if (g.pne) b = __ldg(north + 1);
        if (g.psw) c = __ldg(south);
        if (g.pse) d = __ldg(south + 1);
    }
}

__device__ __forceinline__ float blend4(const PixGeom &g, float a, float b, float c, float d)
{
    float r = __fmul_rn(a, g.wnw);      // ATen's CPU accumulation order (oracle/oracle.c)
    r = __fmaf_rn(b, g.wne, r);
    r = __fmaf_rn(c, g.wsw, r);
    return __fmaf_rn(d, g.wse, r);
}

// 0/1 mask byte -> 0.0f/1.0f without the conversion pipe: 1 * 0x3f800000 is the bit pattern of 1.0f
__device__ __forceinline__ float mask_to_float(uint8_t m) { return __uint_as_float((unsigned)m * 0x3f800000u); }

struct WarpBwdParams {
    const float *flow; int64_t flow_bs; float sign;
    const float *src; int64_t src_bs;
    const float *gout; float add_flow;
    float *gflow; float *gsrc; int64_t gsrc_bs;
    int N, C, H, W;
    Window win;
    const int32_t *sel;   // device-side early-out decision (ofb_early_decide), or nullptr
    const float *dtaps;   // (N, 2C, H, W) d out[c] / d(ix, iy) saved by the forward (ofb_warp_t_fwd_d), or nullptr
    int gout_bcast;       // 1: gout is ONE value broadcast over (N, C, H, W) -- the gradient of out.sum(), never materialised
    int32_t *za;          // zero-ahead workspace of the tile adjoint (warp_t_tile.cu), or nullptr
};

// Predicated fire-and-forget reduction: `@p red.global.add.f32 [a + imm], v` -- no branch, no return value.
// (atomicAdd() on a pointer the optimiser cannot prove global compiles to a generic ATOM with a return path.)
template <int BYTE_OFF>
__device__ __forceinline__ void red_add(const float *a, float v, bool pred)
{
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.s32 q, %2, 0;\n\t"
        "@q red.global.add.f32 [%0+%3], %1;\n\t"
        "}" ::"l"(a), "f"(v), "r"((int)pred), "n"(BYTE_OFF));   // no "memory" clobber: gsrc is write-only here, and
                                                                   // the clobber would pin every load behind the REDs
}


// all-true window built from H, W so that everything folds to the plain-warp code at compile time
__host__ __device__ inline Window identity_window(int H, int W) { return Window{0, 0, H, W, 0, 0, H, W}; }

// warp_t_adj.cu: ring adjoint.  Returns 0 when it handled the call, 1 when the caller must use the register-merge
// kernel (shape / layout outside its domain), or a CUDA error.
int warp_t_bwd_ring_launch(const WarpBwdParams &p, const CoordCtx &cc, cudaStream_t st);

// warp_t_tile.cu: tile-owned accumulation (the default adjoint of the plain warp).  Same return convention.
int warp_t_bwd_tile_launch(const WarpBwdParams &p, const CoordCtx &cc, cudaStream_t st);

}  // namespace ofb
