/*
 * oflib_b200.h -- C ABI of the B200-native dense flow-field hot path.
 *
 * The reference (oflibpytorch 2.1.1) has no FFI layer: its hot path is Python that calls
 * torch.nn.functional.grid_sample (utils.py:555) and Tensor.scatter_add_ (utils.py:1133,1139).
 * Each entry point below replaces one of those call sites together with the elementwise glue
 * around it; the reference-side binding (ctypes) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device; tensors are dense, row-major
 *     "planar" NCHW / NHW as in the reference ((N,2,H,W) flows, (N,C,H,W) payloads, (N,H,W) masks);
 *   - masks are 1 byte per pixel, 0 or 1 (torch.bool storage);
 *   - a "*_bs" argument is the batch stride of that tensor in ELEMENTS; 0 broadcasts one item over
 *     the batch (the reference's 1<->N batch broadcasting, utils.py:527-537);
 *   - optional pointers may be NULL where the comment says so;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - the return value is 0 on success, a positive cudaError_t, or a negative OFB_E* code;
 *     ofb_error_string() renders either.  No call allocates, synchronises or throws.
 *   - fp32 arithmetic, IEEE division, no fast-math; the operation order is that of oracle/oracle.c
 *     (SURVEY.md Appendix A), which is what makes the boolean outputs bit-exact.
 */
#ifndef OFLIB_B200_H
#define OFLIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFB_ABI_VERSION 1

#define OFB_EINVAL (-1)   /* bad argument (NULL where required, non-positive size, C too large) */
#define OFB_ESHAPE (-2)   /* H < 2 or W < 2: the reference divides by (W-1), (H-1) (utils.py:463-464) */

/* bits written by ofb_flow_stats, one int32 per batch item */
#define OFB_STAT_NONFINITE   1   /* some component is NaN/Inf            (utils.py:98)            */
#define OFB_STAT_NONZERO     2   /* some component != 0                  (utils.py:938, thresholded=False) */
#define OFB_STAT_NONZERO_T   4   /* some |component| >= 1e-3             (utils.py:642,938)        */
#define OFB_STAT_NONZERO_M   8   /* as NONZERO,   only where mask is set (flow_class.py:1243)      */
#define OFB_STAT_NONZERO_TM 16   /* as NONZERO_T, only where mask is set                           */

/* validity rules for ofb_splat_s_fwd */
#define OFB_VALID_NONE 0
#define OFB_VALID_GT   1   /* mask channel > 0.99999   (flow_class.py:922)        */
#define OFB_VALID_EQ1  2   /* mask channel == 1        (flow_class.py:1117,1170)  */

int ofb_abi_version(void);
const char *ofb_error_string(int status);

/* Number of kernels this library has launched since load / since the last reset (host counter,
 * used by bench.py for its "gpu_launches" claim). */
int64_t ofb_launch_count(void);
void ofb_launch_count_reset(void);

/* Tuning knobs for profiling sweeps (the defaults are the measured best; not needed in production).
 *   key 1  W adjoint implementation: -1 auto (= 7), 0 register-merge kernel, 1 ring + RED flush, 2 ring + TMA reduce
 *          flush, 7 tile-owned fixed-point accumulation
 *   key 2  W adjoint ring: output rows per warp task
 *   key 4  S forward: 1 = deterministic merge (64-bit fixed-point accumulation: float outputs bit-identical from run
 *          to run, about 3x slower), 0 = float atomics (default).  ofb_splat_s_workspace follows the mode.
 *   key 5  W forward: 0 = lean kernel (default), 1 = direct kernel (the general path: windows, > 4 channels)
 *   key 6  W adjoint (tile kernel): 1 = the kernel zero-fills the payload gradient itself, tile rows ahead of its
 *          REDs, instead of a memset pass (less DRAM traffic, measured slower; uses one workspace per device, so calls
 *          must not overlap on two streams of one device), 0 = memset (default)
 * Returns 0, or OFB_EINVAL for an unknown key.  Process-wide, not thread-safe. */
int ofb_set_option(int key, int value);

/* Self-test: counts the float bit patterns b in [lo_bits, hi_bits) (both signs) for which the kernels' coordinate
 * round trip (constant-reciprocal division, d = W-1 or H-1) differs from the one evaluated with IEEE division as
 * the reference does (utils.py:462-465 + ATen GridSampler.h:31).  *mismatches (device, zeroed by the caller) must
 * stay 0: the validity masks depend on it. */
int ofb_selftest_div(float d, uint32_t lo_bits, uint32_t hi_bits, unsigned long long *mismatches, void *stream);

/* One pass over a flow field: replaces get_valid_vecs' isfinite().all() (utils.py:98), is_zero_flow
 * (utils.py:919-938) and Flow.is_zero (flow_class.py:1226-1244).  flags[N] must be zero on entry and
 * receives the OR of the OFB_STAT_* bits per batch item.  mask (N,H,W) may be NULL (then the *_M
 * bits equal the unmasked ones). */
int ofb_flow_stats(const float *flow, const uint8_t *mask, int32_t *flags,
                   int N, int H, int W, void *stream);

/* ofb_flow_stats with a batch summary: flags has N + 1 entries (zero on entry); flags[N] receives the OR over all
 * items -- the word the device-side early-outs of ofb_warp_t_fwd_ex / ofb_warp_t_bwd_ex test, i.e. what the
 * reference's all(is_zero_flow(..)) over the batch looks at (utils.py:497, flow_class.py:1046,1729,1738). */
int ofb_flow_stats_batch(const float *flow, const uint8_t *mask, int32_t *flags,
                         int N, int H, int W, void *stream);

/* W -- backward warp of a payload by a target-referenced flow, with fused validity.
 * Replaces apply_flow(ref='t') (utils.py:541-555: grid - flow, normalise_coords, grid_sample(bilinear,
 * zeros, align_corners=True)) and, when `valid` is given, the mask tracking of Flow.apply
 * (flow_class.py:896-898,922-934): the payload mask travels as a 0/1 channel through the same weights
 * and  valid = (warped mask channel > valid_thresh) & flow_mask.
 *   sample position   p - flow_sign * flow(p)            (flow_sign = -1 restates apply_flow(-vecs, ..))
 *   out[c]            bilinear(src[c]) (+ add_flow * flow[c] when C == 2: combine mode 3,
 *                                       flow_class.py:1804,1808)
 *   src      (N|1, C, H, W)  may be NULL iff C == 0 (validity only: valid_target/valid_source 't')
 *   src_mask (N|1, H, W)     NULL = all true           flow_mask (N|1, H, W)  NULL = all true
 *   out      (N, C, H, W)    may be NULL iff C == 0    valid (N, H, W)        NULL = not wanted */
int ofb_warp_t_fwd(const float *flow, int64_t flow_bs, float flow_sign,
                   const float *src, int64_t src_bs,
                   const uint8_t *src_mask, int64_t src_mask_bs,
                   const uint8_t *flow_mask, int64_t flow_mask_bs,
                   float add_flow, float valid_thresh,
                   float *out, uint8_t *valid,
                   int N, int C, int H, int W, void *stream);

/* Adjoint of ofb_warp_t_fwd with respect to flow and src (replaces grid_sampler_2d_backward and the
 * backward of normalise_coords).  gflow (N,2,H,W) and gsrc may each be NULL.  gsrc has batch stride
 * gsrc_bs (0 accumulates a broadcast payload's gradient over the batch) and is zero-filled by the
 * call before accumulation.  With add_flow != 0 (C == 2) gflow also receives add_flow * gout. */
int ofb_warp_t_bwd(const float *flow, int64_t flow_bs, float flow_sign,
                   const float *src, int64_t src_bs,
                   const float *gout, float add_flow,
                   float *gflow, float *gsrc, int64_t gsrc_bs,
                   int N, int C, int H, int W, void *stream);

/* The reference's zero-flow early-outs decided ON THE DEVICE (no host synchronisation).
 * ofb_early_decide evaluates up to three rules against the batch summary words of two flag arrays and leaves the
 * decision in *sel (device int32): 0 = none, or the action of the first rule that fired.
 *   flow_flags, payload_flags   int32[N + 1] as written by ofb_flow_stats_batch (or the out_flags of an earlier
 *                               ofb_warp_t_fwd_ex) for the warping flow and for the payload when it is itself a flow;
 *                               may be NULL
 *   rules                       one rule per byte, tested in order; a rule fires when its OFB_STAT bit is CLEAR in
 *                               flags[N]:   bit 7     rule present
 *                                           bit 5     0: test flow_flags, 1: test payload_flags
 *                                           bits 2-4  log2 of the OFB_STAT_* bit
 *                                           bits 0-1  action
 *   actions   1  out = add_flow*flow + src,  valid = src_mask & flow_mask   (apply_flow's early-out, utils.py:497-498)
 *             2  out = src,                  valid = src_mask               (combine_with hands back one of its
 *             3  out = add_flow*flow,        valid = flow_mask               inputs, flow_class.py:1729-1744)
 * ofb_warp_t_fwd_ex / ofb_warp_t_bwd_ex are ofb_warp_t_fwd / ofb_warp_t_bwd honouring *sel (NULL = no early-out);
 * the adjoint must be given the same sel.  out_flags (NULL ok; needs C == 2 and valid, i.e. a flow-valued result)
 * receives the OFB_STAT_* flags of (out, valid): int32[ofb_warp_t_flags_ints(N, H, W)], the first N + 1 entries are
 * the flags and must be zero on entry, the rest is workspace. */
int ofb_early_decide(const int32_t *flow_flags, const int32_t *payload_flags, int rules, int N, int32_t *sel,
                     void *stream);
int64_t ofb_warp_t_flags_ints(int N, int H, int W);
int ofb_warp_t_fwd_ex(const float *flow, int64_t flow_bs, float flow_sign,
                      const float *src, int64_t src_bs,
                      const uint8_t *src_mask, int64_t src_mask_bs,
                      const uint8_t *flow_mask, int64_t flow_mask_bs,
                      float add_flow, float valid_thresh,
                      float *out, uint8_t *valid,
                      int N, int C, int H, int W,
                      const int32_t *sel, int32_t *out_flags, void *stream);
int ofb_warp_t_bwd_ex(const float *flow, int64_t flow_bs, float flow_sign,
                      const float *src, int64_t src_bs,
                      const float *gout, float add_flow,
                      float *gflow, float *gsrc, int64_t gsrc_bs,
                      int N, int C, int H, int W,
                      const int32_t *sel, void *stream);

/* W with saved derivatives.  ofb_warp_t_fwd_d is ofb_warp_t_fwd_ex that additionally writes
 *   dtaps (N, 2C, H, W)   plane 2c = d out[c] / d ix, plane 2c+1 = d out[c] / d iy (the bilinear derivative of the
 *                         payload at the sample position, before the chain through the flow; the add_flow term is not
 *                         part of it)
 * and ofb_warp_t_bwd_d is ofb_warp_t_bwd_ex that, given the same dtaps, forms the flow gradient as
 * sum_c gout[c] * dtaps[2c .. 2c+1] instead of gathering the 4C payload taps again (src may then be NULL).  This is
 * the arithmetic of ATen's grid_sampler_2d_backward (behind utils.py:555) split at the point where the taps have been
 * combined: the values are bit-identical to ofb_warp_t_bwd's.  dtaps = NULL makes both behave exactly like the _ex
 * functions.  Needs 1 <= C <= 4 and H, W <= 65536 (OFB_EINVAL otherwise).
 * gout_broadcast = 1: gout points at ONE float that stands for the whole (N, C, H, W) upstream gradient -- what
 * autograd hands down for out.sum() as a stride-0 view -- and is never materialised (needs the
 * default adjoint kernel; OFB_EINVAL otherwise, the caller then passes a dense gout). */
int ofb_warp_t_fwd_d(const float *flow, int64_t flow_bs, float flow_sign,
                     const float *src, int64_t src_bs,
                     const uint8_t *src_mask, int64_t src_mask_bs,
                     const uint8_t *flow_mask, int64_t flow_mask_bs,
                     float add_flow, float valid_thresh,
                     float *out, uint8_t *valid,
                     int N, int C, int H, int W,
                     const int32_t *sel, int32_t *out_flags, float *dtaps, void *stream);
int ofb_warp_t_bwd_d(const float *flow, int64_t flow_bs, float flow_sign,
                     const float *src, int64_t src_bs,
                     const float *gout, float add_flow,
                     float *gflow, float *gsrc, int64_t gsrc_bs,
                     int N, int C, int H, int W,
                     const int32_t *sel, const float *dtaps, int gout_broadcast, void *stream);

/* W with an INSET flow: flow (N|1, 2, fh, fw) and flow_mask (N|1, fh, fw) cover the window at (fwin[0], fwin[1]) =
 * (y0, x0) of size fwin[2] x fwin[3] = fh x fw inside the H x W frame of src / out / valid; outside the window the flow
 * is zero and valid is false.  Replaces Flow.apply(padding=..., cut=False) for 't' flows (reference
 * flow_class.py:905-934: zero-pad the flow, warp the whole frame, keep the valid area inside the window only) without
 * materialising the padded flow and mask; per pixel it is the arithmetic of ofb_warp_t_fwd, so the results are
 * bit-identical to warping with the padded flow.  `fwin` is a HOST pointer to 4 ints.  The adjoint returns gflow of the
 * window's size (the padded vectors are constants) and gsrc of the frame's size (zero-filled by the call). */
int ofb_warp_t_inset_fwd(const float *flow, int64_t flow_bs,
                         const float *src, int64_t src_bs,
                         const uint8_t *src_mask, int64_t src_mask_bs,
                         const uint8_t *flow_mask, int64_t flow_mask_bs,
                         float valid_thresh, float *out, uint8_t *valid,
                         int N, int C, int H, int W, const int *fwin, void *stream);
int ofb_warp_t_inset_bwd(const float *flow, int64_t flow_bs,
                         const float *src, int64_t src_bs,
                         const float *gout, float *gflow, float *gsrc, int64_t gsrc_bs,
                         int N, int C, int H, int W, const int *fwin, void *stream);

/* Windowed W: the flow/out/valid/flow_mask tensors cover an H x W WINDOW located at (window[0], window[1]) = (y0, x0)
 * of a global field of window[2] x window[3] pixels; coordinates, the round trip and zero padding live in the global
 * frame.  src / src_mask / gsrc store rows [window[4], window[4]+window[6]) x cols [window[5], window[5]+window[7])
 * of the global field (src_bs, src_mask_bs, gsrc_bs refer to that stored size); taps outside the stored region read as
 * zero.  Uses: Flow.apply(padding=[t,b,l,r], cut=True) for 't' flows without materialising a padded flow (reference
 * flow_class.py:905-918), and row-band sharding of one oversized field with halo rows.  `window` is a
 * HOST pointer to 8 ints.  Everything else as ofb_warp_t_fwd / ofb_warp_t_bwd. */
int ofb_warp_t_win_fwd(const float *flow, int64_t flow_bs, float flow_sign,
                       const float *src, int64_t src_bs,
                       const uint8_t *src_mask, int64_t src_mask_bs,
                       const uint8_t *flow_mask, int64_t flow_mask_bs,
                       float add_flow, float valid_thresh,
                       float *out, uint8_t *valid,
                       int N, int C, int H, int W, const int *window, void *stream);
int ofb_warp_t_win_bwd(const float *flow, int64_t flow_bs, float flow_sign,
                       const float *src, int64_t src_bs,
                       const float *gout, float add_flow,
                       float *gflow, float *gsrc, int64_t gsrc_bs,
                       int N, int C, int H, int W, const int *window, void *stream);

/* Workspace (in floats) ofb_splat_s_fwd needs for this problem; has_data_mask != 0 iff data_mask will be given. */
int64_t ofb_splat_s_workspace(int N, int C, int H, int W, int has_data_mask);

/* S -- normalised bilinear forward splat.  Replaces get_flow_endpoints (utils.py:1045-1058),
 * grid_from_unstructured_data (utils.py:1061-1154) and apply_s_flow (utils.py:1157-1205), and, with
 * `data_mask`/`valid`, the mask tracking of Flow.apply for 's' flows (flow_class.py:880-898,922).
 *   coordinates   relative != 0:  x = sign*cx + j, y = sign*cy + i  (cx, cy = the two planes of a flow)
 *                 relative == 0:  x = cx, y = cy                     (grid_from_unstructured_data)
 *   cx, cy    (N|1, H, W) planes, batch stride coord_bs
 *   data      (N|1, C, H, W), C may be 0 (then data/out may be NULL)
 *   data_mask (N|1, H, W) the 0/1 "mask channel" splatted along with the payload; NULL = all ones
 *   flow_mask (N|1, H, W) sources to ignore (apply_s_flow's `mask`); NULL = all true
 *   occlude_zero  != 0: sources with |f0|<1e-3 and |f1|<1e-3 are left out and filled back in where
 *                 nothing landed (utils.py:1190-1203); requires relative != 0
 *   acc       workspace of acc_floats floats (ask ofb_splat_s_workspace), zero-filled and reused by the call
 *   out (N,C,H,W); density (N,H,W) NULL ok; warped_mask (N,H,W) = density > 0, NULL ok;
 *   unocc (N,H,W) = pixels filled back in by the occlusion rule (needed by the adjoint), NULL ok;
 *   valid (N,H,W) by rule valid_mode, NULL ok. */
int ofb_splat_s_fwd(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                    const float *data, int64_t data_bs,
                    const uint8_t *data_mask, int64_t data_mask_bs,
                    const uint8_t *flow_mask, int64_t flow_mask_bs,
                    int occlude_zero, float *acc, int64_t acc_floats,
                    float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                    uint8_t *valid, int valid_mode,
                    int N, int C, int H, int W, void *stream);

/* ofb_splat_s_fwd for a flow that is a WINDOW of a larger frame -- the padded 's' apply without the padding
 * (Flow.apply(padding=..., cut=True), flow_class.py:905-918, where the reference pads the flow to the target's size,
 * splats the whole frame and slices the window back out).  All tensors are the H x W window; window (host, 4 ints) =
 * {y0, x0, frame_h, frame_w}.  End points are computed in the frame (y + y0 + sign*cy, x + x0 + sign*cx: the same
 * floats as the reference's), tested against the frame, and dropped when they land outside the window.  Relative
 * coordinates only; payloads of up to 6 channels (OFB_EINVAL otherwise: use the padded path). */
int ofb_splat_s_win_fwd(const float *cx, const float *cy, int64_t coord_bs, float sign,
                        const float *data, int64_t data_bs,
                        const uint8_t *data_mask, int64_t data_mask_bs,
                        const uint8_t *flow_mask, int64_t flow_mask_bs,
                        int occlude_zero, float *acc, int64_t acc_floats,
                        float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                        uint8_t *valid, int valid_mode,
                        int N, int C, int H, int W, const int *window, void *stream);

/* Adjoint of ofb_splat_s_fwd with respect to the coordinates (hence the flow) and the data.
 *   out, density, unocc   the forward results (out BEFORE any dtype rounding; unocc NULL = none)
 *   gout (N,C,H,W); gdensity (N,H,W) NULL = zero
 *   gcx, gcy       (N,H,W) planes with batch stride gcoord_bs (the two planes of gflow), NULL ok (both)
 *   gdata          (N,C,H,W), NULL ok
 *   work           scratch of work_floats floats (ofb_splat_s_bwd_workspace), NULL ok */
int ofb_splat_s_bwd(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                    const float *data, int64_t data_bs,
                    const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                    const float *out, const float *density, const uint8_t *unocc,
                    const float *gout, const float *gdensity,
                    float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                    float *work, int64_t work_floats,
                    int N, int C, int H, int W, void *stream);

/* Adjoint of ofb_splat_s_win_fwd; `work` (ofb_splat_s_bwd_workspace) is required, C <= 7. */
int ofb_splat_s_win_bwd(const float *cx, const float *cy, int64_t coord_bs, float sign,
                        const float *data, int64_t data_bs,
                        const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                        const float *out, const float *density, const uint8_t *unocc,
                        const float *gout, const float *gdensity,
                        float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                        float *work, int64_t work_floats,
                        int N, int C, int H, int W, const int *window, void *stream);

/* Workspace (in floats) ofb_splat_s_bwd wants for its vector-gather path; 0 = none.  `work` may be NULL (then the
 * generic scalar-gather kernel runs). */
int64_t ofb_splat_s_bwd_workspace(int N, int C, int H, int W);

/* Row-band splat for ONE field shared by several GPUs: the H x W tensors are rows
 * [band[0], band[0]+H) of a global field of band[1] rows; the accumulator holds global rows [band[2], band[2]+band[3])
 * as (N, band[3], W, KP) floats, KP = ofb_splat_s_kp(C, data_mask != NULL) (0: payload too wide).  accumulate
 * zero-fills and fills the accumulator (targets outside its rows are dropped: size the halo for the flow); the caller
 * adds the neighbours' halo rows (P2P) and then calls normalise for its own rows.  Arguments as ofb_splat_s_fwd;
 * `band` is a HOST pointer to 4 ints. */
int ofb_splat_s_kp(int C, int has_data_mask);
int ofb_splat_s_band_accumulate(const float *cx, const float *cy, int64_t coord_bs, int relative, float sign,
                                const float *data, int64_t data_bs,
                                const uint8_t *data_mask, int64_t data_mask_bs,
                                const uint8_t *flow_mask, int64_t flow_mask_bs,
                                int occlude_zero, float *acc, int N, int C, int H, int W, const int *band, void *stream);
int ofb_splat_s_band_normalise(const float *cx, const float *cy, int64_t coord_bs, int relative,
                               const float *data, int64_t data_bs,
                               const uint8_t *data_mask, int64_t data_mask_bs,
                               const uint8_t *flow_mask, int64_t flow_mask_bs,
                               int occlude_zero, const float *acc,
                               float *out, float *density, uint8_t *warped_mask, uint8_t *unocc,
                               uint8_t *valid, int valid_mode,
                               int N, int C, int H, int W, const int *band, void *stream);

/* Row-band adjoint of the splat (one oversized field split over GPUs, see ofb_splat_s_band_accumulate).
 *   1. ofb_splat_s_band_bwd_prepass on the rank's own target rows: G (N,H,W,KG), KG = ofb_splat_s_kg(C), holds per
 *      target pixel {gout_c / max(den, 1e-3), d loss / d den}; out/density/unocc are the band's forward results.
 *   2. the caller copies the neighbours' halo rows of G in (sharding.exchange_row_halos);
 *   3. ofb_splat_s_band_bwd_gather for the rank's own sources against the extended G:
 *      band (host, 4 ints) = {src_y0, global_h, acc_y0, acc_h}, G = (N, acc_h, W, KG) = frame rows
 *      [acc_y0, acc_y0 + acc_h); unocc / gout are the rank's own rows (gout only read where unocc is set). */
int ofb_splat_s_kg(int C);
int ofb_splat_s_band_bwd_prepass(const float *out, const float *density, const uint8_t *unocc,
                                 const float *gout, const float *gdensity, float *G,
                                 int N, int C, int H, int W, void *stream);
int ofb_splat_s_band_bwd_gather(const float *cx, const float *cy, int64_t coord_bs, float sign,
                                const float *data, int64_t data_bs,
                                const uint8_t *flow_mask, int64_t flow_mask_bs, int occlude_zero,
                                const float *G, const uint8_t *unocc, const float *gout,
                                float *gcx, float *gcy, int64_t gcoord_bs, float *gdata,
                                int N, int C, int H, int W, const int *band, void *stream);

/* Point sampling of a flow field -- the sparse half of track_pts (utils.py:1003-1019): out[n,m,:] = (flow_y, flow_x)
 * bilinearly interpolated at pts[n,m,:] = (y, x) with grid_sample(align_corners=True, zeros) semantics.
 *   flow (N|1, 2, H, W); pts, out (N, M, 2).  The adjoint yields gflow (N|1,2,H,W; zero-filled by the call; stride 0
 *   accumulates over the batch) and gpts (N, M, 2); either may be NULL. */
int ofb_sample_pts_fwd(const float *flow, int64_t flow_bs, const float *pts, float *out,
                       int N, int M, int H, int W, void *stream);
int ofb_sample_pts_bwd(const float *flow, int64_t flow_bs, const float *pts, const float *gout,
                       float *gflow, int64_t gflow_bs, float *gpts, int N, int M, int H, int W, void *stream);

/* ---- callers and data formats either side of the path ----------------------------- */

/* Masked extent of the flow end points: extent (N,4) floats = [min_y, max_y, min_x, max_x] of
 * pos = grid -/+ thresholded(flow) over the pixels where mask is set (mask NULL = all), i.e. the reductions of
 * Flow.get_padding (flow_class.py:1195-1216) for a whole batch in one pass and one host read.  An item without a
 * single masked pixel reports [+inf, -inf, +inf, -inf].  is_source_ref != 0 for 's' flows. */
int ofb_flow_extent(const float *flow, const uint8_t *mask, int is_source_ref, float *extent,
                    int N, int H, int W, void *stream);

/* Sintel .flo payload, (N,H,W,2) interleaved little-endian floats as they sit in the file after the 12-byte
 * header, to (N,2,H,W) planes (load_sintel + to_tensor, utils.py:844-855). */
int ofb_flo_to_planar(const float *interleaved, float *planar, int N, int H, int W, void *stream);

/* KITTI flow png payload, (N,H,W,3) uint16 in the channel order png decoders deliver it with cv2 ([valid, v, u]),
 * to flow planes (u16 - 2^15) / 64 and a 0/1 valid mask (load_kitti, utils.py:821-829); valid may be NULL. */
int ofb_kitti_decode(const uint16_t *png_bgr, float *flow, uint8_t *valid, int N, int H, int W, void *stream);

/* Bilinear resize with align_corners = False and the scale factor used for the source index, as
 * torch.nn.functional.interpolate(scale_factor=...) does it (resize_flow, utils.py:878-916; Flow.resize,
 * flow_class.py:694-712).  Output extent along each axis = ofb_resize_output_size(in, scale).
 *   flow (N,2,H,W) -> out (N,2,Ho,Wo) with the x component times scale_x and the y component times scale_y
 *   mask (N,H,W)   -> out_mask (N,Ho,Wo) = round-half-even of the resized 0/1 mask
 * Either pair may be NULL.  ofb_resize_bwd is the adjoint of the flow part (gin is zero-filled by the call). */
int ofb_resize_output_size(int in_size, double scale);
int ofb_resize_fwd(const float *flow, const uint8_t *mask, double scale_y, double scale_x,
                   float *out, uint8_t *out_mask, int N, int H, int W, void *stream);
int ofb_resize_bwd(const float *gout, double scale_y, double scale_x, float *gin,
                   int N, int H, int W, void *stream);

/* Row-band sharding, halo step: dst[n, k] += src[n, k] for k < floats_per_item, n < N, with item
 * strides in floats.  `src` is typically PEER memory (another GPU's halo rows, mapped through CUDA IPC / torch
 * symmetric memory), read over NVLink by this kernel; called once per neighbour on one stream, which fixes the order
 * of the additions (own + upper + lower). */
int ofb_add_rows(float *dst, int64_t dst_item_stride, const float *src, int64_t src_item_stride,
                 int N, int64_t floats_per_item, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OFLIB_B200_H */
