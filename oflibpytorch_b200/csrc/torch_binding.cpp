// TORCH_LIBRARY registration of the two hot W ops (warp_t_fwd, warp_t_bwd) from C++.
//
// The kernels live in liboflib_b200.so behind the C ABI of include/oflib_b200.h (no torch types); this shim is the thin
// PyTorch side of it: it allocates the outputs with ATen, turns batch broadcasting into element strides (reference
// utils.py:527-537), picks the entry point and launches on the current CUDA stream -- what oflibpytorch_b200/_ops.py
// does in Python for every other op.  CUDA tensors only; nothing here computes.  The autograd.Function wrappers stay in
// _ops.py and call these ops (torch.ops.oflib_b200n.*).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <string>
#include <tuple>

#include "../../include/oflib_b200.h"

namespace {

using at::Tensor;
using c10::optional;

void check(int status, const char *what)
{
    if (status != 0) {
        const char *msg = ofb_error_string(status);
        TORCH_CHECK(false, "oflibpytorch_b200: ", what, " failed: ", msg ? msg : "?", " (status ", status, ")");
    }
}

Tensor f32c(const Tensor &t)
{
    TORCH_CHECK(t.is_cuda(), "oflibpytorch_b200: tensor is on device '", t.device(),
                "'. This library runs its hot path on CUDA (sm_100a) only and has no CPU fallback; move the tensors "
                "to a CUDA device first.");
    Tensor r = t.scalar_type() == at::kFloat ? t : t.to(at::kFloat);
    return r.is_contiguous() ? r : r.contiguous();
}

optional<Tensor> maskc(const optional<Tensor> &t)
{
    if (!t.has_value()) return t;
    Tensor r = t->scalar_type() == at::kBool ? *t : t->to(at::kBool);
    return r.is_contiguous() ? r : r.contiguous();
}

int64_t batch_of(int64_t n, const optional<Tensor> &t)
{
    if (!t.has_value()) return n;
    const int64_t b = t->size(0);
    TORCH_CHECK(b == n || b == 1 || n == 1, "oflibpytorch_b200: batch sizes ", n, " and ", b, " cannot be broadcast");
    return n > b ? n : b;
}

// batch stride in elements; 0 broadcasts a single item
int64_t bs(const optional<Tensor> &t, int64_t n, int64_t per_item)
{
    if (!t.has_value()) return 0;
    return (t->size(0) == 1 && n > 1) ? 0 : per_item;
}

template <typename T>
const T *cptr(const optional<Tensor> &t) { return t.has_value() ? static_cast<const T *>(t->data_ptr()) : nullptr; }

// (out, valid, out_flags, dtaps): see _ops._warp_t_fwd
std::tuple<Tensor, Tensor, Tensor, Tensor> warp_t_fwd(const Tensor &flow_in, double sign, const optional<Tensor> &src_in,
                                                      const optional<Tensor> &src_mask_in,
                                                      const optional<Tensor> &flow_mask_in, double add_flow, double thresh,
                                                      bool want_valid, const optional<Tensor> &sel, bool want_flags,
                                                      bool want_dtaps)
{
    const Tensor flow = f32c(flow_in);
    optional<Tensor> src;
    if (src_in.has_value()) src = f32c(*src_in);
    const optional<Tensor> src_mask = maskc(src_mask_in), flow_mask = maskc(flow_mask_in);
    const int64_t h = flow.size(2), w = flow.size(3), hw = h * w;
    const int64_t c = src.has_value() ? src->size(1) : 0;
    int64_t n = flow.size(0);
    n = batch_of(n, src);
    n = batch_of(n, src_mask);
    n = batch_of(n, flow_mask);
    const c10::cuda::CUDAGuard guard(flow.device());
    const auto fopt = flow.options();
    Tensor out = at::empty({n, c, h, w}, fopt);
    Tensor valid = want_valid ? at::empty({n, h, w}, fopt.dtype(at::kBool)) : at::empty({0}, fopt.dtype(at::kBool));
    want_dtaps = want_dtaps && c >= 1 && c <= 4 && h <= 65536 && w <= 65536;
    Tensor dtaps = want_dtaps ? at::empty({n, 2 * c, h, w}, fopt) : at::empty({0}, fopt);
    Tensor buf = at::empty({0}, fopt.dtype(at::kInt)), out_flags = buf;
    void *stream = at::cuda::getCurrentCUDAStream().stream();
    const float *srcp = c ? static_cast<const float *>(src->data_ptr()) : nullptr;
    float *outp = c ? static_cast<float *>(out.data_ptr()) : nullptr;
    uint8_t *validp = want_valid ? static_cast<uint8_t *>(valid.data_ptr()) : nullptr;
    const optional<Tensor> flow_o(flow);
    if (sel.has_value() || want_flags || want_dtaps) {
        if (want_flags) {      // flags (zeroed) followed by the kernel's workspace (ofb_warp_t_flags_ints)
            buf = at::empty({ofb_warp_t_flags_ints((int)n, (int)h, (int)w)}, fopt.dtype(at::kInt));
            out_flags = buf.narrow(0, 0, n + 1);
            out_flags.zero_();
        }
        int32_t *flagp = want_flags ? static_cast<int32_t *>(buf.data_ptr()) : nullptr;
        if (want_dtaps)
            check(ofb_warp_t_fwd_d(static_cast<const float *>(flow.data_ptr()), bs(flow_o, n, 2 * hw), (float)sign, srcp,
                                   bs(src, n, c * hw), cptr<uint8_t>(src_mask), bs(src_mask, n, hw),
                                   cptr<uint8_t>(flow_mask), bs(flow_mask, n, hw), (float)add_flow, (float)thresh, outp,
                                   validp, (int)n, (int)c, (int)h, (int)w, cptr<int32_t>(sel), flagp,
                                   static_cast<float *>(dtaps.data_ptr()), stream),
                  "ofb_warp_t_fwd_d");
        else
            check(ofb_warp_t_fwd_ex(static_cast<const float *>(flow.data_ptr()), bs(flow_o, n, 2 * hw), (float)sign, srcp,
                                    bs(src, n, c * hw), cptr<uint8_t>(src_mask), bs(src_mask, n, hw),
                                    cptr<uint8_t>(flow_mask), bs(flow_mask, n, hw), (float)add_flow, (float)thresh, outp,
                                    validp, (int)n, (int)c, (int)h, (int)w, cptr<int32_t>(sel), flagp, stream),
                  "ofb_warp_t_fwd_ex");
    } else {
        check(ofb_warp_t_fwd(static_cast<const float *>(flow.data_ptr()), bs(flow_o, n, 2 * hw), (float)sign, srcp,
                             bs(src, n, c * hw), cptr<uint8_t>(src_mask), bs(src_mask, n, hw), cptr<uint8_t>(flow_mask),
                             bs(flow_mask, n, hw), (float)add_flow, (float)thresh, outp, validp, (int)n, (int)c, (int)h,
                             (int)w, stream),
              "ofb_warp_t_fwd");
    }
    return std::make_tuple(out, valid, out_flags, dtaps);
}

// (gflow, gsrc): see _ops._warp_t_bwd
std::tuple<Tensor, Tensor> warp_t_bwd(const Tensor &flow_in, double sign, const Tensor &src_in, const Tensor &gout_in,
                                      double add_flow, bool need_gflow, bool need_gsrc, const optional<Tensor> &sel,
                                      const optional<Tensor> &dtaps)
{
    const Tensor flow = f32c(flow_in), src = f32c(src_in);
    const int64_t n = gout_in.size(0), c = gout_in.size(1), h = gout_in.size(2), w = gout_in.size(3), hw = h * w;
    const c10::cuda::CUDAGuard guard(flow.device());
    const auto fopt = flow.options();
    Tensor gflow = need_gflow ? at::empty({n, 2, h, w}, fopt) : at::empty({0}, fopt);
    Tensor gsrc = need_gsrc ? at::empty(src.sizes(), fopt) : at::empty({0}, fopt);
    void *stream = at::cuda::getCurrentCUDAStream().stream();
    const optional<Tensor> flow_o(flow), src_o(src);
    const float *flowp = static_cast<const float *>(flow.data_ptr()), *srcp = static_cast<const float *>(src.data_ptr());
    float *gflowp = need_gflow ? static_cast<float *>(gflow.data_ptr()) : nullptr;
    float *gsrcp = need_gsrc ? static_cast<float *>(gsrc.data_ptr()) : nullptr;
    const bool has_d = dtaps.has_value() && dtaps->numel() > 0 && need_gflow;
    const float *dp = has_d ? static_cast<const float *>(dtaps->data_ptr()) : nullptr;
    // the gradient of out.sum() arrives as a stride-0 view of one value: the default adjoint reads that value instead
    // of a materialised (N, C, H, W) copy
    bool bcast = gout_in.scalar_type() == at::kFloat && gout_in.numel() > 1 && need_gsrc && h <= 65536 && w <= 65536;
    for (int d = 0; bcast && d < 4; ++d) bcast = gout_in.stride(d) == 0;
    if (bcast) {
        const int st = ofb_warp_t_bwd_d(flowp, bs(flow_o, n, 2 * hw), (float)sign, srcp, bs(src_o, n, c * hw),
                                        static_cast<const float *>(gout_in.data_ptr()), (float)add_flow, gflowp, gsrcp,
                                        bs(src_o, n, c * hw), (int)n, (int)c, (int)h, (int)w, cptr<int32_t>(sel), dp, 1,
                                        stream);
        if (st == 0) return std::make_tuple(gflow, gsrc);
        if (st != OFB_EINVAL) check(st, "ofb_warp_t_bwd_d");   // OFB_EINVAL: another adjoint kernel is selected
    }
    const Tensor gout = f32c(gout_in);
    const float *gp = static_cast<const float *>(gout.data_ptr());
    if (has_d)
        check(ofb_warp_t_bwd_d(flowp, bs(flow_o, n, 2 * hw), (float)sign, srcp, bs(src_o, n, c * hw), gp, (float)add_flow,
                               gflowp, gsrcp, bs(src_o, n, c * hw), (int)n, (int)c, (int)h, (int)w, cptr<int32_t>(sel), dp, 0,
                               stream),
              "ofb_warp_t_bwd_d");
    else if (sel.has_value())
        check(ofb_warp_t_bwd_ex(flowp, bs(flow_o, n, 2 * hw), (float)sign, srcp, bs(src_o, n, c * hw), gp, (float)add_flow,
                                gflowp, gsrcp, bs(src_o, n, c * hw), (int)n, (int)c, (int)h, (int)w, cptr<int32_t>(sel),
                                stream),
              "ofb_warp_t_bwd_ex");
    else
        check(ofb_warp_t_bwd(flowp, bs(flow_o, n, 2 * hw), (float)sign, srcp, bs(src_o, n, c * hw), gp, (float)add_flow,
                             gflowp, gsrcp, bs(src_o, n, c * hw), (int)n, (int)c, (int)h, (int)w, stream),
              "ofb_warp_t_bwd");
    return std::make_tuple(gflow, gsrc);
}

}  // namespace

TORCH_LIBRARY(oflib_b200n, m)
{
    m.def("warp_t_fwd(Tensor flow, float sign, Tensor? src, Tensor? src_mask, Tensor? flow_mask, float add_flow, "
          "float thresh, bool want_valid, Tensor? sel, bool want_flags, bool want_dtaps) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("warp_t_bwd(Tensor flow, float sign, Tensor src, Tensor gout, float add_flow, bool need_gflow, bool need_gsrc, "
          "Tensor? sel, Tensor? dtaps) -> (Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(oflib_b200n, CUDA, m)
{
    m.impl("warp_t_fwd", &warp_t_fwd);
    m.impl("warp_t_bwd", &warp_t_bwd);
}
