"""torch.library registration of the sm_100a kernels and their autograd wrappers.

``torch.ops.oflib_b200.*`` are thin launchers over the C ABI (include/oflib_b200.h): they allocate the
outputs with torch (device memory + caching allocator are torch's job here), resolve batch broadcasting
into element strides and launch on the current CUDA stream.  CUDA tensors only -- there is no CPU kernel
registered and no fallback.

Two autograd.Functions give the two primitives of the hot path their hand-written adjoints:
``warp_t`` (W, reference utils.py:541-555) and ``splat_s`` (S, reference utils.py:1061-1205).  Boolean
outputs are marked non-differentiable.
"""
import ctypes
from typing import Optional, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _native

VALID_NONE, VALID_GT, VALID_EQ1 = 0, 1, 2

_lib = torch.library.Library("oflib_b200", "DEF")
_lib.define("flow_stats(Tensor flow, Tensor? mask) -> Tensor")
_lib.define("flow_stats_batch(Tensor flow, Tensor? mask) -> Tensor")
_lib.define("warp_t_fwd(Tensor flow, float sign, Tensor? src, Tensor? src_mask, Tensor? flow_mask, "
            "float add_flow, float thresh, bool want_valid, Tensor? sel, bool want_flags, bool want_dtaps) -> "
            "(Tensor, Tensor, Tensor, Tensor)")
_lib.define("early_decide(Tensor? fflags, Tensor? pflags, int rules, int n) -> Tensor")
_lib.define("warp_t_bwd(Tensor flow, float sign, Tensor src, Tensor gout, float add_flow, "
            "bool need_gflow, bool need_gsrc, Tensor? sel, Tensor? dtaps) -> (Tensor, Tensor)")
_lib.define("splat_s_fwd(Tensor cx, Tensor cy, bool relative, float sign, Tensor? data, Tensor? data_mask, "
            "Tensor? flow_mask, bool occlude_zero, int valid_mode, bool want_aux, int[] window) -> "
            "(Tensor, Tensor, Tensor, Tensor, Tensor)")
_lib.define("splat_s_bwd(Tensor cx, Tensor cy, bool relative, float sign, Tensor data, Tensor? flow_mask, "
            "bool occlude_zero, Tensor out, Tensor density, Tensor? unocc, Tensor? gout, Tensor? gdensity, "
            "bool need_gcoord, bool need_gdata, int[] window) -> (Tensor, Tensor)")


# The two hot W ops are also registered from C++ (csrc/torch_binding.cpp, TORCH_LIBRARY oflib_b200n): same C-ABI calls,
# outputs allocated by ATen, no ctypes on the way.  The autograd wrapper below dispatches to them when the shim is built
# (csrc/build.sh builds it); OFB_NO_TORCH_BINDING=1 keeps the Python-registered twins (tuning / comparison).
def _load_native_ops():
    import os
    path = os.path.join(os.path.dirname(_native.LIB_PATH), 'liboflib_b200_torch.so')
    if os.environ.get('OFB_NO_TORCH_BINDING') or os.environ.get('OFB_LIB_OVERRIDE') or not os.path.exists(path):
        return None
    _native.lib()                      # liboflib_b200.so first: a missing kernel library must fail loudly, by name
    torch.ops.load_library(path)
    return torch.ops.oflib_b200n


_NATIVE = _load_native_ops()


def _native_op(name: str, abi_name: str, *args):
    """Call a C++-registered op, bracketed by CUDA events like _call when bench.py profiles."""
    fn = getattr(_NATIVE, name)
    if PROFILE is None:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = fn(*args)
    e1.record()
    PROFILE.append((abi_name, e0, e1))
    return res


# Optional per-call device timing (bench.py): when PROFILE is a list, every C-ABI call is bracketed by CUDA events
# recorded on the launching stream and (name, start, end) is appended.  None (default) = no events.
PROFILE = None


def _call(name: str, *args):
    fn = getattr(_native.lib(), name)
    if PROFILE is None:
        _native.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _native.check(fn(*args), name)
    e1.record()
    PROFILE.append((name, e0, e1))


def _try_call(name: str, *args) -> int:
    """_call that hands the status back instead of raising (for entry points with a documented OFB_EINVAL answer)."""
    fn = getattr(_native.lib(), name)
    if PROFILE is None:
        return int(fn(*args))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    status = int(fn(*args))
    e1.record()
    if status == 0:
        PROFILE.append((name, e0, e1))
    return status


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            "oflibpytorch_b200: %s is on device '%s'. This library runs its hot path on CUDA (sm_100a) only and "
            "has no CPU fallback; move the tensors to a CUDA device first." % (what, t.device))


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _maskc(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.bool:
        t = t.to(torch.bool)
    return t if t.is_contiguous() else t.contiguous()


def _batch(n: int, *tensors) -> int:
    for t in tensors:
        if t is not None:
            b = t.shape[0]
            if b != n and b != 1 and n != 1:
                raise ValueError("oflibpytorch_b200: batch sizes %d and %d cannot be broadcast" % (n, b))
            n = max(n, b)
    return n


def _bs(t: Optional[torch.Tensor], n: int, per_item: int) -> int:
    """Batch stride in elements; 0 broadcasts a single item (reference utils.py:527-537)."""
    if t is None:
        return 0
    return 0 if (t.shape[0] == 1 and n > 1) else per_item


# ----------------------------------------------------------------------------------------------------
# raw ops
# ----------------------------------------------------------------------------------------------------

def _flow_stats(flow: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    require_cuda(flow, "flow")
    flow = _f32c(flow)
    mask = _maskc(mask)
    n, _, h, w = flow.shape
    if mask is not None and mask.shape[0] != n:
        mask = mask.expand(n, -1, -1).contiguous()
    with torch.cuda.device(flow.device):
        flags = torch.zeros((n,), dtype=torch.int32, device=flow.device)
        _call("ofb_flow_stats", _p(flow), _p(mask), _p(flags), n, h, w, _stream())
    return flags


def _flow_stats_batch(flow: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """(N + 1,) int32 on the device: per-item OFB_STAT_* bits and, last, their OR over the batch.  No sync."""
    require_cuda(flow, "flow")
    flow = _f32c(flow)
    mask = _maskc(mask)
    n, _, h, w = flow.shape
    if mask is not None and mask.shape[0] != n:
        mask = mask.expand(n, -1, -1).contiguous()
    with torch.cuda.device(flow.device):
        flags = torch.zeros((n + 1,), dtype=torch.int32, device=flow.device)
        _call("ofb_flow_stats_batch", _p(flow), _p(mask), _p(flags), n, h, w, _stream())
    return flags


def early_rule(which: str, stat_bit: int, action: int) -> int:
    """One rule byte of ofb_warp_t_fwd_ex: fires when `stat_bit` (an OFB_STAT_* value) is clear in the batch summary of
    the warping flow's ('f') or the payload flow's ('p') flags.  action 1: identity (+ add_flow*flow), both masks;
    2: the payload and its mask; 3: add_flow*flow and the flow mask."""
    return 0x80 | (0x20 if which == 'p' else 0) | ((int(stat_bit).bit_length() - 1) << 2) | action


def early_rules(*rules: int) -> int:
    prog = 0
    for k, r in enumerate(rules):
        prog |= r << (8 * k)
    return prog


def _early_decide(fflags, pflags, rules, n) -> torch.Tensor:
    """(1,) int32 on the device: which early-out (0 = none) the rules select for these flags.  No sync."""
    t = fflags if fflags is not None else pflags
    with torch.cuda.device(t.device):
        sel = torch.empty((1,), dtype=torch.int32, device=t.device)
        _call("ofb_early_decide", _p(fflags), _p(pflags), int(rules), int(n), _p(sel), _stream())
    return sel


# Saving the bilinear derivatives in the forward (ofb_warp_t_fwd_d) makes the adjoint's flow gradient a stream instead
# of a second gather of the payload taps.  Measured on B200 (tools/ab_step.py, batch 32 x 3 x 1080p): adjoint 1.32 ->
# 1.20 ms, forward 0.57 -> 0.84 ms (8C more bytes per pixel written and read back: the pair turns HBM-bound) -- a net
# loss, hence opt-in.
SAVE_DERIVATIVES = False


def set_save_derivatives(on: bool) -> None:
    global SAVE_DERIVATIVES
    SAVE_DERIVATIVES = bool(on)


def _warp_t_fwd(flow, sign, src, src_mask, flow_mask, add_flow, thresh, want_valid, sel=None,
                want_flags=False, want_dtaps=False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    require_cuda(flow, "flow")
    flow = _f32c(flow)
    src = None if src is None else _f32c(src)
    src_mask, flow_mask = _maskc(src_mask), _maskc(flow_mask)
    _, _, h, w = flow.shape
    c = 0 if src is None else src.shape[1]
    n = _batch(flow.shape[0], src, src_mask, flow_mask)
    hw = h * w
    dev = flow.device
    with torch.cuda.device(dev):
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
        valid = torch.empty((n, h, w) if want_valid else (0,), dtype=torch.bool, device=dev)
        want_dtaps = bool(want_dtaps) and 1 <= c <= 4 and h <= 65536 and w <= 65536
        dtaps = torch.empty((n, 2 * c, h, w) if want_dtaps else (0,), dtype=torch.float32, device=dev)
        head = (_p(flow), _bs(flow, n, 2 * hw), sign,
                _p(src) if c else None, _bs(src, n, c * hw),
                _p(src_mask), _bs(src_mask, n, hw), _p(flow_mask), _bs(flow_mask, n, hw),
                add_flow, thresh, _p(out) if c else None, _p(valid) if want_valid else None,
                n, c, h, w)
        if sel is not None or want_flags or want_dtaps:
            if want_flags:      # flags (zeroed) followed by the kernel's workspace (ofb_warp_t_flags_ints)
                buf = torch.empty((int(_native.lib().ofb_warp_t_flags_ints(n, h, w)),), dtype=torch.int32, device=dev)
                out_flags = buf[:n + 1]
                out_flags.zero_()
            else:
                buf = out_flags = torch.empty((0,), dtype=torch.int32, device=dev)
            if want_dtaps:
                _call("ofb_warp_t_fwd_d", *head, _p(sel), _p(buf) if want_flags else None, _p(dtaps), _stream())
            else:
                _call("ofb_warp_t_fwd_ex", *head, _p(sel), _p(buf) if want_flags else None, _stream())
        else:
            out_flags = torch.empty((0,), dtype=torch.int32, device=dev)
            _call("ofb_warp_t_fwd", *head, _stream())
    return out, valid, out_flags, dtaps


def _warp_t_bwd(flow, sign, src, gout, add_flow, need_gflow, need_gsrc, sel=None,
                dtaps=None) -> Tuple[torch.Tensor, torch.Tensor]:
    require_cuda(flow, "flow")
    flow, src = _f32c(flow), _f32c(src)
    n, c, h, w = gout.shape
    hw = h * w
    dev = flow.device
    # the gradient of out.sum() arrives as a stride-0 view of one value: the default adjoint reads that value instead
    # of a materialised (N, C, H, W) copy (100 MB per 4K frame)
    bcast = (gout.dtype == torch.float32 and gout.numel() > 1 and all(st == 0 for st in gout.stride())
             and need_gsrc and h <= 65536 and w <= 65536)
    with torch.cuda.device(dev):
        gflow = torch.empty((n, 2, h, w) if need_gflow else (0,), dtype=torch.float32, device=dev)
        gsrc = torch.empty(tuple(src.shape) if need_gsrc else (0,), dtype=torch.float32, device=dev)

        def head(g):
            return (_p(flow), _bs(flow, n, 2 * hw), sign, _p(src), _bs(src, n, c * hw), _p(g), add_flow,
                    _p(gflow) if need_gflow else None, _p(gsrc) if need_gsrc else None, _bs(src, n, c * hw),
                    n, c, h, w)

        has_d = dtaps is not None and dtaps.numel() > 0 and need_gflow
        if bcast:
            status = _try_call("ofb_warp_t_bwd_d", *head(gout), _p(sel), _p(dtaps) if has_d else None, 1, _stream())
            if status == 0:
                return gflow, gsrc
            if status != _native.OFB_EINVAL:     # OFB_EINVAL: another adjoint kernel is selected, it needs a dense gout
                _native.check(status, "ofb_warp_t_bwd_d")
        gout = _f32c(gout)
        if has_d:
            _call("ofb_warp_t_bwd_d", *head(gout), _p(sel), _p(dtaps), 0, _stream())
        elif sel is not None:
            _call("ofb_warp_t_bwd_ex", *head(gout), _p(sel), _stream())
        else:
            _call("ofb_warp_t_bwd", *head(gout), _stream())
    return gflow, gsrc


def _coord_planes(cx: torch.Tensor, cy: torch.Tensor):
    """Two (N,H,W) fp32 planes with unit inner strides and a common batch stride (views of one flow tensor
    qualify, so flow[:, 0], flow[:, 1] are passed without a copy)."""
    h, w = cx.shape[-2:]

    def ok(t):
        return t.dtype == torch.float32 and t.stride(-1) == 1 and t.stride(-2) == w

    if not (ok(cx) and ok(cy) and cx.stride(0) == cy.stride(0)):
        cx, cy = _f32c(cx), _f32c(cy)
    return cx, cy


def _splat_s_fwd(cx, cy, relative, sign, data, data_mask, flow_mask, occlude_zero, valid_mode, want_aux, window):
    require_cuda(cx, "flow")
    cx, cy = _coord_planes(cx, cy)
    data = None if data is None else _f32c(data)
    data_mask, flow_mask = _maskc(data_mask), _maskc(flow_mask)
    h, w = cx.shape[-2:]
    hw = h * w
    c = 0 if data is None else data.shape[1]
    n = _batch(cx.shape[0], data, data_mask, flow_mask)
    dev = cx.device
    want_valid = valid_mode != VALID_NONE
    with torch.cuda.device(dev):
        acc_floats = int(_native.lib().ofb_splat_s_workspace(n, c, h, w, int(data_mask is not None)))
        acc = torch.empty((acc_floats,), dtype=torch.float32, device=dev)
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
        # density / warped mask / un-occlusion mask are only materialised when the caller or the adjoint needs them
        density = torch.empty((n, h, w) if want_aux else (0,), dtype=torch.float32, device=dev)
        wmask = torch.empty((n, h, w) if want_aux else (0,), dtype=torch.bool, device=dev)
        unocc = torch.empty((n, h, w) if (occlude_zero and want_aux) else (0,), dtype=torch.bool, device=dev)
        valid = torch.empty((n, h, w) if want_valid else (0,), dtype=torch.bool, device=dev)
        head = (_p(cx), _p(cy), 0 if (cx.shape[0] == 1 and n > 1) else cx.stride(0))
        tail = (_p(data) if c else None, _bs(data, n, c * hw),
                _p(data_mask), _bs(data_mask, n, hw), _p(flow_mask), _bs(flow_mask, n, hw),
                int(occlude_zero), _p(acc), acc_floats, _p(out) if c else None, _p(density) if want_aux else None,
                _p(wmask) if want_aux else None, _p(unocc) if (occlude_zero and want_aux) else None,
                _p(valid) if want_valid else None, valid_mode, n, c, h, w)
        if len(window):      # (y0, x0, frame_h, frame_w): the tensors are a window of a larger frame
            _call("ofb_splat_s_win_fwd", *head, sign, *tail, _win_arr4(window), _stream())
        else:
            _call("ofb_splat_s_fwd", *head, int(relative), sign, *tail, _stream())
    return out, density, wmask, unocc, valid


def _win_arr4(window):
    if len(window) != 4:
        raise ValueError("window needs 4 integers (y0, x0, frame_h, frame_w)")
    return (ctypes.c_int * 4)(*[int(v) for v in window])


def _splat_s_bwd(cx, cy, relative, sign, data, flow_mask, occlude_zero, out, density, unocc, gout, gdensity,
                 need_gcoord, need_gdata, window):
    require_cuda(cx, "flow")
    cx, cy = _coord_planes(cx, cy)
    data, out, density = _f32c(data), _f32c(out), _f32c(density)
    flow_mask = _maskc(flow_mask)
    n, c, h, w = out.shape
    hw = h * w
    dev = cx.device
    with torch.cuda.device(dev):
        gout = torch.zeros_like(out) if gout is None else _f32c(gout)
        gdensity = None if gdensity is None else _f32c(gdensity)
        gcoord = torch.empty((n, 2, h, w) if need_gcoord else (0,), dtype=torch.float32, device=dev)
        gdata = torch.empty((n, c, h, w) if need_gdata else (0,), dtype=torch.float32, device=dev)
        work_floats = int(_native.lib().ofb_splat_s_bwd_workspace(n, c, h, w))
        work = torch.empty((work_floats,), dtype=torch.float32, device=dev) if work_floats else None
        gcx = ctypes.c_void_p(gcoord.data_ptr()) if need_gcoord else None
        gcy = ctypes.c_void_p(gcoord.data_ptr() + 4 * hw) if need_gcoord else None
        head = (_p(cx), _p(cy), 0 if (cx.shape[0] == 1 and n > 1) else cx.stride(0))
        tail = (_p(data), _bs(data, n, c * hw), _p(flow_mask), _bs(flow_mask, n, hw), int(occlude_zero),
                _p(out), _p(density), _p(unocc) if (unocc is not None and unocc.numel()) else None,
                _p(gout), _p(gdensity), gcx, gcy, 2 * hw, _p(gdata) if need_gdata else None,
                _p(work), work_floats, n, c, h, w)
        if len(window):
            _call("ofb_splat_s_win_bwd", *head, sign, *tail, _win_arr4(window), _stream())
        else:
            _call("ofb_splat_s_bwd", *head, int(relative), sign, *tail, _stream())
    return gcoord, gdata


_lib.impl("flow_stats", _flow_stats, "CUDA")
_lib.impl("flow_stats_batch", _flow_stats_batch, "CUDA")
_lib.impl("early_decide", _early_decide, "CUDA")
_lib.impl("warp_t_fwd", _warp_t_fwd, "CUDA")
_lib.impl("warp_t_bwd", _warp_t_bwd, "CUDA")
_lib.impl("splat_s_fwd", _splat_s_fwd, "CUDA")
_lib.impl("splat_s_bwd", _splat_s_bwd, "CUDA")

_ops = torch.ops.oflib_b200


# ----------------------------------------------------------------------------------------------------
# autograd wrappers
# ----------------------------------------------------------------------------------------------------

def _reduce_to(g: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Sum a per-item gradient over the batch when the input was a broadcast single item."""
    if like.shape[0] == 1 and g.shape[0] != 1:
        g = g.sum(0, keepdim=True)
    return g


class _WarpT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, src, src_mask, flow_mask, sign, add_flow, thresh, want_valid, sel, want_flags):
        want_dtaps = SAVE_DERIVATIVES and ctx.needs_input_grad[0]
        if _NATIVE is not None:
            require_cuda(flow, "flow")
            out, valid, out_flags, dtaps = _native_op("warp_t_fwd", "ofb_warp_t_fwd", flow, sign, src, src_mask,
                                                      flow_mask, add_flow, thresh, want_valid, sel, want_flags, want_dtaps)
        else:
            out, valid, out_flags, dtaps = _ops.warp_t_fwd(flow, sign, src, src_mask, flow_mask, add_flow, thresh,
                                                           want_valid, sel, want_flags, want_dtaps)
        ctx.save_for_backward(flow, src, sel, dtaps if dtaps.numel() else None)
        ctx.sign, ctx.add_flow = sign, add_flow
        ctx.mark_non_differentiable(valid, out_flags)
        ctx.set_materialize_grads(False)     # no zero-filled stand-in for the gradient of the boolean output
        return out, valid, out_flags

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, _gvalid, _gflags):
        flow, src, sel, dtaps = ctx.saved_tensors
        need_gflow, need_gsrc = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gflow = gsrc = None
        if gout is not None and (need_gflow or need_gsrc):
            if _NATIVE is not None:
                gf, gs = _native_op("warp_t_bwd", "ofb_warp_t_bwd", flow, ctx.sign, src, gout, ctx.add_flow, need_gflow,
                                    need_gsrc, sel, dtaps)
            else:
                gf, gs = _ops.warp_t_bwd(flow, ctx.sign, src, gout, ctx.add_flow, need_gflow, need_gsrc, sel, dtaps)
            if need_gflow:
                gflow = _reduce_to(gf, flow)
            if need_gsrc:
                gsrc = gs
        return gflow, gsrc, None, None, None, None, None, None, None, None


class _SplatS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cx, cy, data, data_mask, flow_mask, relative, sign, occlude_zero, valid_mode):
        out, density, wmask, unocc, valid = _ops.splat_s_fwd(cx, cy, relative, sign, data, data_mask, flow_mask,
                                                             occlude_zero, valid_mode, True, [])
        ctx.save_for_backward(cx, cy, data, flow_mask, out, density, unocc)
        ctx.relative, ctx.sign, ctx.occlude_zero = relative, sign, occlude_zero
        ctx.mark_non_differentiable(wmask, valid)
        ctx.set_materialize_grads(False)     # an unused density output must not cost a zero-filled gradient
        return out, density, wmask, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, gdensity, _gw, _gv):
        cx, cy, data, flow_mask, out, density, unocc = ctx.saved_tensors
        need_gc = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_gd = ctx.needs_input_grad[2]
        gcx = gcy = gdata = None
        if (gout is not None or gdensity is not None) and (need_gc or need_gd):
            gcoord, gd = _ops.splat_s_bwd(cx, cy, ctx.relative, ctx.sign, data, flow_mask, ctx.occlude_zero,
                                          out, density, unocc, gout, gdensity, need_gc, need_gd, [])
            if need_gc:
                gcx = _reduce_to(gcoord[:, 0], cx)
                gcy = _reduce_to(gcoord[:, 1], cy)
            if need_gd:
                gdata = _reduce_to(gd, data)
        return gcx, gcy, gdata, None, None, None, None, None, None


def flow_stats(flow: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-item OFB_STAT_* bits (int32, on the device)."""
    return _ops.flow_stats(flow.detach(), mask)


def flow_stats_batch(flow: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N + 1,) int32 on the device: per-item OFB_STAT_* bits, then their OR over the batch."""
    return _ops.flow_stats_batch(flow.detach(), mask)


def warp_t(flow, src, src_mask=None, flow_mask=None, sign=1.0, add_flow=0.0, thresh=0.99999,
           want_valid=False, fflags=None, pflags=None, early=0, want_flags=False):
    """W.  Returns (out, valid), or (out, valid, out_flags) with `want_flags`; valid is an empty tensor unless
    want_valid.  src may be None (C = 0).  fflags / pflags / early: device-side early-outs (ofb_early_decide)."""
    sel = None
    if early:
        f = flow.detach()
        n = _batch(f.shape[0], src, src_mask, flow_mask)
        for fl in (fflags, pflags):
            if fl is not None and fl.shape[0] != n + 1:
                raise ValueError("oflibpytorch_b200: device flags of a batch of %d used with a batch of %d"
                                 % (fl.shape[0] - 1, n))
        sel = _ops.early_decide(fflags, pflags, int(early), n)
    if src is None:
        out, valid, fl, _ = _ops.warp_t_fwd(flow.detach(), float(sign), None, src_mask, flow_mask, 0.0, float(thresh),
                                            True, sel, False, False)
    else:
        out, valid, fl = _WarpT.apply(flow, src, src_mask, flow_mask, float(sign), float(add_flow), float(thresh),
                                      bool(want_valid), sel, bool(want_flags))
    return (out, valid, fl) if want_flags else (out, valid)


def splat_flow(flow, data, data_mask=None, flow_mask=None, sign=1.0, occlude_zero=True, valid_mode=VALID_NONE,
               want_aux=False, window=()):
    """See _splat_flow; `window` = (y0, x0, frame_h, frame_w) makes the tensors a window of a larger frame
    (ofb_splat_s_win_fwd)."""
    return _splat_flow(flow, data, data_mask, flow_mask, sign, occlude_zero, valid_mode, want_aux,
                       [int(v) for v in window])


def _splat_flow(flow, data, data_mask, flow_mask, sign, occlude_zero, valid_mode, want_aux, window):
    """S driven by a flow field (N,2,H,W).  Returns (out, density, warped_mask, valid); density and warped_mask are
    empty tensors unless `want_aux` (or autograd needs them)."""
    if data is None:
        f = flow.detach()
        out, density, wmask, _u, valid = _ops.splat_s_fwd(f[:, 0], f[:, 1], True, float(sign), None, data_mask,
                                                          flow_mask, bool(occlude_zero), int(valid_mode), bool(want_aux),
                                                          window)
        return out, density, wmask, valid
    if not want_aux and not (torch.is_grad_enabled() and (flow.requires_grad or data.requires_grad)):
        f = flow.detach()
        out, density, wmask, _u, valid = _ops.splat_s_fwd(f[:, 0], f[:, 1], True, float(sign), data.detach(), data_mask,
                                                          flow_mask, bool(occlude_zero), int(valid_mode), False, window)
        return out, density, wmask, valid
    return _SplatFlow.apply(flow, data, data_mask, flow_mask, float(sign), bool(occlude_zero), int(valid_mode), window)


class _SplatFlow(torch.autograd.Function):
    """S with the coordinates given as the two planes of one flow tensor (no copies)."""

    @staticmethod
    def forward(ctx, flow, data, data_mask, flow_mask, sign, occlude_zero, valid_mode, window):
        flow = _f32c(flow)
        out, density, wmask, unocc, valid = _ops.splat_s_fwd(flow[:, 0], flow[:, 1], True, sign, data, data_mask,
                                                             flow_mask, occlude_zero, valid_mode, True, window)
        ctx.save_for_backward(flow, data, flow_mask, out, density, unocc)
        ctx.sign, ctx.occlude_zero, ctx.window = sign, occlude_zero, list(window)
        ctx.mark_non_differentiable(wmask, valid)
        ctx.set_materialize_grads(False)
        return out, density, wmask, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, gdensity, _gw, _gv):
        flow, data, flow_mask, out, density, unocc = ctx.saved_tensors
        need_gf, need_gd = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gflow = gdata = None
        if (gout is not None or gdensity is not None) and (need_gf or need_gd):
            gcoord, gd = _ops.splat_s_bwd(flow[:, 0], flow[:, 1], True, ctx.sign, data, flow_mask,
                                          ctx.occlude_zero, out, density, unocc, gout, gdensity, need_gf, need_gd,
                                          ctx.window)
            if need_gf:
                gflow = _reduce_to(gcoord, flow)
            if need_gd:
                gdata = _reduce_to(gd, data)
        return gflow, gdata, None, None, None, None, None, None


_lib.define("warp_t_win_fwd(Tensor flow, float sign, Tensor? src, Tensor? src_mask, Tensor? flow_mask, float thresh, "
            "bool want_valid, int[] window) -> (Tensor, Tensor)")
_lib.define("warp_t_win_bwd(Tensor flow, float sign, Tensor src, Tensor gout, bool need_gflow, bool need_gsrc, "
            "int[] window) -> (Tensor, Tensor)")


def _win_arr(window):
    if len(window) != 8:
        raise ValueError("oflibpytorch_b200: window needs 8 integers (y0, x0, gh, gw, sy0, sx0, sh, sw)")
    return (ctypes.c_int * 8)(*[int(v) for v in window])


def _warp_t_win_fwd(flow, sign, src, src_mask, flow_mask, thresh, want_valid, window):
    require_cuda(flow, "flow")
    flow = _f32c(flow)
    src = None if src is None else _f32c(src)
    src_mask, flow_mask = _maskc(src_mask), _maskc(flow_mask)
    _, _, h, w = flow.shape
    sh, sw = int(window[6]), int(window[7])
    c = 0 if src is None else src.shape[1]
    if src is not None and tuple(src.shape[-2:]) != (sh, sw):
        raise ValueError("oflibpytorch_b200: stored source size does not match the window")
    n = _batch(flow.shape[0], src, src_mask, flow_mask)
    dev = flow.device
    with torch.cuda.device(dev):
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
        valid = torch.empty((n, h, w) if want_valid else (0,), dtype=torch.bool, device=dev)
        _call("ofb_warp_t_win_fwd", _p(flow), _bs(flow, n, 2 * h * w), sign,
              _p(src) if c else None, _bs(src, n, c * sh * sw), _p(src_mask), _bs(src_mask, n, sh * sw),
              _p(flow_mask), _bs(flow_mask, n, h * w), 0.0, thresh, _p(out) if c else None,
              _p(valid) if want_valid else None, n, c, h, w, _win_arr(window), _stream())
    return out, valid


def _warp_t_win_bwd(flow, sign, src, gout, need_gflow, need_gsrc, window):
    require_cuda(flow, "flow")
    flow, src, gout = _f32c(flow), _f32c(src), _f32c(gout)
    n, c, h, w = gout.shape
    sh, sw = int(window[6]), int(window[7])
    dev = flow.device
    with torch.cuda.device(dev):
        gflow = torch.empty((n, 2, h, w) if need_gflow else (0,), dtype=torch.float32, device=dev)
        gsrc = torch.empty(tuple(src.shape) if need_gsrc else (0,), dtype=torch.float32, device=dev)
        _call("ofb_warp_t_win_bwd", _p(flow), _bs(flow, n, 2 * h * w), sign, _p(src), _bs(src, n, c * sh * sw),
              _p(gout), 0.0, _p(gflow) if need_gflow else None, _p(gsrc) if need_gsrc else None,
              _bs(src, n, c * sh * sw), n, c, h, w, _win_arr(window), _stream())
    return gflow, gsrc


_lib.impl("warp_t_win_fwd", _warp_t_win_fwd, "CUDA")
_lib.impl("warp_t_win_bwd", _warp_t_win_bwd, "CUDA")


class _WarpTWin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, src, src_mask, flow_mask, sign, thresh, want_valid, window):
        out, valid = _ops.warp_t_win_fwd(flow, sign, src, src_mask, flow_mask, thresh, want_valid, window)
        ctx.save_for_backward(flow, src)
        ctx.sign, ctx.window = sign, list(window)
        ctx.mark_non_differentiable(valid)
        ctx.set_materialize_grads(False)
        return out, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, _gvalid):
        flow, src = ctx.saved_tensors
        need_gflow, need_gsrc = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gflow = gsrc = None
        if gout is not None and (need_gflow or need_gsrc):
            gf, gs = _ops.warp_t_win_bwd(flow, ctx.sign, src, gout, need_gflow, need_gsrc, ctx.window)
            gflow = _reduce_to(gf, flow) if need_gflow else None
            gsrc = gs if need_gsrc else None
        return gflow, gsrc, None, None, None, None, None, None


_lib.define("warp_t_inset_fwd(Tensor flow, Tensor src, Tensor? src_mask, Tensor? flow_mask, float thresh, "
            "bool want_valid, int[] fwin) -> (Tensor, Tensor)")
_lib.define("warp_t_inset_bwd(Tensor flow, Tensor src, Tensor gout, bool need_gflow, bool need_gsrc, int[] fwin) -> "
            "(Tensor, Tensor)")


def _fwin_arr(fwin):
    if len(fwin) != 4:
        raise ValueError("fwin needs 4 integers (y0, x0, flow_h, flow_w)")
    return (ctypes.c_int * 4)(*[int(v) for v in fwin])


def _warp_t_inset_fwd(flow, src, src_mask, flow_mask, thresh, want_valid, fwin):
    require_cuda(flow, "flow")
    flow, src = _f32c(flow), _f32c(src)
    src_mask, flow_mask = _maskc(src_mask), _maskc(flow_mask)
    _, _, fh, fw = flow.shape
    _, c, h, w = src.shape
    n = _batch(flow.shape[0], src, src_mask, flow_mask)
    dev = flow.device
    with torch.cuda.device(dev):
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
        valid = torch.empty((n, h, w) if want_valid else (0,), dtype=torch.bool, device=dev)
        _call("ofb_warp_t_inset_fwd", _p(flow), _bs(flow, n, 2 * fh * fw), _p(src), _bs(src, n, c * h * w),
              _p(src_mask), _bs(src_mask, n, h * w), _p(flow_mask), _bs(flow_mask, n, fh * fw), thresh,
              _p(out), _p(valid) if want_valid else None, n, c, h, w, _fwin_arr(fwin), _stream())
    return out, valid


def _warp_t_inset_bwd(flow, src, gout, need_gflow, need_gsrc, fwin):
    require_cuda(flow, "flow")
    flow, src, gout = _f32c(flow), _f32c(src), _f32c(gout)
    n, c, h, w = gout.shape
    _, _, fh, fw = flow.shape
    dev = flow.device
    with torch.cuda.device(dev):
        gflow = torch.empty((n, 2, fh, fw) if need_gflow else (0,), dtype=torch.float32, device=dev)
        gsrc = torch.empty(tuple(src.shape) if need_gsrc else (0,), dtype=torch.float32, device=dev)
        _call("ofb_warp_t_inset_bwd", _p(flow), _bs(flow, n, 2 * fh * fw), _p(src), _bs(src, n, c * h * w), _p(gout),
              _p(gflow) if need_gflow else None, _p(gsrc) if need_gsrc else None, _bs(src, n, c * h * w),
              n, c, h, w, _fwin_arr(fwin), _stream())
    return gflow, gsrc


_lib.impl("warp_t_inset_fwd", _warp_t_inset_fwd, "CUDA")
_lib.impl("warp_t_inset_bwd", _warp_t_inset_bwd, "CUDA")


class _WarpTInset(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, src, src_mask, flow_mask, thresh, want_valid, fwin):
        out, valid = _ops.warp_t_inset_fwd(flow, src, src_mask, flow_mask, thresh, want_valid, fwin)
        ctx.save_for_backward(flow, src)
        ctx.fwin = list(fwin)
        ctx.mark_non_differentiable(valid)
        ctx.set_materialize_grads(False)
        return out, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, _gvalid):
        flow, src = ctx.saved_tensors
        need_gflow, need_gsrc = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gflow = gsrc = None
        if gout is not None and (need_gflow or need_gsrc):
            gf, gs = _ops.warp_t_inset_bwd(flow, src, gout, need_gflow, need_gsrc, ctx.fwin)
            gflow = _reduce_to(gf, flow) if need_gflow else None
            gsrc = gs if need_gsrc else None
        return gflow, gsrc, None, None, None, None, None


def warp_t_inset(flow, src, top_left, src_mask=None, flow_mask=None, thresh=0.99999, want_valid=False):
    """W of the whole frame of `src` by a flow that covers only the window at `top_left` = (y0, x0) and is zero
    elsewhere; valid is false outside the window (include/oflib_b200.h: ofb_warp_t_inset_fwd)."""
    fwin = [int(top_left[0]), int(top_left[1]), int(flow.shape[-2]), int(flow.shape[-1])]
    return _WarpTInset.apply(flow, src, src_mask, flow_mask, float(thresh), bool(want_valid), fwin)


def warp_t_win(flow, src, window, src_mask=None, flow_mask=None, sign=1.0, thresh=0.99999, want_valid=False):
    """Windowed W (see include/oflib_b200.h: ofb_warp_t_win_fwd).  window = (y0, x0, gh, gw, sy0, sx0, sh, sw)."""
    return _WarpTWin.apply(flow, src, src_mask, flow_mask, float(sign), float(thresh), bool(want_valid),
                           [int(v) for v in window])


_lib.define("sample_pts_fwd(Tensor flow, Tensor pts) -> Tensor")
_lib.define("sample_pts_bwd(Tensor flow, Tensor pts, Tensor gout, bool need_gflow, bool need_gpts) -> (Tensor, Tensor)")


def _sample_pts_fwd(flow, pts):
    require_cuda(flow, "flow")
    flow, pts = _f32c(flow), _f32c(pts)
    n, m, _ = pts.shape
    _, _, h, w = flow.shape
    with torch.cuda.device(flow.device):
        out = torch.empty((n, m, 2), dtype=torch.float32, device=flow.device)
        _call("ofb_sample_pts_fwd", _p(flow), _bs(flow, n, 2 * h * w), _p(pts), _p(out), n, m, h, w, _stream())
    return out


def _sample_pts_bwd(flow, pts, gout, need_gflow, need_gpts):
    require_cuda(flow, "flow")
    flow, pts, gout = _f32c(flow), _f32c(pts), _f32c(gout)
    n, m, _ = pts.shape
    _, _, h, w = flow.shape
    with torch.cuda.device(flow.device):
        gflow = torch.empty(tuple(flow.shape) if need_gflow else (0,), dtype=torch.float32, device=flow.device)
        gpts = torch.empty((n, m, 2) if need_gpts else (0,), dtype=torch.float32, device=flow.device)
        _call("ofb_sample_pts_bwd", _p(flow), _bs(flow, n, 2 * h * w), _p(pts), _p(gout),
              _p(gflow) if need_gflow else None, _bs(flow, n, 2 * h * w), _p(gpts) if need_gpts else None,
              n, m, h, w, _stream())
    return gflow, gpts


_lib.impl("sample_pts_fwd", _sample_pts_fwd, "CUDA")
_lib.impl("sample_pts_bwd", _sample_pts_bwd, "CUDA")


class _SamplePts(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, pts):
        ctx.save_for_backward(flow, pts)
        return _ops.sample_pts_fwd(flow, pts)

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        flow, pts = ctx.saved_tensors
        gf, gp = _ops.sample_pts_bwd(flow, pts, gout, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return (gf if ctx.needs_input_grad[0] else None), (gp if ctx.needs_input_grad[1] else None)


def sample_pts(flow, pts):
    """(N,M,2) flow vectors in (y, x) order interpolated at float points pts (N,M,2) in (y, x) order."""
    return _SamplePts.apply(flow, pts)


def add_rows(dst: torch.Tensor, src: torch.Tensor):
    """dst[l] += src[l] for every leading index l, where dst[l] / src[l] are contiguous runs of floats (views whose
    leading dimension may be strided).  `src` may be PEER memory: the kernel reads it over NVLink (ofb_add_rows)."""
    if dst.dtype != torch.float32 or src.dtype != torch.float32 or dst.shape != src.shape:
        raise ValueError("oflibpytorch_b200: add_rows needs two fp32 tensors of equal shape")
    if dst.numel() == 0:
        return dst
    n = dst.shape[0]
    per = dst[0].numel()
    if not (dst[0].is_contiguous() and src[0].is_contiguous()):
        raise ValueError("oflibpytorch_b200: add_rows needs contiguous items")
    with torch.cuda.device(dst.device):
        _call("ofb_add_rows", _p(dst), dst.stride(0) if n > 1 else per, _p(src), src.stride(0) if n > 1 else per,
              n, per, _stream())
    return dst


def splat_band_accumulate(flow, data, data_mask, flow_mask, sign, occlude_zero, band, out=None):
    """Row-band splat, step 1 (forward only): returns the interleaved accumulator (N, acc_h, W, KP) of this band.
    band = (src_y0, global_h, acc_y0, acc_h); see include/oflib_b200.h ofb_splat_s_band_accumulate.  `out`: where to
    accumulate (e.g. peer-mapped memory, so that the neighbours can read the halo rows in place)."""
    require_cuda(flow, "flow")
    flow = _f32c(flow.detach())
    data = None if data is None else _f32c(data.detach())
    data_mask, flow_mask = _maskc(data_mask), _maskc(flow_mask)
    _, _, h, w = flow.shape
    c = 0 if data is None else data.shape[1]
    n = _batch(flow.shape[0], data, data_mask, flow_mask)
    kp = int(_native.lib().ofb_splat_s_kp(c, int(data_mask is not None)))
    if kp == 0:
        raise ValueError("oflibpytorch_b200: payloads wider than 6 channels are not supported by the row-band splat")
    hw = h * w
    with torch.cuda.device(flow.device):
        if out is not None:
            acc = out
            if tuple(acc.shape) != (n, int(band[3]), w, kp) or not acc.is_contiguous() or acc.dtype != torch.float32:
                raise ValueError("oflibpytorch_b200: accumulator must be a contiguous fp32 (N, acc_h, W, KP) tensor")
        else:
            acc = torch.empty((n, int(band[3]), w, kp), dtype=torch.float32, device=flow.device)
        cx, cy = flow[:, 0], flow[:, 1]
        _call("ofb_splat_s_band_accumulate", _p(cx), _p(cy), 0 if (flow.shape[0] == 1 and n > 1) else 2 * hw, 1, sign,
              _p(data) if c else None, _bs(data, n, c * hw), _p(data_mask), _bs(data_mask, n, hw),
              _p(flow_mask), _bs(flow_mask, n, hw), int(occlude_zero), _p(acc), n, c, h, w,
              (ctypes.c_int * 4)(*[int(v) for v in band]), _stream())
    return acc


def splat_band_normalise(acc, flow, data, data_mask, flow_mask, occlude_zero, valid_mode, band, want_aux=False):
    """Row-band splat, step 2: normalise this rank's own rows from the (halo-summed) accumulator.
    Returns (out (N,C,H,W), valid (N,H,W) or empty); with `want_aux` also (density, unocc) for the adjoint."""
    flow = _f32c(flow.detach())
    data = None if data is None else _f32c(data.detach())
    data_mask, flow_mask = _maskc(data_mask), _maskc(flow_mask)
    _, _, h, w = flow.shape
    c = 0 if data is None else data.shape[1]
    n = acc.shape[0]
    hw = h * w
    want_valid = valid_mode != VALID_NONE
    with torch.cuda.device(flow.device):
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=flow.device)
        valid = torch.empty((n, h, w) if want_valid else (0,), dtype=torch.bool, device=flow.device)
        density = torch.empty((n, h, w) if want_aux else (0,), dtype=torch.float32, device=flow.device)
        unocc = torch.empty((n, h, w) if (want_aux and occlude_zero) else (0,), dtype=torch.bool, device=flow.device)
        cx, cy = flow[:, 0], flow[:, 1]
        acc = acc.contiguous()                 # bound to a name: the pointer must outlive the launch
        _call("ofb_splat_s_band_normalise", _p(cx), _p(cy), 0 if (flow.shape[0] == 1 and n > 1) else 2 * hw, 1,
              _p(data) if c else None, _bs(data, n, c * hw), _p(data_mask), _bs(data_mask, n, hw),
              _p(flow_mask), _bs(flow_mask, n, hw), int(occlude_zero), _p(acc), _p(out) if c else None,
              _p(density) if want_aux else None, None, _p(unocc) if (want_aux and occlude_zero) else None,
              _p(valid) if want_valid else None, valid_mode, n, c, h, w,
              (ctypes.c_int * 4)(*[int(v) for v in band]), _stream())
    if want_aux:
        return out, valid, density, (unocc if occlude_zero else None)
    return out, valid


def splat_band_bwd_prepass(out, density, unocc, gout):
    """Row-band splat adjoint, step 1: G (N,H,W,KG) for this rank's own target rows (ofb_splat_s_band_bwd_prepass)."""
    out, density, gout = _f32c(out), _f32c(density), _f32c(gout)
    n, c, h, w = out.shape
    kg = int(_native.lib().ofb_splat_s_kg(c))
    if kg == 0:
        raise ValueError("oflibpytorch_b200: the row-band splat adjoint supports 1..7 payload channels")
    with torch.cuda.device(out.device):
        g = torch.empty((n, h, w, kg), dtype=torch.float32, device=out.device)
        _call("ofb_splat_s_band_bwd_prepass", _p(out), _p(density), _p(unocc), _p(gout), None, _p(g), n, c, h, w,
              _stream())
    return g


def splat_band_bwd_gather(flow, data, flow_mask, occlude_zero, g_ext, unocc, gout, band, need_gflow, need_gdata,
                          sign=1.0):
    """Row-band splat adjoint, step 3: gradients of this rank's own sources against the halo-extended G
    (ofb_splat_s_band_bwd_gather).  band = (src_y0, global_h, acc_y0, acc_h).  Returns (gflow, gdata)."""
    flow, data, gout = _f32c(flow.detach()), _f32c(data.detach()), _f32c(gout)
    flow_mask = _maskc(flow_mask)
    n, c, h, w = data.shape
    hw = h * w
    with torch.cuda.device(flow.device):
        gflow = torch.empty((n, 2, h, w) if need_gflow else (0,), dtype=torch.float32, device=flow.device)
        gdata = torch.empty((n, c, h, w) if need_gdata else (0,), dtype=torch.float32, device=flow.device)
        gcx = ctypes.c_void_p(gflow.data_ptr()) if need_gflow else None
        gcy = ctypes.c_void_p(gflow.data_ptr() + 4 * hw) if need_gflow else None
        g_ext = g_ext.contiguous()             # bound to a name: the pointer must outlive the launch
        _call("ofb_splat_s_band_bwd_gather", _p(flow[:, 0]), _p(flow[:, 1]), 2 * hw, float(sign), _p(data), c * hw,
              _p(flow_mask), _bs(flow_mask, n, hw), int(occlude_zero), _p(g_ext),
              _p(unocc) if (unocc is not None and unocc.numel()) else None, _p(gout), gcx, gcy, 2 * hw,
              _p(gdata) if need_gdata else None, n, c, h, w, (ctypes.c_int * 4)(*[int(v) for v in band]), _stream())
    return gflow, gdata


def splat_xy(x, y, data, mask=None):
    """S with absolute coordinates (grid_from_unstructured_data, utils.py:1061).  Returns (out, density)."""
    out, density, _w, _v = _SplatS.apply(x, y, data, None, mask, False, 1.0, False, VALID_NONE)
    return out, density


# ------------------------------------------------------------------------------------------------------------------
# callers and data formats either side of the path (csrc/ingest.cu)
# ------------------------------------------------------------------------------------------------------------------
_lib.define("flow_extent(Tensor flow, Tensor? mask, bool source_ref) -> Tensor")
_lib.define("flo_to_planar(Tensor interleaved) -> Tensor")
_lib.define("kitti_decode(Tensor png_bgr, bool want_valid) -> (Tensor, Tensor)")
_lib.define("resize_fwd(Tensor? flow, Tensor? mask, float scale_y, float scale_x) -> (Tensor, Tensor)")
_lib.define("resize_bwd(Tensor gout, float scale_y, float scale_x, int h, int w) -> Tensor")


def _flow_extent(flow, mask, source_ref):
    require_cuda(flow, "flow")
    flow, mask = _f32c(flow), _maskc(mask)
    n, _, h, w = flow.shape
    with torch.cuda.device(flow.device):
        ext = torch.empty((n, 4), dtype=torch.float32, device=flow.device)
        _call("ofb_flow_extent", _p(flow), _p(mask), int(source_ref), _p(ext), n, h, w, _stream())
    return ext


def _flo_to_planar(interleaved):
    require_cuda(interleaved, "flow data")
    x = interleaved if (interleaved.dtype == torch.float32 and interleaved.is_contiguous()) else \
        interleaved.to(torch.float32).contiguous()
    n, h, w, _ = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((n, 2, h, w), dtype=torch.float32, device=x.device)
        _call("ofb_flo_to_planar", _p(x), _p(out), n, h, w, _stream())
    return out


def _kitti_decode(png_bgr, want_valid):
    require_cuda(png_bgr, "KITTI data")
    if png_bgr.dtype not in (torch.int16, torch.uint16) or png_bgr.dim() != 4 or png_bgr.shape[-1] != 3:
        raise TypeError("oflibpytorch_b200: KITTI data must be an (N,H,W,3) 16-bit tensor, got %s %s"
                        % (png_bgr.dtype, tuple(png_bgr.shape)))
    x = png_bgr.contiguous()
    n, h, w, _ = x.shape
    with torch.cuda.device(x.device):
        flow = torch.empty((n, 2, h, w), dtype=torch.float32, device=x.device)
        valid = torch.empty((n, h, w) if want_valid else (0,), dtype=torch.bool, device=x.device)
        _call("ofb_kitti_decode", _p(x), _p(flow), _p(valid) if want_valid else None, n, h, w, _stream())
    return flow, valid


def _resize_fwd(flow, mask, scale_y, scale_x):
    t = flow if flow is not None else mask
    require_cuda(t, "flow")
    flow = None if flow is None else _f32c(flow)
    mask = _maskc(mask)
    n, h, w = t.shape[0], t.shape[-2], t.shape[-1]
    ho = int(_native.lib().ofb_resize_output_size(h, scale_y))
    wo = int(_native.lib().ofb_resize_output_size(w, scale_x))
    with torch.cuda.device(t.device):
        out = torch.empty((n, 2, ho, wo) if flow is not None else (0,), dtype=torch.float32, device=t.device)
        out_mask = torch.empty((n, ho, wo) if mask is not None else (0,), dtype=torch.bool, device=t.device)
        _call("ofb_resize_fwd", _p(flow), _p(mask), scale_y, scale_x, _p(out) if flow is not None else None,
              _p(out_mask) if mask is not None else None, n, h, w, _stream())
    return out, out_mask


def _resize_bwd(gout, scale_y, scale_x, h, w):
    require_cuda(gout, "gradient")
    gout = _f32c(gout)
    n = gout.shape[0]
    with torch.cuda.device(gout.device):
        gin = torch.empty((n, 2, h, w), dtype=torch.float32, device=gout.device)
        _call("ofb_resize_bwd", _p(gout), scale_y, scale_x, _p(gin), n, h, w, _stream())
    return gin


_lib.impl("flow_extent", _flow_extent, "CUDA")
_lib.impl("flo_to_planar", _flo_to_planar, "CUDA")
_lib.impl("kitti_decode", _kitti_decode, "CUDA")
_lib.impl("resize_fwd", _resize_fwd, "CUDA")
_lib.impl("resize_bwd", _resize_bwd, "CUDA")


class _ResizeFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, scale_y, scale_x):
        ctx.geom = (scale_y, scale_x, flow.shape[-2], flow.shape[-1])
        return _ops.resize_fwd(flow, None, scale_y, scale_x)[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        sy, sx, h, w = ctx.geom
        return _ops.resize_bwd(gout, sy, sx, h, w), None, None


def flow_extent(flow, mask=None, source_ref=False):
    """(N,4) = [min_y, max_y, min_x, max_x] of the masked flow end points (flow_class.py:1195-1216)."""
    return _ops.flow_extent(flow.detach(), mask, bool(source_ref))


def flo_to_planar(interleaved):
    """(N,H,W,2) .flo payload on the device -> (N,2,H,W)."""
    return _ops.flo_to_planar(interleaved)


def kitti_decode(png_bgr, want_valid=True):
    """(N,H,W,3) uint16 KITTI png payload (cv2 channel order) on the device -> flow (N,2,H,W), valid (N,H,W)."""
    return _ops.kitti_decode(png_bgr, bool(want_valid))


def resize_flow(flow, scale_y, scale_x):
    """Differentiable bilinear resize of (N,2,H,W) with rescaled vectors (utils.py:906-909)."""
    return _ResizeFlow.apply(flow, float(scale_y), float(scale_x))


def resize_mask(mask, scale_y, scale_x):
    """round(bilinear resize of the 0/1 mask) as bool (flow_class.py:707-712)."""
    return _ops.resize_fwd(None, mask, float(scale_y), float(scale_x))[1]
