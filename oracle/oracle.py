"""CPU oracle for the oflibpytorch dense flow-field hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``oflibpytorch_b200`` never does (tests/test_capi_symbols.py::test_product_never_imports_the_oracle enforces it).

Parity status: PINNED against the unmodified reference (tests/test_oracle_pinned.py,
golden vectors in tests/golden/ made by tests/golden/make_golden.py) and against the
reference's own known-answer tests.

Two layers:

* the primitives W (backward warp) and S (forward splat) and their adjoints, implemented in
  plain C (oracle/oracle.c, loaded with ctypes) and wrapped in ``torch.autograd.Function`` on CPU
  tensors so composite gradients come from torch's autograd over oracle primitives;
* a compact restatement of the reference's compositions on top of them (``OFlow``), each method
  citing the reference lines it follows.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_F = ctypes.c_void_p
_I = ctypes.c_int


def build(force: bool = False) -> str:
    """Compile oracle.c with the Makefile next to it (idempotent)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_warp_t_fwd.argtypes = [_F, ctypes.c_float, _F, _F, _I, _I, _I, _I]
        _lib.orc_warp_t_bwd.argtypes = [_F, ctypes.c_float, _F, _F, _F, _F, _I, _I, _I, _I]
        _lib.orc_grid_from_unstructured.argtypes = [_F, _F, _F, _F, _F, _F, _I, _I, _I, _I]
        _lib.orc_grid_from_unstructured_bwd.argtypes = [_F] * 10 + [_I] * 4
        _lib.orc_flow_endpoints.argtypes = [_F, ctypes.c_float, _F, _F, _I, _I, _I]
        _lib.orc_zero_mask.argtypes = [_F, _F, _I, _I, _I]
        _lib.orc_is_zero_flow.argtypes = [_F, _I, _F, _I, _I, _I]
        _lib.orc_sample_pts_fwd.argtypes = [_F, _F, _F, _I, _I, _I, _I]
        _lib.orc_sample_pts_bwd.argtypes = [_F, _F, _F, _F, _F, _I, _I, _I, _I]
        _lib.orc_sample_pts_fwd.restype = _lib.orc_sample_pts_bwd.restype = None
        for f in ("orc_warp_t_fwd", "orc_warp_t_bwd", "orc_grid_from_unstructured",
                  "orc_grid_from_unstructured_bwd", "orc_flow_endpoints", "orc_zero_mask",
                  "orc_is_zero_flow"):
            getattr(_lib, f).restype = None
    return _lib


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _u8(t):
    return None if t is None else t.detach().to(torch.uint8).contiguous()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------

class _WarpT(torch.autograd.Function):
    """W: out[n,c,p] = bilinear(target[n,c], p - sign*flow[n,:,p]), zero padding (utils.py:541-555)."""

    @staticmethod
    def forward(ctx, flow, target, sign):
        f, t = _f32(flow), _f32(target)
        n, c, h, w = t.shape
        out = torch.empty_like(t)
        lib().orc_warp_t_fwd(_p(f), sign, _p(t), _p(out), n, c, h, w)
        ctx.save_for_backward(f, t)
        ctx.sign = sign
        return out

    @staticmethod
    def backward(ctx, gout):
        f, t = ctx.saved_tensors
        n, c, h, w = t.shape
        g = _f32(gout)
        gflow, gtgt = torch.empty_like(f), torch.empty_like(t)
        lib().orc_warp_t_bwd(_p(f), ctx.sign, _p(t), _p(g), _p(gflow), _p(gtgt), n, c, h, w)
        return gflow, gtgt, None


class _Splat(torch.autograd.Function):
    """S core: grid_from_unstructured_data (utils.py:1061-1154)."""

    @staticmethod
    def forward(ctx, x, y, data, mask):
        x_, y_, d_ = _f32(x), _f32(y), _f32(data)
        m_ = _u8(mask)
        n, c, h, w = d_.shape
        out = torch.empty_like(d_)
        dens = torch.empty((n, h, w), dtype=torch.float32)
        lib().orc_grid_from_unstructured(_p(x_), _p(y_), _p(d_), _p(m_), _p(out), _p(dens), n, c, h, w)
        ctx.save_for_backward(x_, y_, d_)
        ctx.m = m_
        return out, dens

    @staticmethod
    def backward(ctx, gout, gdens):
        x_, y_, d_ = ctx.saved_tensors
        n, c, h, w = d_.shape
        gx, gy, gd = torch.empty_like(x_), torch.empty_like(y_), torch.empty_like(d_)
        go = _f32(gout) if gout is not None else torch.zeros_like(d_)
        gdn = _f32(gdens) if gdens is not None else None
        lib().orc_grid_from_unstructured_bwd(_p(x_), _p(y_), _p(d_), _p(ctx.m), _p(go), _p(gdn), None,
                                             _p(gx), _p(gy), _p(gd), n, c, h, w)
        return gx, gy, gd, None


def warp_t(flow: torch.Tensor, target: torch.Tensor, sign: float = 1.0) -> torch.Tensor:
    return _WarpT.apply(flow, target, float(sign))


def grid_from_unstructured_data(x, y, data, mask=None):
    return _Splat.apply(x, y, data, mask)


def get_flow_endpoints(flow: torch.Tensor, ref: str):
    """utils.py:1045-1058 (torch ops so that autograd reaches the flow)."""
    s = 1.0 if ref == 's' else -1.0
    h, w = flow.shape[-2:]
    x = s * flow[:, 0] + torch.arange(w, dtype=torch.float32)[None, None, :]
    y = s * flow[:, 1] + torch.arange(h, dtype=torch.float32)[None, :, None]
    return x, y


def zero_mask(flow: torch.Tensor) -> torch.Tensor:
    """utils.py:1191: both components strictly inside (-1e-3, 1e-3) in fp32."""
    f = _f32(flow)
    n, _, h, w = f.shape
    z = torch.empty((n, h, w), dtype=torch.uint8)
    lib().orc_zero_mask(_p(f), _p(z), n, h, w)
    return z.bool()


def is_zero_flow(flow: torch.Tensor, thresholded: bool = True) -> torch.Tensor:
    """utils.py:919-938."""
    f = _f32(flow)
    n, _, h, w = f.shape
    z = torch.empty((n,), dtype=torch.uint8)
    lib().orc_is_zero_flow(_p(f), int(bool(thresholded)), _p(z), n, h, w)
    return z.bool()


def apply_s_flow(flow, data, mask=None, occlude_zero_flow=True):
    """utils.py:1157-1205."""
    n, c, h, w = data.shape
    x, y = get_flow_endpoints(flow, 's')
    if mask is None:
        mask = torch.ones((n, h, w), dtype=torch.bool)
    zm = None
    if occlude_zero_flow:
        zm = zero_mask(flow)
        fm = mask & ~zm
    else:
        fm = mask
    out, dens = grid_from_unstructured_data(x, y, data, fm)
    warped_mask = dens > 0
    if occlude_zero_flow:
        unocc = (mask & zm & ~warped_mask).unsqueeze(1)
        out = torch.where(unocc, data.float(), out)   # utils.py:1202-1203
    return out, warped_mask


def apply_flow(flow, target, ref, mask=None):
    """utils.py:469-620, PURE_PYTORCH path.  flow (N,2,H,W) float tensor; target (H,W)/(C,H,W)/(N,C,H,W)."""
    flow = flow.float()
    if bool(is_zero_flow(flow, True).all()):         # utils.py:497-498
        return target
    dt = target.dtype
    t = target.float()
    nd = t.dim()
    if nd == 2:
        t = t[None, None]
    elif nd == 3:
        t = t[None]
    if t.shape[0] < flow.shape[0]:                    # utils.py:532-537
        t = t.expand(flow.shape[0], -1, -1, -1)
    elif flow.shape[0] < t.shape[0]:
        flow = flow.expand(t.shape[0], -1, -1, -1)
        if mask is not None:
            mask = mask.expand(t.shape[0], -1, -1)
    if ref == 't':
        res = warp_t(flow, t, 1.0)
    else:
        res, _ = apply_s_flow(flow, t, mask, True)
    if res.shape[0] == 1:                             # utils.py:603-610
        if nd == 2:
            res = res[0, 0]
        elif nd == 3:
            res = res[0]
    elif nd == 2:
        res = res[:, 0]
    if not dt.is_floating_point:                      # utils.py:614-618
        res = torch.round(res)
        if dt == torch.uint8:
            res = torch.clamp(res, 0, 255)
    return res.to(dt)


# ----------------------------------------------------------------------------------------------
# compositions (flow_class.py)
# ----------------------------------------------------------------------------------------------

class OFlow:
    """Minimal (vecs, ref, mask) triple with the hot-path methods of the reference's Flow."""

    def __init__(self, vecs, ref='t', mask=None):
        if isinstance(vecs, np.ndarray):
            vecs = torch.tensor(vecs, dtype=torch.float32)
        if vecs.dim() == 3:
            vecs = vecs[None]
        self.vecs = vecs.float()
        self.ref = ref
        n, _, h, w = self.vecs.shape
        if mask is None:
            mask = torch.ones((n, h, w), dtype=torch.bool)
        elif isinstance(mask, np.ndarray):
            mask = torch.tensor(mask)
        if mask.dim() == 2:
            mask = mask[None]
        self.mask = mask.bool()

    @property
    def shape(self):
        return (self.vecs.shape[0],) + tuple(self.vecs.shape[2:])

    def __neg__(self):
        return OFlow(self.vecs * -1.0, self.ref, self.mask)            # flow_class.py:692,549

    def __mul__(self, k):
        return OFlow(self.vecs * float(k), self.ref, self.mask)        # flow_class.py:549

    def __add__(self, o):
        return OFlow(self.vecs + o.vecs, self.ref, self.mask & o.mask)  # flow_class.py:483-488

    def __sub__(self, o):
        return OFlow(self.vecs - o.vecs, self.ref, self.mask & o.mask)  # flow_class.py:526-531

    def is_zero(self, thresholded=True):
        f = self.vecs.detach().clone()                                   # flow_class.py:1241-1244
        f[~self.mask.unsqueeze(1).expand(-1, 2, -1, -1)] = 0
        return is_zero_flow(f, thresholded)

    def pad(self, padding, mode='constant'):
        """flow_class.py:715-736."""
        pad4 = (padding[2], padding[3], padding[0], padding[1])
        v = torch.nn.functional.pad(self.vecs, pad4, mode=mode)
        m = torch.nn.functional.pad(self.mask.unsqueeze(1), pad4).squeeze(1)
        return OFlow(v, self.ref, m)

    def apply(self, target, target_mask=None, return_valid_area=False, consider_mask=True,
              padding=None, cut=True):
        """flow_class.py:755-959 (PURE_PYTORCH path)."""
        return_flow = isinstance(target, OFlow)
        r2d = r3d = False
        if return_flow:
            t, mask = target.vecs, target.mask
            ret_dtype = torch.float32
        else:
            ret_dtype = target.dtype
            t = target
            if t.dim() == 3:
                t, r3d = t[None], True
            elif t.dim() == 2:
                t, r2d = t[None, None], True
            t = t.float()
            tshape = (t.shape[0],) + tuple(t.shape[2:])
            mask = torch.ones(tshape, dtype=torch.bool) if target_mask is None else \
                (target_mask[None] if target_mask.dim() == 2 else target_mask).clone()
        th, tw = t.shape[-2:]
        if return_flow or return_valid_area:
            if self.ref == 's':
                if tuple(mask.shape[-2:]) != tuple(self.shape[-2:]):      # flow_class.py:882-893
                    sm = self.mask if self.shape[0] >= mask.shape[0] else \
                        self.mask.expand(mask.shape[0], -1, -1)
                    mask = mask.clone()
                    for i in range(mask.shape[0]):
                        win = mask[i, padding[0]:padding[0] + self.shape[1],
                                   padding[2]:padding[2] + self.shape[2]].clone()
                        mask[i] = False
                        mask[i, padding[0]:padding[0] + self.shape[1],
                             padding[2]:padding[2] + self.shape[2]] = win & sm[i]
                else:
                    mask = mask & self.mask                               # flow_class.py:895
            if mask.shape[0] != t.shape[0]:
                t = t.expand(mask.shape[0], -1, -1, -1)
            t = torch.cat((t.float(), mask.unsqueeze(1).float()), dim=1)  # flow_class.py:898
        if padding is None:
            warped = apply_flow(self.vecs, t, self.ref, self.mask if consider_mask else None)
        else:
            fl = self.pad(padding, 'constant' if self.ref == 't' else 'replicate')
            warped = apply_flow(fl.vecs, t, fl.ref, fl.mask if consider_mask else None)
        if padding is not None and cut:
            warped = warped[..., padding[0]:padding[0] + self.shape[1],
                            padding[2]:padding[2] + self.shape[2]]
        if return_flow or return_valid_area:
            mask = warped[:, -1] > 0.99999                                # flow_class.py:922
            if self.ref == 't':
                if tuple(mask.shape[1:]) != tuple(self.mask.shape[1:]):   # flow_class.py:926-932
                    tmp = mask[:, padding[0]:padding[0] + self.shape[1],
                               padding[2]:padding[2] + self.shape[2]].clone()
                    mask = torch.zeros_like(mask)
                    mask[:, padding[0]:padding[0] + self.shape[1],
                         padding[2]:padding[2] + self.shape[2]] = tmp & self.mask
                else:
                    mask = mask & self.mask
        if return_flow:
            return OFlow(warped[:, :2], target.ref, mask)
        if return_valid_area:
            warped = warped[:, :-1]
        if not ret_dtype.is_floating_point:                               # flow_class.py:943-949
            warped = torch.round(warped.float())
            if ret_dtype == torch.uint8:
                warped = torch.clamp(warped, 0, 255)
        else:
            warped = warped.to(ret_dtype)
        if r2d and warped.shape[0] == 1:
            warped = warped[0, 0]
        if r3d and warped.shape[0] == 1:
            warped = warped[0]
        return (warped, mask) if return_valid_area else warped

    def switch_ref(self, mode='valid'):
        """flow_class.py:1022-1062."""
        other = 't' if self.ref == 's' else 's'
        if mode == 'invalid' or bool(self.is_zero(thresholded=False).all()):
            return OFlow(self.vecs, other, self.mask)
        if self.ref == 's':
            r = self.apply(self)
            r.ref = 't'
            return r
        s = OFlow(self.vecs, 's', self.mask)
        return (-s).apply(s)

    def invert(self, ref=None):
        """flow_class.py:1064-1086."""
        ref = self.ref if ref is None else ref
        if self.ref == 's':
            return self.apply(-self) if ref == 's' else OFlow(-self.vecs, 't', self.mask)
        return OFlow(-self.vecs, 's', self.mask) if ref == 's' else self.invert('s').switch_ref()

    def valid_target(self, consider_mask=True):
        """flow_class.py:1088-1127."""
        if self.ref == 's':
            a = apply_flow(self.vecs, self.mask.unsqueeze(1).float(), 's',
                           self.mask if consider_mask else None).squeeze(1)
            return a == 1
        a = apply_flow(self.vecs, torch.ones((self.shape[0], 1) + self.shape[1:]), 't').squeeze(1)
        return (a > 0.9999) & self.mask

    def valid_source(self, consider_mask=True):
        """flow_class.py:1129-1172."""
        if self.ref == 's':
            a = apply_flow(-self.vecs, torch.ones((self.shape[0], 1) + self.shape[1:]), 't').squeeze(1)
            return (a > 0.9999) & self.mask
        a = apply_flow(-self.vecs, self.mask.unsqueeze(1).float(), 's',
                       self.mask if consider_mask else None).squeeze(1)
        return a == 1

    def combine_with(self, flow, mode, thresholded=False):
        """flow_class.py:1648-1810 (PURE_PYTORCH path)."""
        if bool(self.is_zero(thresholded=thresholded).all()):
            return flow
        if bool(flow.is_zero(thresholded=thresholded).all()):
            return self if mode == 3 else self.invert()
        s = self.ref == 's'
        if mode == 1:
            if s:
                inv_t = flow.invert('t')
                return flow - (inv_t + inv_t.apply(self.switch_ref())).apply(self)
            return self.invert().apply(flow - self)
        if mode == 2:
            if s:
                return self.apply(flow - self)
            return flow - flow.apply(self.invert().apply(self))
        if s:
            return self + self.invert(ref='t').apply(flow)
        return flow + flow.apply(self)


class _SamplePts(torch.autograd.Function):
    """flow vectors (y, x) interpolated at float points (utils.py:1010-1018)."""

    @staticmethod
    def forward(ctx, flow, pts):
        f, p = _f32(flow), _f32(pts)
        n, m, _ = p.shape
        out = torch.empty((n, m, 2), dtype=torch.float32)
        lib().orc_sample_pts_fwd(_p(f), _p(p), _p(out), n, m, f.shape[2], f.shape[3])
        ctx.save_for_backward(f, p)
        return out

    @staticmethod
    def backward(ctx, gout):
        f, p = ctx.saved_tensors
        n, m, _ = p.shape
        gf, gp = torch.empty_like(f), torch.empty_like(p)
        lib().orc_sample_pts_bwd(_p(f), _p(p), _p(_f32(gout)), _p(gf), _p(gp), n, m, f.shape[2], f.shape[3])
        return gf, gp


def track_pts(flow, ref, pts, int_out=False):
    """utils.py:941-1042, PURE_PYTORCH path; flow (N,2,H,W) tensor, pts (M,2) or (N,M,2) in (y, x) order."""
    flow = flow.float()
    return_2d = pts.dim() == 2
    if return_2d:
        pts = pts.unsqueeze(0).expand(flow.shape[0], -1, -1)
    elif pts.shape[0] != flow.shape[0]:
        pts = pts.expand(flow.shape[0], -1, -1)
    if bool(is_zero_flow(flow, True).all()):
        warped = pts
    else:
        if ref == 't':
            x, y = get_flow_endpoints(-flow, 's')
            flow, _ = grid_from_unstructured_data(x, y, flow)
        if not pts.dtype.is_floating_point:
            n, _, h, w = flow.shape
            idx = (pts[..., 0] * w + pts[..., 1]).long().unsqueeze(-1).expand(-1, -1, 2)
            vec = torch.gather(flow.permute(0, 2, 3, 1).reshape(n, h * w, 2), 1, idx).flip(-1)
        else:
            vec = _SamplePts.apply(flow, pts.float().contiguous())
        warped = pts.float() + vec
        warped = torch.where(torch.isnan(warped).any(-1, keepdim=True), torch.zeros_like(warped), warped)
    if int_out:
        warped = torch.round(warped).long()
    return warped.squeeze(0) if return_2d else warped


def flow_track(fl: "OFlow", pts, int_out=False, get_valid_status=False):
    """flow_class.py:961-1020"""
    input_2d = pts.dim() == 2
    warped = track_pts(fl.vecs, fl.ref, pts, int_out)
    if input_2d and fl.shape[0] == 1:
        warped = warped.squeeze(0)
    if not get_valid_status:
        return warped
    if pts.dtype.is_floating_point:
        pts = torch.round(pts)
    pts2 = pts.unsqueeze(0) if input_2d else pts
    if pts2.shape[0] != fl.shape[0]:
        pts2 = pts2.expand(fl.shape[0], -1, -1)
    vs = fl.valid_source().view(fl.shape[0], -1)
    status = torch.gather(vs, 1, (pts2[..., 0] * fl.shape[-1] + pts2[..., 1]).long())
    return warped, (status.squeeze(0) if warped.dim() == 2 else status)


def combine(a: OFlow, b: OFlow, mode: int, ref=None) -> OFlow:
    """flow_class.py:1812-1939, restated with explicit time frames (0, 1, 2 = the three images)."""
    ref = a.ref if ref is None else ref
    span = {1: (0, 1), 2: (1, 2), 3: (0, 2)}
    sign = [[0, 1, 1], [-1, 0, 1], [-1, -1, 0]]
    ia, ib = {1: (2, 3), 2: (1, 3), 3: (1, 2)}[mode]
    s_time, t_time = span[mode]
    goal = s_time if ref == 's' else t_time
    rest = ({0, 1, 2} - {s_time, t_time}).pop()
    anchor = lambda f, i: span[i][0] if f.ref == 's' else span[i][1]   # noqa: E731
    if (mode == 1 and goal == s_time) or (mode in (2, 3) and goal == t_time):
        close, ic, far, jf = b, ib, a, ia
    else:
        close, ic, far, jf = a, ia, b, ib
    cf, ff = anchor(close, ic), anchor(far, jf)
    if ff in (s_time, t_time):
        far = far.switch_ref()
    if cf == goal:
        far = close.apply(far) if sign[rest][goal] == 1 else close.invert('t').apply(far)
    if goal == t_time:
        res = far * sign[s_time][rest] + close * sign[rest][t_time]
    else:
        res = close * sign[s_time][rest] + far * sign[rest][t_time]
    if cf != goal:
        res = close.apply(res) if sign[rest][goal] == 1 else close.invert('s').apply(res)
    res.ref = ref
    return res


# ----------------------------------------------------------------------------------------------
# raw (no-autograd) entry points used as the CPU baseline in bench.py
# ----------------------------------------------------------------------------------------------

def raw_warp_t_fwd(flow, target, sign=1.0):
    f, t = _f32(flow), _f32(target)
    n, c, h, w = t.shape
    out = torch.empty_like(t)
    lib().orc_warp_t_fwd(_p(f), sign, _p(t), _p(out), n, c, h, w)
    return out


def raw_warp_t_bwd(flow, target, gout, sign=1.0):
    f, t, g = _f32(flow), _f32(target), _f32(gout)
    n, c, h, w = t.shape
    gflow, gtgt = torch.empty_like(f), torch.empty_like(t)
    lib().orc_warp_t_bwd(_p(f), sign, _p(t), _p(g), _p(gflow), _p(gtgt), n, c, h, w)
    return gflow, gtgt


# ----------------------------------------------------------------------------------------------------------------------
# callers and data formats either side of the path: padding, resize, ingest
# ----------------------------------------------------------------------------------------------------------------------
def get_padding(fl: "OFlow"):
    """reference flow_class.py:1174-1224, one item at a time like the reference (fp32 tensor ops, then ceil)."""
    v = fl.vecs.clone()
    v[(fl.vecs < 1e-3) & (fl.vecs > -1e-3)] = 0                       # threshold_vectors, utils.py:642
    if fl.ref == 's':
        v = v * -1
    n, _, h, w = v.shape
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    v[:, 0] -= gx
    v[:, 1] -= gy
    v *= -1
    out = []
    for i in range(n):
        m = fl.mask[i]
        pad = [max(-torch.min(v[i, 1, m]), 0), max(torch.max(v[i, 1, m]) - (h - 1), 0),
               max(-torch.min(v[i, 0, m]), 0), max(torch.max(v[i, 0, m]) - (w - 1), 0)]
        out.append([int(np.ceil(float(p))) for p in pad])
    return out


def _fma32(a, b, c):
    """fp32 fused multiply-add emulated in fp64 (the product of two fp32 numbers is exact in fp64)."""
    return (np.asarray(a, np.float32).astype(np.float64) * np.asarray(b, np.float32).astype(np.float64) +
            np.asarray(c, np.float32).astype(np.float64)).astype(np.float32)


def _resize_geom(n_out: int, scale: float, n_in: int):
    """ATen upsample_bilinear2d, align_corners=False, scale factor given: src = fma(rs, dst + 0.5, -0.5) clamped at
    0 with rs = float(1 / scale) (UpSample.h area_pixel_compute_source_index / guard_index_and_lambda; the FMA is the
    CPU build's contraction, pinned by bit comparison against torch in tests/golden/make_ingest.py)."""
    f32 = np.float32
    d = np.arange(n_out, dtype=f32)
    src = np.maximum(_fma32(f32(1.0 / scale), d + f32(0.5), f32(-0.5)), f32(0)).astype(f32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = np.clip((src - i0.astype(f32)).astype(f32), f32(0), f32(1)).astype(f32)
    return i0, i1, (f32(1) - l1).astype(f32), l1


def resize_bilinear(a: np.ndarray, sy: float, sx: float) -> np.ndarray:
    """(N,C,H,W) fp32 -> (N,C,floor(H*sy),floor(W*sx)) as
    torch.nn.functional.interpolate(a, scale_factor=[sy, sx], mode='bilinear', align_corners=False) computes it on
    CPU: bit-identical in the vectorised body of ATen's kernel, within an ulp or two in its scalar tail (which
    contracts differently); the rounded masks of Flow.resize are bit-identical in every pinned case."""
    a = np.asarray(a, np.float32)
    h, w = a.shape[-2:]
    ho, wo = int(np.floor(h * sy)), int(np.floor(w * sx))
    y0, y1, h0, h1 = _resize_geom(ho, sy, h)
    x0, x1, w0, w1 = _resize_geom(wo, sx, w)
    top = _fma32(w0, a[:, :, y0][:, :, :, x0], (w1 * a[:, :, y0][:, :, :, x1]).astype(np.float32))
    bot = _fma32(w0, a[:, :, y1][:, :, :, x0], (w1 * a[:, :, y1][:, :, :, x1]).astype(np.float32))
    return _fma32(h0[:, None], top, (h1[:, None] * bot).astype(np.float32))


def resize_flow(flow: torch.Tensor, scale) -> torch.Tensor:
    """reference utils.py:878-916 (forward values)"""
    s = [scale, scale] if isinstance(scale, (int, float)) else list(scale)
    three = flow.dim() == 3
    v = flow.unsqueeze(0) if three else flow
    r = resize_bilinear(v.detach().numpy(), float(s[0]), float(s[1]))
    r[:, 0] *= np.float32(s[1])
    r[:, 1] *= np.float32(s[0])
    r = torch.from_numpy(r)
    return r.squeeze(0) if three else r


def resize_flow_bwd(gout: torch.Tensor, scale, h: int, w: int) -> torch.Tensor:
    """adjoint of resize_flow with respect to the (N,2,h,w) input flow"""
    s = [scale, scale] if isinstance(scale, (int, float)) else list(scale)
    g = gout.detach().numpy().astype(np.float64).copy()
    g[:, 0] *= np.float32(s[1])
    g[:, 1] *= np.float32(s[0])
    n, c, ho, wo = g.shape
    y0, y1, h0, h1 = _resize_geom(ho, float(s[0]), h)
    x0, x1, w0, w1 = _resize_geom(wo, float(s[1]), w)
    gin = np.zeros((n, c, h, w), np.float64)
    for ys, wy in ((y0, h0), (y1, h1)):
        for xs, wx in ((x0, w0), (x1, w1)):
            wgt = wy[:, None].astype(np.float64) * wx[None, :].astype(np.float64)
            np.add.at(gin, (slice(None), slice(None), ys[:, None], xs[None, :]), g * wgt)
    return torch.from_numpy(gin.astype(np.float32))


def resize_mask(mask: torch.Tensor, scale) -> torch.Tensor:
    """reference flow_class.py:707-712: round(bilinear(mask.float())), torch.round = half to even"""
    s = [scale, scale] if isinstance(scale, (int, float)) else list(scale)
    r = resize_bilinear(mask.to(torch.float32).unsqueeze(1).numpy(), float(s[0]), float(s[1]))
    return torch.from_numpy(np.rint(r)[:, 0] != 0)


def load_sintel(path: str) -> torch.Tensor:
    """reference utils.py:832-855"""
    with open(path, 'rb') as fh:
        if fh.read(4) != b'PIEH':
            raise ValueError("not a .flo file")
        w, h = int.from_bytes(fh.read(4), 'little'), int.from_bytes(fh.read(4), 'little')
        flow = np.frombuffer(fh.read(), dtype='<f4').reshape(h, w, 2)
    return torch.from_numpy(np.ascontiguousarray(np.moveaxis(flow, -1, 0)).astype(np.float32))


def decode_kitti(png_bgr: np.ndarray) -> torch.Tensor:
    """reference utils.py:824-829 applied to the decoded (H,W,3) uint16 array (cv2 channel order)"""
    inp = png_bgr[..., ::-1].astype('float64')
    inp[..., :2] = (inp[..., :2] - 2 ** 15) / 64
    inp[inp[..., 2] > 0, 2] = 1
    return torch.from_numpy(np.ascontiguousarray(np.moveaxis(inp, -1, 0)).astype(np.float32))
