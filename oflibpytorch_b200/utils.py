"""Functional layer of the hot path: same names, arguments, return conventions and error behaviour as the
reference's ``oflibpytorch/utils.py``, served by the sm_100a kernels.

Reference map (the reference's ``src/oflibpytorch/utils.py``):

=============================  ========================  =============================================
here                           reference                 what changed
=============================  ========================  =============================================
get_valid_* / validation       utils.py:72-232           same checks and messages; the finite test and
                                                         the zero tests are ONE kernel pass + ONE sync
apply_flow                     utils.py:469-620          't': one fused gather kernel; 's': splat kernels
get_flow_endpoints             utils.py:1045-1058        kept for API parity (the kernels fuse it)
grid_from_unstructured_data    utils.py:1061-1154        accumulate + normalise kernels, no index tensors
apply_s_flow                   utils.py:1157-1205        occlusion rule and fill fused into the splat
threshold_vectors/is_zero_flow utils.py:623-643,919-938  is_zero_flow is a flag reduction
=============================  ========================  =============================================

CUDA tensors only.  CPU inputs are validated like the reference validates them (same exceptions) and then
rejected with a RuntimeError: there is deliberately no CPU path in this package.
"""
from typing import Any, List, Optional, Union

import numpy as np
import torch

from . import _ops
from ._ops import VALID_EQ1, VALID_GT, VALID_NONE  # noqa: F401

DEFAULT_THRESHOLD = 1e-3
PURE_PYTORCH = True

STAT_NONFINITE, STAT_NONZERO, STAT_NONZERO_T, STAT_NONZERO_M, STAT_NONZERO_TM = 1, 2, 4, 8, 16


# ----------------------------------------------------------------------------------------------------
# toolbox-wide switch (reference utils.py:24-69).  Only the differentiable splat mode exists here.
# ----------------------------------------------------------------------------------------------------

def get_pure_pytorch() -> bool:
    return PURE_PYTORCH


def set_pure_pytorch(warn: bool = None):
    global PURE_PYTORCH
    PURE_PYTORCH = True
    if warn:
        print("Pure Pytorch mode set: no use of scipy.interpolate.griddata. Differentiable, significantly faster, "
              "but more approximate")


def unset_pure_pytorch(warn: bool = None):
    raise NotImplementedError(
        "oflibpytorch_b200 implements the PURE_PYTORCH=True path only (the scipy.interpolate.griddata branch of the "
        "reference is CPU-only, non-differentiable and out of scope); use the reference package for it.")


# ----------------------------------------------------------------------------------------------------
# validation mode.  'eager' (default) is the reference's contract: a flow with NaN/Inf raises ValueError where it
# is handed over (utils.py:98), which costs one device->host read per user-supplied flow.  'deferred' never
# synchronises on the 't' path: the finite test and the zero-flow questions stay on the DEVICE (ofb_flow_stats_batch),
# the reference's early-outs (utils.py:497-498, flow_class.py:1729-1744) are taken by the kernels themselves
# (ofb_warp_t_fwd_ex), and a non-finite flow raises at the next `check_deferred()` instead of at construction.
# ----------------------------------------------------------------------------------------------------
_VALIDATION = 'eager'
_PENDING = []          # [(device flags, message)] of flows whose finite test has not been read back yet


def get_validation() -> str:
    return _VALIDATION


def set_validation(mode: str):
    """'eager': reference behaviour (ValueError at hand-over, one host sync per user-supplied flow);
    'deferred': no host sync on the 't' path, ValueError raised by `check_deferred()`."""
    global _VALIDATION
    if mode not in ('eager', 'deferred'):
        raise ValueError("validation mode must be 'eager' or 'deferred', got {!r}".format(mode))
    _VALIDATION = mode


class validation(object):
    """Context manager: ``with of.validation('deferred'): ...``"""

    def __init__(self, mode: str):
        self.mode, self.prev = mode, None

    def __enter__(self):
        self.prev = get_validation()
        set_validation(self.mode)
        return self

    def __exit__(self, *exc):
        set_validation(self.prev)
        return False


def set_deterministic(on: bool):
    """Deterministic splat merge: the forward 's' path accumulates in 64-bit fixed point, so its float outputs are
    bit-identical from run to run (integer addition does not care about the order of the atomics).  About 3x slower
    than the default float atomics, whose results differ in the last bits between runs (masks never do)."""
    from . import _native
    _native.check(_native.lib().ofb_set_option(4, 1 if on else 0), "ofb_set_option")


def _defer_finite_check(flags: torch.Tensor, error_string: str):
    if len(_PENDING) >= 4096:          # bounded: drop the oldest entries unchecked rather than grow without limit
        del _PENDING[:2048]
    _PENDING.append((flags, error_string))


def check_deferred():
    """Read back the finite tests deferred since the last call (ONE host sync) and raise the reference's ValueError
    (utils.py:98-99) if any flow handed over in 'deferred' mode held NaN, Inf or -Inf."""
    pending, bad = list(_PENDING), None
    del _PENDING[:]
    if pending:
        summary = torch.stack([f[-1] for f, _ in pending]).tolist()
        for (_, msg), fl in zip(pending, summary):
            if fl & STAT_NONFINITE:
                bad = msg
                break
    if bad is not None:
        raise ValueError(bad + "Input contains NaN, Inf or -Inf values")


# ----------------------------------------------------------------------------------------------------
# validation (reference utils.py:72-232) -- the error contract of the boundary
# ----------------------------------------------------------------------------------------------------

def _valid_vecs(vecs: Any, desired_shape=None, error_string: str = None, want_stats: bool = False,
                mask: torch.Tensor = None, lazy_ok: bool = False):
    """get_valid_vecs plus, optionally, the host copy of the per-item OFB_STAT_* flags obtained in the same
    pass that checks for NaN/Inf (so that later zero-flow questions cost nothing)."""
    error_string = '' if error_string is None else error_string
    if not isinstance(vecs, (np.ndarray, torch.Tensor)):
        raise TypeError(error_string + "Input is not a numpy array or a torch tensor")
    ndim = len(vecs.shape)
    if ndim != 3 and ndim != 4:
        raise ValueError(error_string + "Input has {} dimensions, should be 3 or 4".format(ndim))
    if isinstance(vecs, np.ndarray):
        vecs = torch.tensor(vecs, dtype=torch.float, device='cpu')
    if ndim == 3:
        vecs = vecs.unsqueeze(0)
    if vecs.shape[1] != 2:
        if vecs.shape[3] == 2:
            vecs = vecs.permute(0, 3, 1, 2)
        else:
            raise ValueError(error_string + "Input needs to be shape (N-)H-W-2 or (N-)2-H-W")
    stats = None
    if vecs.is_cuda:
        v32 = vecs.float().contiguous()
        if _VALIDATION == 'eager' or (want_stats and not lazy_ok):
            flags = _ops.flow_stats(v32, mask).tolist()          # the one sync of the call
            if any(f & STAT_NONFINITE for f in flags):
                raise ValueError(error_string + "Input contains NaN, Inf or -Inf values")
            stats = flags
        # 'deferred': nothing is launched here; the caller computes device flags when (and if) it needs them
        vecs = v32
    else:
        if not torch.isfinite(vecs).all():
            raise ValueError(error_string + "Input contains NaN, Inf or -Inf values")
        vecs = vecs.float()
    if desired_shape is not None:
        d = get_valid_shape(desired_shape)
        if vecs.shape[0] != d[0] or vecs.shape[2] != d[1] or vecs.shape[3] != d[2]:
            raise ValueError(error_string + "Input shape does not match the desired shape")
    return (vecs, stats) if want_stats else vecs


def get_valid_vecs(vecs: Any, desired_shape: Union[tuple, list] = None, error_string: str = None) -> torch.Tensor:
    """Checks array or tensor input and returns an N-2-H-W float tensor (reference utils.py:72-118)."""
    return _valid_vecs(vecs, desired_shape, error_string)


def get_valid_shape(shape: Any) -> tuple:
    """reference utils.py:121-131"""
    if not isinstance(shape, (list, tuple)):
        raise TypeError("Error creating flow from matrix: Dims need to be a list or a tuple")
    if len(shape) not in (2, 3):
        raise ValueError("Error creating flow from matrix: Dims need to be a list or a tuple of length 2 or 3")
    if any((item <= 0 or not isinstance(item, int)) for item in shape):
        raise ValueError("Error creating flow from matrix: Dims need to be a list or a tuple of integers above zero")
    return ((1,) + tuple(shape)) if len(shape) == 2 else tuple(shape)


def get_valid_ref(ref: Any) -> str:
    """reference utils.py:134-148"""
    if ref is None:
        return 't'
    if not isinstance(ref, str):
        raise TypeError("Error setting flow reference: Input is not a string")
    if ref not in ('s', 't'):
        raise ValueError("Error setting flow reference: Input is not 's' or 't', but {}".format(ref))
    return ref


def get_valid_mask(mask: Any, desired_shape: Union[tuple, list] = None, error_string: str = None) -> torch.Tensor:
    """reference utils.py:151-186.  A torch.bool input cannot hold anything but 0/1, so its value check is
    skipped (no kernel, no sync)."""
    error_string = '' if error_string is None else error_string
    if not isinstance(mask, (np.ndarray, torch.Tensor)):
        raise TypeError(error_string + "Input is not a numpy array or a torch tensor")
    ndim = len(mask.shape)
    if ndim != 2 and ndim != 3:
        raise ValueError(error_string + "Input has {} dimensions, should be 2 or 3".format(ndim))
    if isinstance(mask, np.ndarray):
        mask = torch.tensor(mask)
    if mask.dtype != torch.bool and bool(((mask != 0) & (mask != 1)).any()):
        raise ValueError(error_string + "Values must be 0 or 1")
    if ndim == 2:
        mask = mask.unsqueeze(0)
    if desired_shape is not None:
        if tuple(mask.shape) != get_valid_shape(desired_shape):
            raise ValueError(error_string + "Input shape does not match the desired shape")
    return mask.to(torch.bool)


def get_valid_device(device: Any) -> torch.device:
    """reference utils.py:189-212"""
    if device is None:
        device = torch.device('cpu')
    elif not isinstance(device, torch.device):
        try:
            device = torch.device(device)
        except RuntimeError:
            raise ValueError("Error setting tensor device: Input needs to be a torch.device, or valid input to "
                             "torch.device(). Instead found {}".format(device))
    if device.type == 'cuda':
        if not torch.cuda.is_available():
            raise ValueError("Error setting tensor device: Input is 'cuda', but cuda is not available")
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
    return device


def get_valid_padding(padding: Any, error_string: str = None) -> list:
    """reference utils.py:215-232"""
    error_string = '' if error_string is None else error_string
    if not isinstance(padding, (list, tuple)):
        raise TypeError(error_string + "Padding needs to be a list [top, bot, left, right]")
    if len(padding) != 4:
        raise ValueError(error_string + "Padding list needs to be a list of length 4 [top, bot, left, right]")
    if not all(isinstance(item, int) for item in padding):
        raise ValueError(error_string + "Padding list [top, bot, left, right] items need to be integers")
    if not all(item >= 0 for item in padding):
        raise ValueError(error_string + "Padding list [top, bot, left, right] items need to be 0 or larger")
    return padding


def move_axis(input_tensor: torch.Tensor, source: int, destination: int) -> torch.Tensor:
    """np.moveaxis for tensors (reference utils.py:235-252)."""
    return torch.movedim(input_tensor, source, destination)


def to_numpy(tensor: torch.Tensor, switch_channels: bool = None) -> np.ndarray:
    """reference utils.py:255-273"""
    arr = tensor.detach().cpu().numpy()
    return np.moveaxis(arr, 1, -1) if switch_channels else arr


def normalise_coords(coords: torch.Tensor, shape: Union[tuple, list]) -> torch.Tensor:
    """Pixel coordinates (x, y) -> [-1, 1] (reference utils.py:445-466).  Kept for API parity; the warp kernel
    evaluates the same three fp32 operations (x2, /(size-1), -1) in registers."""
    if len(shape) != 2:
        raise ValueError("Error normalising coords: Given shape needs to be list or tuple of length 2")
    out = coords.float() * 2
    out[..., 0] /= (shape[1] - 1)
    out[..., 1] /= (shape[0] - 1)
    out -= 1
    return out


# ----------------------------------------------------------------------------------------------------
# zero-flow helpers (reference utils.py:623-643, 919-938)
# ----------------------------------------------------------------------------------------------------

def threshold_vectors(vecs: torch.Tensor, threshold: Union[float, int] = None, use_mag: bool = None) -> torch.Tensor:
    """reference utils.py:623-643 (elementwise; differentiable).  The kernels never materialise this tensor:
    they evaluate the predicate |v| < 1e-3 per component in registers."""
    threshold = DEFAULT_THRESHOLD if threshold is None else threshold
    if use_mag:
        small = (torch.norm(vecs, dim=1, keepdim=True) < threshold).expand(-1, 2, -1, -1)
    else:
        small = (vecs < threshold) & (vecs > -threshold)
    return torch.where(small, torch.zeros((), dtype=vecs.dtype, device=vecs.device), vecs)


def _is_zero_from_stats(stats: List[int], thresholded: bool, masked: bool) -> List[bool]:
    bit = (STAT_NONZERO_TM if thresholded else STAT_NONZERO_M) if masked else \
        (STAT_NONZERO_T if thresholded else STAT_NONZERO)
    return [not (f & bit) for f in stats]


def is_zero_flow(flow: Union[np.ndarray, torch.Tensor], thresholded: bool = None) -> torch.Tensor:
    """reference utils.py:919-938: (N,) bool, True where the (thresholded) flow is zero everywhere."""
    flow, stats = _valid_vecs(flow, error_string="Error checking whether flow is zero: ", want_stats=True)
    thresholded = True if thresholded is None else thresholded
    if not isinstance(thresholded, bool):
        raise TypeError("Error checking whether flow is zero: Thresholded needs to be a boolean")
    _ops.require_cuda(flow, "flow")
    return torch.tensor(_is_zero_from_stats(stats, thresholded, False), dtype=torch.bool, device=flow.device)


# ----------------------------------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------------------------------

def get_flow_endpoints(flow: torch.Tensor, ref: str) -> tuple:
    """reference utils.py:1045-1058.  Differentiable w.r.t. the flow.  (Inside apply_flow / apply_s_flow the
    same two adds happen in registers; this function exists for callers that want the grids.)"""
    s = 1.0 if ref == 's' else -1.0
    h, w = flow.shape[-2:]
    x = s * flow[:, 0] + torch.arange(w, device=flow.device, dtype=torch.float32)[None, None, :]
    y = s * flow[:, 1] + torch.arange(h, device=flow.device, dtype=torch.float32)[None, :, None]
    return x, y


def grid_from_unstructured_data(x: torch.Tensor, y: torch.Tensor, data: torch.Tensor,
                                mask: torch.Tensor = None) -> tuple:
    """Inverse-bilinear splat of `data` located at (x, y) onto the regular grid (reference utils.py:1061-1154).
    Returns (grid_data (N,C,H,W), density (N,H,W)); both differentiable w.r.t. x, y and data."""
    _ops.require_cuda(data, "data")
    return _ops.splat_xy(x, y, data, mask)


def apply_s_flow(flow: torch.Tensor, data: torch.Tensor, mask: torch.Tensor = None,
                 occlude_zero_flow: bool = None) -> tuple:
    """Forward warp by an 's' flow (reference utils.py:1157-1205).  Returns (warped (N,C,H,W), mask (N,H,W))."""
    occlude_zero_flow = True if occlude_zero_flow is None else occlude_zero_flow
    _ops.require_cuda(flow, "flow")
    out, _density, wmask, _valid = _ops.splat_flow(flow, data.float(), None, mask, 1.0, bool(occlude_zero_flow),
                                                   VALID_NONE, want_aux=True)
    return out, wmask


def _restore_shape(result: torch.Tensor, target_dims: int) -> torch.Tensor:
    """reference utils.py:603-610"""
    if result.shape[0] == 1:
        if target_dims == 2:
            return result[0, 0]
        if target_dims == 3:
            return result[0]
    elif target_dims == 2:
        return result[:, 0]
    return result


def _restore_dtype(result: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """reference utils.py:614-618"""
    if not dtype.is_floating_point:
        result = torch.round(result)
        if dtype == torch.uint8:
            result = torch.clamp(result, 0, 255)
    return result.to(dtype)


def apply_flow(flow: Union[np.ndarray, torch.Tensor], target: torch.Tensor, ref: str,
               mask: Union[np.ndarray, torch.Tensor] = None) -> torch.Tensor:
    """Warps `target` with `flow` (reference utils.py:469-620).  Differentiable w.r.t. flow and target.

    :param flow: (2,H,W), (H,W,2), (N,2,H,W) or (N,H,W,2)
    :param target: (H,W), (C,H,W) or (N,C,H,W); batch sizes 1 and N broadcast
    :param ref: 't' or 's'
    :param mask: (H,W) or (N,H,W) flow mask, only relevant for 's' flows
    """
    ref = get_valid_ref(ref)
    # 'deferred' validation, 't' flows: finite test and zero-flow early-out stay on the device (no host sync)
    flow, stats = _valid_vecs(flow, error_string="Error applying flow to a target: ", want_stats=True,
                              lazy_ok=(ref == 't'))
    if stats is not None and all(_is_zero_from_stats(stats, True, False)):       # utils.py:497-498
        return target
    if not isinstance(target, torch.Tensor):
        raise TypeError("Error applying flow to a target: Target needs to be a torch tensor")
    if len(target.shape) not in [2, 3, 4]:
        raise ValueError("Error applying flow to a target: Target tensor needs to have shape H-W, C-H-W, or N-C-H-W")
    if target.shape[-2:] != flow.shape[-2:]:
        raise ValueError("Error applying flow to a target: Target height and width needs to match flow field array")
    if mask is not None:
        mask = get_valid_mask(mask, desired_shape=(flow.shape[0],) + tuple(flow.shape[2:]))
    _ops.require_cuda(flow, "flow")

    target_dtype = target.dtype
    t = target.to(torch.float)
    if t.device != flow.device:
        t = t.to(flow.device)
    if mask is not None and mask.device != flow.device:
        mask = mask.to(flow.device)
    target_dims = t.dim()
    if target_dims == 2:
        t = t[None, None]
    elif target_dims == 3:
        t = t[None]
    if t.shape[0] != flow.shape[0] and t.shape[0] != 1 and flow.shape[0] != 1:
        raise ValueError("Error applying flow to target: Batch dimensions for flow ({}) and target ({}) don't match"
                         .format(flow.shape[0], t.shape[0]))

    if ref == 't' and stats is None:
        # the kernel takes the early-out itself (the result then equals `target`, broadcast to the flow's batch)
        fl = _ops.flow_stats_batch(flow, None)
        _defer_finite_check(fl, "Error applying flow to a target: ")
        result, _ = _ops.warp_t(flow, t, fflags=fl, early=_ops.early_rules(_ops.early_rule('f', STAT_NONZERO_T, 1)))
    elif ref == 't':
        result, _ = _ops.warp_t(flow, t)                                          # utils.py:541-555
    else:
        result, _d, _w, _v = _ops.splat_flow(flow, t, None, mask, 1.0, True, VALID_NONE)   # utils.py:576
    return _restore_dtype(_restore_shape(result, target_dims), target_dtype)


def track_pts(flow: Union[np.ndarray, torch.Tensor], ref: str, pts: torch.Tensor, int_out: bool = None) -> torch.Tensor:
    """Warp points with a flow field (reference utils.py:941-1042, PURE_PYTORCH path).  Differentiable w.r.t. the flow
    and float points.  `pts` is (M,2) or (N,M,2) in (y, x) order.  SURVEY.md section 8(f2)."""
    flow, stats = _valid_vecs(flow, error_string="Error tracking points: ", want_stats=True)
    ref = get_valid_ref(ref)
    if not isinstance(pts, torch.Tensor):
        raise TypeError("Error tracking points: Pts needs to be a numpy array or a torch tensor")
    return_2d = False
    if pts.dim() == 2:
        return_2d = True
        pts = pts.unsqueeze(0).expand(flow.shape[0], -1, -1)
    elif pts.dim() == 3:
        if pts.shape[0] != flow.shape[0]:
            if pts.shape[0] == 1:
                pts = pts.expand(flow.shape[0], -1, -1)
            else:
                raise ValueError("Error tracking points: "
                                 "If used, pts batch size needs to be equal to the flow batch size")
    else:
        raise ValueError("Error tracking points: Pts needs to have shape M-2 or N-M-2")
    if pts.shape[-1] != 2:
        raise ValueError("Error tracking points: Pts needs to have shape M-2 or N-M-2")
    int_out = False if int_out is None else int_out
    if not isinstance(int_out, bool):
        raise TypeError("Error tracking points: Int_out needs to be a boolean")
    _ops.require_cuda(flow, "flow")
    pts = pts.to(flow.device)

    if all(_is_zero_from_stats(stats, True, False)):                           # utils.py:993
        warped_pts = pts
    else:
        if ref == 't':
            # utils.py:999-1001: bring the 't' flow to the source grid with one splat whose payload is the flow itself
            flow, _d, _w, _v = _ops.splat_flow(flow, flow, None, None, -1.0, False, VALID_NONE)
        if not pts.dtype.is_floating_point:                                    # utils.py:1003-1008
            n, _, h, w = flow.shape
            idx = (pts[..., 0] * w + pts[..., 1]).long().unsqueeze(-1).expand(-1, -1, 2)
            flow_vecs = torch.gather(flow.permute(0, 2, 3, 1).reshape(n, h * w, 2), 1, idx).flip(-1)
        else:                                                                  # utils.py:1010-1018
            flow_vecs = _ops.sample_pts(flow, pts.to(torch.float))
        warped_pts = pts.float() + flow_vecs
        warped_pts = torch.where(torch.isnan(warped_pts).any(-1, keepdim=True), torch.zeros_like(warped_pts), warped_pts)
    if int_out:
        warped_pts = torch.round(warped_pts).long()
    if return_2d:
        warped_pts = warped_pts.squeeze(0)
    return warped_pts


# ------------------------------------------------------------------------------------------------------------------
# ingest and resize: the data formats either side of the warp/splat path (reference utils.py:810-916)
# ------------------------------------------------------------------------------------------------------------------
def _read_flo_header(path: str):
    """12-byte Sintel header: tag 'PIEH', width, height as little-endian int32 (utils.py:844-852)."""
    if not isinstance(path, str):
        raise TypeError("Error loading flow from Sintel data: Path needs to be a string")
    with open(path, 'rb') as file:
        if file.read(4) != b'PIEH':
            raise ValueError("Error loading flow from Sintel data: Path not a valid .flo file")
        w, h = int.from_bytes(file.read(4), 'little'), int.from_bytes(file.read(4), 'little')
    if w < 1 or h < 1:
        raise ValueError("Error loading flow from Sintel data: Invalid size read from file ('{}', '{}')".format(w, h))
    return h, w


def _ingest_device(device: Any) -> torch.device:
    """the loaders of this package deliver to the current CUDA device unless told otherwise"""
    return torch.device('cuda', torch.cuda.current_device()) if device is None else get_valid_device(device)


def load_sintel_batch(paths: List[str], device: Any = None) -> torch.Tensor:
    """B200-native batched form of :func:`load_sintel`: N ``.flo`` files of one size -> (N,2,H,W) fp32 on `device`.

    The payload of every file is read straight into ONE pinned host buffer (no numpy copy, no host-side axis
    swap), moved with one host-to-device copy and de-interleaved on the device (``ofb_flo_to_planar``).
    """
    if not isinstance(paths, (list, tuple)) or len(paths) == 0:
        raise TypeError("Error loading flow from Sintel data: Paths needs to be a non-empty list of strings")
    sizes = [_read_flo_header(p) for p in paths]
    if any(s != sizes[0] for s in sizes):
        raise ValueError("Error loading flow from Sintel data: All files of a batch need to have the same size")
    h, w = sizes[0]
    device = _ingest_device(device)
    staging = torch.empty((len(paths), h, w, 2), dtype=torch.float32).pin_memory()
    raw = staging.numpy()
    for i, p in enumerate(paths):
        with open(p, 'rb') as file:
            file.seek(12)
            got = file.readinto(memoryview(raw[i]).cast('B'))
        if got != h * w * 8:
            raise ValueError("Error loading flow from Sintel data: File '{}' is shorter than its header says".format(p))
    if np.dtype('<f4') != np.dtype('=f4'):                     # the file format is little-endian
        raw.byteswap(inplace=True)
    return _ops.flo_to_planar(staging.to(device, non_blocking=True))


def load_sintel(path: str, device: Any = None) -> torch.Tensor:
    """(2,H,W) flow from a Sintel ``.flo`` file (reference utils.py:832-855), on the current CUDA device unless
    `device` says otherwise."""
    return load_sintel_batch([path], device)[0]


def load_sintel_mask(path: str, device: Any = None) -> torch.Tensor:
    """(H,W) bool mask of the VALID pixels from a Sintel invalid-pixel ``.png`` (reference utils.py:858-875).
    The png decode is host work (cv2, as in the reference); the tensor lands on the CUDA device."""
    import cv2
    if not isinstance(path, str):
        raise TypeError("Error loading flow from Sintel data: Path needs to be a string")
    mask = cv2.imread(path, 0)
    if mask is None:
        raise ValueError("Error loading flow from Sintel data: Invalid mask could not be loaded from path")
    return torch.from_numpy(mask).to(_ingest_device(device)) == 0


def load_kitti_batch(paths: List[str], device: Any = None, load_valid: bool = True):
    """B200-native batched form of :func:`load_kitti`: N KITTI flow pngs of one size -> flow (N,2,H,W) fp32 and the
    valid mask (N,H,W) bool on `device`.  The png inflate is host work (cv2, as in the reference); the uint16
    payloads go up in one copy and the channel flip, the ``(u16 - 2**15) / 64`` conversion and the valid test run in
    one kernel (``ofb_kitti_decode``)."""
    import cv2
    if not isinstance(paths, (list, tuple)) or len(paths) == 0:
        raise TypeError("Error loading flow from KITTI data: Paths needs to be a non-empty list of strings")
    decoded = []
    for p in paths:
        inp = cv2.imread(p, cv2.IMREAD_UNCHANGED) if isinstance(p, str) else None
        if inp is None:
            raise ValueError("Error loading flow from KITTI data: Flow data could not be loaded")
        if inp.ndim != 3 or inp.shape[-1] != 3 or inp.dtype != np.uint16:
            raise ValueError("Error loading flow from KITTI data: Loaded flow data has the wrong shape")
        if inp.shape != decoded[0].shape if decoded else False:
            raise ValueError("Error loading flow from KITTI data: All files of a batch need to have the same size")
        decoded.append(inp)
    device = _ingest_device(device)
    staging = torch.empty((len(paths),) + decoded[0].shape, dtype=torch.int16).pin_memory()
    for i, inp in enumerate(decoded):
        staging[i].numpy().view(np.uint16)[...] = inp
    flow, valid = _ops.kitti_decode(staging.to(device, non_blocking=True), load_valid)
    return (flow, valid) if load_valid else (flow, None)


def load_kitti(path: str, device: Any = None) -> torch.Tensor:
    """(3,H,W) fp32 tensor: flow x, flow y, valid (0/1) from a KITTI uint16 ``.png`` (reference utils.py:811-829)."""
    flow, valid = load_kitti_batch([path], device, True)
    return torch.cat((flow[0], valid[0].to(torch.float32).unsqueeze(0)), dim=0)


def _valid_scale(scale) -> list:
    """reference utils.py:893-903"""
    if isinstance(scale, (float, int)):
        scale = [scale, scale]
    elif isinstance(scale, (tuple, list)):
        if len(scale) != 2:
            raise ValueError("Error resizing flow: Scale {} must have a length of 2".format(type(scale)))
        if not all(isinstance(item, (float, int)) for item in scale):
            raise ValueError("Error resizing flow: Scale {} items must be integers or floats".format(type(scale)))
    else:
        raise TypeError("Error resizing flow: "
                        "Scale must be an integer, float, or list or tuple of integers or floats")
    if any(s <= 0 for s in scale):
        raise ValueError("Error resizing flow: Scale values must be larger than 0")
    return [float(scale[0]), float(scale[1])]


def resize_flow(flow: Union[np.ndarray, torch.Tensor], scale: Union[float, int, list, tuple]) -> torch.Tensor:
    """Bilinear resize of a flow field with the vectors rescaled accordingly; differentiable with respect to the
    input flow.  Same arguments, errors and results as the reference (utils.py:878-916): `scale` is a number or
    ``[vertical, horizontal]``; the output has ``floor(H * sv) x floor(W * sh)`` pixels."""
    valid_flow = get_valid_vecs(flow, error_string="Error resizing flow: ")
    scale = _valid_scale(scale)
    resized = _ops.resize_flow(valid_flow, scale[0], scale[1])
    if len(flow.shape) == 3:
        resized = resized.squeeze(0)
    return resized
