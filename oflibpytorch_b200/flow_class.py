"""The ``Flow`` container and its hot-path methods, API-compatible with the reference's
``oflibpytorch/flow_class.py`` for: construction and properties (flow_class.py:37-236), arithmetic
(:450-692), ``pad``/``unpad`` (:715-753), ``apply`` (:755-959), ``switch_ref`` (:1022-1062), ``invert``
(:1064-1086), ``valid_target``/``valid_source`` (:1088-1172), ``is_zero`` (:1226-1244), ``combine_with``
(:1648-1810) and ``combine`` (:1812-1939).

What is different from the reference, by design:

* every hot-path method is a fixed, short sequence of kernel launches: the payload
  is never concatenated with a float copy of the mask, masks are read and written as bytes by the same kernel
  that warps the payload, `-vecs` is a kernel argument (sign) rather than a tensor, and `combine_with` mode 3 is
  ONE kernel;
* a Flow validates its vectors once (one reduction kernel, one host sync) and remembers the per-item
  finite/zero flags; results of hot-path methods are built without re-validation (the reference re-runs
  ``isfinite().all()`` and a mask value scan on every returned Flow, flow_class.py:92,180);
* CUDA tensors only for the hot-path methods (no CPU fallback); a Flow may hold CPU tensors as a container.
"""
import warnings
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as f

from . import _ops
from ._ops import VALID_EQ1, VALID_GT, VALID_NONE
from . import utils as _utils
from .utils import (STAT_NONFINITE, STAT_NONZERO, STAT_NONZERO_M, STAT_NONZERO_T, STAT_NONZERO_TM,
                    _is_zero_from_stats, _restore_dtype, _valid_vecs, get_pure_pytorch,
                    get_valid_device, get_valid_mask, get_valid_padding, get_valid_ref, get_valid_shape, track_pts)

FlowAlias = 'Flow'


class Flow(object):
    _vecs: torch.Tensor
    _mask: torch.Tensor
    _ref: str
    _device: torch.device

    def __init__(self, flow_vectors: Union[np.ndarray, torch.Tensor], ref: str = None,
                 mask: Union[np.ndarray, torch.Tensor] = None, device: Union[torch.device, int, str] = None):
        """Same arguments and validation as the reference constructor (flow_class.py:37-66).

        :param flow_vectors: (2,H,W), (N,2,H,W), (H,W,2) or (N,H,W,2) array or tensor, OpenCV convention
        :param ref: 't' (default) or 's'
        :param mask: (H,W) or (N,H,W), True where the vectors are valid; defaults to all True
        :param device: tensor device; defaults to the device of `flow_vectors`
        """
        self._stats = None
        self._dflags = None          # device copy of the flags (+ batch summary), see _get_dflags
        self._user = True            # built from caller data: owes a finite test (deferred mode)
        # If the mask is already a device-resident bool tensor of the right shape, the single validation pass
        # over the vectors (finite test, utils.py:98) also yields the mask-aware zero flags (flow_class.py:1243).
        self._mask_hint = None
        if isinstance(mask, torch.Tensor) and mask.dtype == torch.bool and mask.is_cuda and \
                isinstance(flow_vectors, torch.Tensor) and flow_vectors.device == mask.device and \
                flow_vectors.dim() == 4 and flow_vectors.shape[1] == 2 and \
                tuple(mask.shape) == (flow_vectors.shape[0],) + tuple(flow_vectors.shape[2:]):
            self._mask_hint = mask
        self.vecs = flow_vectors
        self.ref = ref
        self.mask = mask
        self.device = device
        self._mask_hint = None

    # ------------------------------------------------------------------ construction without re-validation
    @classmethod
    def _make(cls, vecs: torch.Tensor, ref: str, mask: torch.Tensor) -> FlowAlias:
        """Internal: wrap kernel outputs (already N-2-H-W fp32 / N-H-W bool on one device)."""
        obj = cls.__new__(cls)
        obj._vecs, obj._ref, obj._mask, obj._device, obj._stats = vecs, ref, mask, vecs.device, None
        obj._dflags, obj._user = None, False
        return obj

    # ------------------------------------------------------------------ properties
    @property
    def vecs(self) -> torch.Tensor:
        return self._vecs

    @vecs.setter
    def vecs(self, input_vecs):
        hint = getattr(self, '_mask_hint', None)
        v, stats = _valid_vecs(input_vecs, error_string="Error setting flow vectors: ", want_stats=True, mask=hint,
                               lazy_ok=True)
        self._vecs = v
        self._dflags, self._user = None, True
        # without a mask hint the flags of this pass are unmasked; the masked bits are only trusted once a
        # mask-aware pass ran
        self._stats = None
        self._stats_unmasked = None if hint is not None else stats
        self._stats_hinted = stats if hint is not None else None

    @property
    def vecs_numpy(self) -> np.ndarray:
        return np.moveaxis(self._vecs.detach().cpu().numpy(), 1, -1)

    @property
    def ref(self) -> str:
        return self._ref

    @ref.setter
    def ref(self, input_ref: str = None):
        self._ref = get_valid_ref(input_ref)

    @property
    def mask(self) -> torch.Tensor:
        return self._mask

    @mask.setter
    def mask(self, input_mask=None):
        if input_mask is None:
            self._mask = torch.ones(self.shape, dtype=torch.bool, device=self._vecs.device)
            if getattr(self, '_stats_unmasked', None) is not None:
                # all-true mask: masked bits == unmasked bits
                self._stats = [fl | ((fl & 2) << 2) | ((fl & 4) << 2) for fl in self._stats_unmasked]
            else:
                self._stats = None      # flags cached against the previous mask are stale: recomputed lazily
            self._dflags = None
        else:
            m = get_valid_mask(input_mask, desired_shape=self.shape, error_string="Error setting flow mask: ")
            self._mask = m.to(self._vecs.device)
            hinted = getattr(self, '_stats_hinted', None)
            self._stats = hinted if (hinted is not None and input_mask is getattr(self, '_mask_hint', None)) else None
            self._stats_hinted = None
            self._dflags = None

    @property
    def mask_numpy(self) -> np.ndarray:
        return self._mask.detach().cpu().numpy()

    @property
    def device(self) -> torch.device:
        return self._device

    @device.setter
    def device(self, input_device=None):
        device = self._vecs.device if input_device is None else get_valid_device(input_device)
        self._device = device
        self._vecs = self._vecs.to(device)
        self._mask = self._mask.to(device)

    @property
    def shape(self) -> tuple:
        return (self._vecs.shape[0],) + tuple(self._vecs.shape[2:])

    @classmethod
    def zero(cls, shape, ref: str = None, mask=None, device=None) -> FlowAlias:
        """reference flow_class.py:238-260"""
        s = get_valid_shape(shape)
        dev = get_valid_device(device)
        return cls(torch.zeros(s[0], 2, s[1], s[2], device=dev), ref, mask, dev)

    def copy(self) -> FlowAlias:
        return Flow(self._vecs, self._ref, self._mask, self._device)

    def to_device(self, device) -> FlowAlias:
        device = get_valid_device(device)
        return Flow(self._vecs.to(device), self._ref, self._mask.to(device), device)

    def __str__(self) -> str:
        return "Flow object, reference {}, batch size {}, shape {}*{}, device {}; ".format(
            self._ref, *self.shape, self._device) + self.__repr__()

    def select(self, item: int = None) -> FlowAlias:
        """reference flow_class.py:413-430"""
        if item is None:
            return self
        if not isinstance(item, int):
            raise TypeError("Error selecting from flow object: item needs to be an integer")
        try:
            return Flow(self._vecs[item], self._ref, self._mask[item], self._device)
        except IndexError:
            raise IndexError("Error selecting from flow object: item {} out of bounds for flow with batch size {}"
                             .format(item, self.shape[0]))

    def __getitem__(self, item) -> FlowAlias:
        """Spatial slicing like a (H,W,...) tensor (reference flow_class.py:432-448)."""
        vecs = self._vecs.permute(2, 3, 0, 1).__getitem__(item).permute(2, 3, 0, 1)
        mask = self._mask.permute(1, 2, 0).__getitem__(item).permute(2, 0, 1)
        return Flow(vecs, self._ref, mask, self._device)

    # ------------------------------------------------------------------ arithmetic (flow_class.py:450-692)
    def _other_operand(self, other, verb: str, noun: str, err_a: str):
        if isinstance(other, (np.ndarray, torch.Tensor)):
            ov = _valid_vecs(other, desired_shape=self.shape, error_string="Error adding to flow: ")
            om = None
        elif isinstance(other, Flow):
            ov, om = other._vecs, other._mask
        else:
            raise TypeError("Error {} flow: {} is not a flow object, numpy array, or torch tensor".format(verb, noun))
        if self.shape[0] != ov.shape[0] and self.shape[0] != 1 and ov.shape[0] != 1:
            raise ValueError(err_a)
        return ov.to(self._device), (None if om is None else om.to(self._device))

    def _broadcast_mask(self, om: Optional[torch.Tensor], n_other: int) -> torch.Tensor:
        """Mask of self (+/-) other: the AND of both masks, expanded to the broadcast batch size so that vectors and
        mask of the result always agree (the reference gets there through the Flow constructor, flow_class.py:476)."""
        m = self._mask if om is None else (self._mask & om)
        n = max(self.shape[0], n_other)
        return m if m.shape[0] == n else m.expand(n, -1, -1).contiguous()

    def __add__(self, other) -> FlowAlias:
        ov, om = self._other_operand(other, "adding to", "Addend",
                                     "Error adding to flow: Augend and addend batch dimensions don't match, "
                                     "and neither is 1")
        if self.shape[1:] != tuple(ov.shape[2:]):
            raise ValueError("Error adding to flow: Augend and addend flow objects are not the same shape")
        return Flow._make(self._vecs + ov, self._ref, self._broadcast_mask(om, ov.shape[0]))

    def __sub__(self, other) -> FlowAlias:
        ov, om = self._other_operand(other, "subtracting from", "Subtrahend",
                                     "Error subtracting from flow: Minuend and subtrahend batch dimensions don't "
                                     "match, and neither is 1")
        if self.shape[1:] != tuple(ov.shape[2:]):
            raise ValueError("Error subtracting from flow: Minuend and subtrahend flow objects are not the same shape")
        return Flow._make(self._vecs - ov, self._ref, self._broadcast_mask(om, ov.shape[0]))

    def _scalar_like(self, other, verb: str, noun: str):
        """Shared operand handling of __mul__/__truediv__/__pow__ (flow_class.py:533-678)."""
        try:
            return float(other)
        except (TypeError, ValueError):
            pass
        if isinstance(other, list):
            if len(other) != 2:
                raise ValueError("Error {} flow: {} list not length 2".format(verb, noun))
            other = torch.tensor(other)
        elif isinstance(other, np.ndarray):
            other = torch.tensor(other)
        if not isinstance(other, torch.Tensor):
            raise TypeError("Error {} flow: {} cannot be converted to float, or isn't a list, numpy array, "
                            "or torch tensor".format(verb, noun))
        hw = tuple(self.shape[1:])
        if other.dim() == 1 and other.shape[0] == 2:
            other = other.view(1, 2, 1, 1)
        elif other.dim() == 2 and tuple(other.shape) == hw:
            other = other[None, None]
        elif other.dim() == 3 and tuple(other.shape) == (2,) + hw:
            other = other[None]
        elif other.dim() == 3 and tuple(other.shape) == hw + (2,):
            other = other.permute(2, 0, 1)[None]
        elif other.dim() == 4 and tuple(other.shape[2:]) == hw and other.shape[1] == 2 and \
                (self.shape[0] == 1 or other.shape[0] == 1 or self.shape[0] == other.shape[0]):
            pass
        else:
            raise ValueError("Error {} flow: {} array or tensor needs to be of size 2, of the shape of the flow "
                             "object (H-W), or 2-H-W or H-W-2, or N-2-H-W".format(verb, noun))
        return other.to(self._device)

    def _with_vecs(self, v: torch.Tensor) -> FlowAlias:
        m = self._mask if v.shape[0] == self.shape[0] else self._mask.repeat(v.shape[0], 1, 1)
        return Flow(v, self._ref, m, self._device)

    def __mul__(self, other) -> FlowAlias:
        return self._with_vecs(self._vecs * self._scalar_like(other, "multiplying", "Multiplier"))

    def __truediv__(self, other) -> FlowAlias:
        return self._with_vecs(self._vecs / self._scalar_like(other, "dividing", "Divisor"))

    def __pow__(self, other) -> FlowAlias:
        return self._with_vecs(self._vecs ** self._scalar_like(other, "exponentiating", "Exponent"))

    def __neg__(self) -> FlowAlias:
        out = Flow._make(self._vecs * -1.0, self._ref, self._mask)
        out._stats = self._stats        # |v| tests are sign-symmetric
        return out

    # ------------------------------------------------------------------ pad / unpad (flow_class.py:715-753)
    def pad(self, padding: list = None, mode: str = None) -> FlowAlias:
        mode = 'constant' if mode is None else mode
        if mode not in ['constant', 'reflect', 'replicate']:
            raise ValueError("Error padding flow: Mode should be one of "
                             "'constant', 'reflect', or 'replicate', but instead got '{}'".format(mode))
        padding = get_valid_padding(padding, "Error padding flow: ")
        pad4 = (padding[2], padding[3], padding[0], padding[1])
        vecs = f.pad(self._vecs, pad4, mode=mode)
        mask = f.pad(self._mask.unsqueeze(1), pad4).squeeze(1)
        return Flow._make(vecs.contiguous(), self._ref, mask.contiguous())

    def unpad(self, padding: list = None) -> FlowAlias:
        padding = get_valid_padding(padding, "Error padding flow: ")
        h, w = self.shape[1:3]
        if sum(padding[0:2]) > h - 1 or sum(padding[2:4]) > w - 1:
            raise ValueError("Error unpadding flow: one or more dimensions cut to zero or less")
        return self[padding[0]:h - padding[1], padding[2]:w - padding[3]]

    # ------------------------------------------------------------------ zero tests
    # ------------------------------------------------------------------ ingest, resize, padding
    @classmethod
    def from_sintel(cls, path: str, inv_path: str = None, device: Union[torch.device, int, str] = None) -> FlowAlias:
        """Flow with reference 's' from a Sintel ``.flo`` file, optionally with the invalid-pixel ``.png`` as mask
        (reference flow_class.py:358-373).  `device` defaults to the current CUDA device."""
        from .utils import load_sintel, load_sintel_mask
        flow = load_sintel(path, device)
        mask = None if inv_path is None else load_sintel_mask(inv_path, flow.device)
        return cls(flow, 's', mask)

    @classmethod
    def from_kitti(cls, path: str, load_valid: bool = None, device: Union[torch.device, int, str] = None) -> FlowAlias:
        """Flow with reference 's' from a KITTI uint16 ``.png``, the valid pixels as mask unless `load_valid` is
        False (reference flow_class.py:331-355).  `device` defaults to the current CUDA device."""
        from .utils import load_kitti_batch
        load_valid = True if load_valid is None else load_valid
        if not isinstance(load_valid, bool):
            raise TypeError("Error loading flow from KITTI data: Load_valid needs to be boolean")
        flow, valid = load_kitti_batch([path], device, load_valid)
        return cls(flow, 's', valid)

    def resize(self, scale: Union[float, int, list, tuple]) -> FlowAlias:
        """Resized flow with rescaled vectors and the bilinearly resized, rounded mask (reference
        flow_class.py:694-712); differentiable with respect to the vectors.  Vectors and mask go through one kernel
        launch each (``ofb_resize_fwd``) instead of two interpolations, two in-place scalings and a round."""
        from .utils import _valid_scale
        s = _valid_scale(scale)
        vecs = _ops.resize_flow(self._vecs, s[0], s[1])
        mask = _ops.resize_mask(self._mask, s[0], s[1])
        return Flow(vecs, self._ref, mask)

    def get_padding(self, item: int = None) -> list:
        """[top, bottom, left, right] per batch item (or for `item`) so that every masked vector finds its source
        ('t') or keeps its destination ('s') inside the padded area (reference flow_class.py:1174-1224).  The masked
        min/max of the end points of the whole batch come from one kernel and one host read; the reference does
        four reductions with a host sync each per item."""
        flow = self.select(item=item)
        ext = _ops.flow_extent(flow._vecs, flow._mask, flow._ref == 's').cpu()
        if bool(torch.isinf(ext).any()):
            raise RuntimeError("Error determining padding: the mask of a batch item has no valid pixel")
        h, w = flow.shape[1], flow.shape[2]
        zero = torch.zeros(())
        pads = torch.stack((torch.maximum(-ext[:, 0], zero), torch.maximum(ext[:, 1] - (h - 1), zero),
                            torch.maximum(-ext[:, 2], zero), torch.maximum(ext[:, 3] - (w - 1), zero)), dim=1)
        padding = [[int(v) for v in row] for row in torch.ceil(pads).tolist()]
        return padding[0] if item is not None else padding

    def _get_dflags(self) -> torch.Tensor:
        """(N + 1,) int32 ON THE DEVICE: per-item OFB_STAT_* flags of (vecs, mask) and their OR over the batch.  One
        kernel, no host sync; cached.  Results of the fused combine kernel arrive with their flags already set."""
        if self._dflags is None:
            _ops.require_cuda(self._vecs, "flow")
            self._dflags = _ops.flow_stats_batch(self._vecs, self._mask)
            if self._user and self._stats is None:       # not validated eagerly: the finite test is owed
                _utils._defer_finite_check(self._dflags, "Error setting flow vectors: ")
        return self._dflags

    def _get_stats(self) -> List[int]:
        """Per-item OFB_STAT_* flags of (vecs, mask) on the HOST: one kernel, one sync, cached."""
        if self._stats is None:
            flags = self._get_dflags().tolist()
            if self._user and any(f & STAT_NONFINITE for f in flags):
                raise ValueError("Error setting flow vectors: Input contains NaN, Inf or -Inf values")
            self._stats = flags[:-1]
        return self._stats

    @staticmethod
    def _deferred() -> bool:
        return _utils.get_validation() == 'deferred'

    def is_zero(self, thresholded: bool = None, masked: bool = None) -> torch.Tensor:
        """reference flow_class.py:1226-1244 (+ utils.py:919-938)"""
        masked = True if masked is None else masked
        if not isinstance(masked, bool):
            raise TypeError("Error checking whether flow is zero: Masked needs to be a boolean")
        thresholded = True if thresholded is None else thresholded
        if not isinstance(thresholded, bool):
            raise TypeError("Error checking whether flow is zero: Thresholded needs to be a boolean")
        z = _is_zero_from_stats(self._get_stats(), thresholded, masked)
        return torch.tensor(z, dtype=torch.bool, device=self._device)

    def _all_zero(self, thresholded: bool, masked: bool) -> bool:
        return all(_is_zero_from_stats(self._get_stats(), thresholded, masked))

    # ------------------------------------------------------------------ the warp itself
    def _warp(self, payload: torch.Tensor, payload_mask: Optional[torch.Tensor], need_valid: bool,
              consider_mask: bool, sign: float = 1.0, flow_mask_and: bool = True
              ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """apply_flow(sign*vecs, [payload | payload_mask], ref, mask) followed by the 0.99999 threshold of
        Flow.apply, as one launch sequence.  For 't' the valid bit is ANDed with self.mask when `flow_mask_and`;
        for 's' `payload_mask` must already contain the AND with self.mask (flow_class.py:880-895).
        Includes the zero-flow early-out of apply_flow (utils.py:497-498)."""
        if self._ref == 't' and self._deferred():
            # the early-out (utils.py:497-498) is taken by the kernel from the device flags: no host sync
            out, valid = _ops.warp_t(self._vecs, payload, payload_mask,
                                     self._mask if (need_valid and flow_mask_and) else None,
                                     sign=sign, thresh=0.99999, want_valid=need_valid, fflags=self._get_dflags(),
                                     early=_ops.early_rules(_ops.early_rule('f', STAT_NONZERO_T, 1)))
            return out, (valid if need_valid else None)
        if self._all_zero(thresholded=True, masked=False):
            n = max(payload.shape[0], self.shape[0], 1 if payload_mask is None else payload_mask.shape[0])
            out = payload
            valid = None
            if need_valid:
                valid = payload_mask if payload_mask is not None else \
                    torch.ones((n,) + tuple(payload.shape[2:]), dtype=torch.bool, device=payload.device)
                if valid.shape[0] != n:
                    valid = valid.expand(n, -1, -1)
                if out.shape[0] != n:
                    out = out.expand(n, -1, -1, -1)
                if (self._ref == 't' and flow_mask_and) or (self._ref == 's' and consider_mask and flow_mask_and):
                    valid = valid & self._mask
            return out, valid
        if self._ref == 't':
            out, valid = _ops.warp_t(self._vecs, payload, payload_mask,
                                     self._mask if (need_valid and flow_mask_and) else None,
                                     sign=sign, thresh=0.99999, want_valid=need_valid)
        else:
            if consider_mask and payload_mask is self._mask:
                payload_mask = None      # every source that is used has mask channel 1: the "bad weight" plane is zero
            # (with consider_mask a None payload mask means "the flow mask itself": same shortcut)
            out, _d, _w, valid = _ops.splat_flow(self._vecs, payload, payload_mask,
                                                 self._mask if consider_mask else None, sign=sign,
                                                 occlude_zero=True, valid_mode=VALID_GT if need_valid else VALID_NONE)
        return out, (valid if need_valid else None)

    def apply(self, target, target_mask: torch.Tensor = None, return_valid_area: bool = None,
              consider_mask: bool = None, padding: list = None, cut: bool = None):
        """Apply the flow to a tensor or to another Flow (reference flow_class.py:755-959; same arguments,
        return values and exceptions).  Differentiable w.r.t. the flow vectors and a tensor target."""
        return_valid_area = False if return_valid_area is None else return_valid_area
        if not isinstance(return_valid_area, bool):
            raise TypeError("Error applying flow: Return_valid_area needs to be a boolean")
        consider_mask = True if consider_mask is None else consider_mask
        if not isinstance(consider_mask, bool):
            raise TypeError("Error applying flow: Consider_mask needs to be a boolean")
        cut = True if cut is None else cut
        if not isinstance(cut, bool):
            raise TypeError("Error applying flow: Cut needs to be a boolean")
        if padding is not None:
            padding = get_valid_padding(padding, "Error applying flow: ")
            if self.shape[1] + padding[0] + padding[1] != target.shape[-2] or \
                    self.shape[2] + padding[2] + padding[3] != target.shape[-1]:
                raise ValueError("Error applying flow: Padding values do not match flow and target shape difference")

        return_dtype = torch.float
        return_2d = return_3d = False
        if isinstance(target, Flow):
            return_flow = True
            t = target._vecs.to(self._device)
            mask = target._mask.to(self._device)
        elif isinstance(target, torch.Tensor):
            return_flow = False
            if target.dim() == 4:
                t = target
            elif target.dim() == 3:
                t, return_3d = target.unsqueeze(0), True
            elif target.dim() == 2:
                t, return_2d = target.unsqueeze(0).unsqueeze(0), True
            else:
                raise ValueError("Error applying flow: Target needs to have the shape H-W (2 dimensions)"
                                 ", C-H-W (3 dimensions), or N-C-H-W (4 dimensions)")
            t = t.to(torch.float).to(self._device)
            t_shape = (t.shape[0],) + tuple(t.shape[2:])
            mask = None                                             # all true, never materialised
            if target_mask is not None:
                if not isinstance(target_mask, torch.Tensor):
                    raise TypeError("Error applying flow: Target_mask needs to be a torch tensor")
                if target_mask.dim() == 2:
                    target_mask = target_mask.unsqueeze(0)
                if tuple(target_mask.shape) != t_shape:
                    raise ValueError("Error applying flow: Target_mask needs to match the target shape")
                if target_mask.dtype != torch.bool:
                    raise TypeError("Error applying flow: Target_mask needs to have dtype 'bool'")
                if not return_valid_area:
                    warnings.warn("Warning applying flow: a mask is passed, but return_valid_area is False - so the "
                                  "mask passed will not affect the output, but possibly make the function slower.")
                mask = target_mask.to(self._device)
            return_dtype = target.dtype
        else:
            raise TypeError("Error applying flow: Target needs to be either a flow object or a torch tensor")
        if padding is None and tuple(target.shape[-2:]) != tuple(self.shape[-2:]):
            raise ValueError("Error applying flow: Flow and target have to have the same shape")

        need_valid = return_flow or return_valid_area
        if padding is not None and cut and self._ref == 't':
            # Fused padding: the flow is a window at (top, left) of the target's frame; the kernel
            # works in that frame directly, so neither the zero-padded flow (flow_class.py:913) nor the cut
            # (flow_class.py:917-918) is materialised.
            return self._finish_apply(*self._warp_window(t, mask, need_valid, padding), target if return_flow else None,
                                      return_dtype, return_2d, return_3d, return_valid_area)
        if padding is not None and not cut and self._ref == 't':
            # Fused padding without the cut (flow_class.py:905-934): the whole target frame is warped by a flow that is
            # zero outside its window; the valid area is false out there.  No padded flow / mask is materialised.
            return self._finish_apply(*self._warp_inset(t, mask, need_valid, padding), target if return_flow else None,
                                      return_dtype, return_2d, return_3d, return_valid_area)
        if padding is not None and cut and self._ref == 's' and consider_mask and t.shape[1] <= 6:
            # Fused padding for 's' flows.  The reference pads the flow with replicated vectors and
            # a False mask (flow_class.py:905-914), splats the whole padded frame and cuts the window back out
            # (:917-918).  With consider_mask the padded sources are ignored, so only the window is splatted -- in
            # the padded FRAME (ofb_splat_s_win_fwd: end points y + top + f, x + left + f, the reference's floats),
            # dropping what lands outside the window.
            return self._finish_apply(*self._splat_window(t, mask, need_valid, padding), target if return_flow else None,
                                      return_dtype, return_2d, return_3d, return_valid_area)
        flow = self
        top = left = 0
        if padding is not None:                                     # flow_class.py:905-914
            flow = self.pad(padding, mode='constant' if self._ref == 't' else 'replicate')
            top, left = padding[0], padding[2]
        fh, fw = self.shape[1], self.shape[2]

        payload_mask = mask
        if need_valid and self._ref == 's':                         # flow_class.py:880-895
            if padding is not None:
                # target mask ANDed with the flow mask inside the flow window, False outside (functional
                # version of the reference's in-place loop; the caller's mask is left untouched)
                base = mask if mask is not None else torch.ones((t.shape[0],) + tuple(t.shape[2:]), dtype=torch.bool,
                                                                device=self._device)
                n = max(base.shape[0], self.shape[0])
                pm = torch.zeros((n,) + tuple(base.shape[1:]), dtype=torch.bool, device=self._device)
                pm[:, top:top + fh, left:left + fw] = base[:, top:top + fh, left:left + fw] & self._mask
                payload_mask = pm
            elif consider_mask:
                # The reference warps (target mask & flow mask) (flow_class.py:895).  Sources outside the flow mask carry
                # no weight when the mask is considered, so the AND is implied and never materialised; the zero-flow
                # early-out of _warp, which returns the mask channel itself, applies it.
                payload_mask = mask
            else:
                payload_mask = self._mask if mask is None else (mask & self._mask)

        uncut = padding is not None and not cut
        warped, valid = flow._warp(t, payload_mask if need_valid else None, need_valid, consider_mask,
                                   flow_mask_and=not (padding is not None))
        if padding is not None and cut:                             # flow_class.py:917-918
            warped = warped[..., top:top + fh, left:left + fw]
            if valid is not None:
                valid = valid[..., top:top + fh, left:left + fw]
        if need_valid and self._ref == 't' and padding is not None:  # flow_class.py:924-934
            if uncut:
                full = torch.zeros_like(valid)
                full[:, top:top + fh, left:left + fw] = valid[:, top:top + fh, left:left + fw] & self._mask
                valid = full
            else:
                valid = valid & self._mask

        return self._finish_apply(warped, valid, target if return_flow else None, return_dtype, return_2d, return_3d,
                                  return_valid_area)

    def _warp_window(self, t: torch.Tensor, mask: Optional[torch.Tensor], need_valid: bool, padding: list):
        """'t' warp of the (top, left) window of a larger target, with apply_flow's zero-flow early-out."""
        top, left = padding[0], padding[2]
        fh, fw = self.shape[1], self.shape[2]
        th, tw = t.shape[-2:]
        if self._all_zero(thresholded=True, masked=False):           # utils.py:497-498, then the cut
            out = t[..., top:top + fh, left:left + fw]
            valid = None
            if need_valid:
                n = max(t.shape[0], self.shape[0])
                valid = (mask[..., top:top + fh, left:left + fw] if mask is not None else
                         torch.ones((n, fh, fw), dtype=torch.bool, device=self._device)) & self._mask
            return out, valid
        out, valid = _ops.warp_t_win(self._vecs, t, [top, left, th, tw, 0, 0, th, tw], mask if need_valid else None,
                                     self._mask if need_valid else None, want_valid=need_valid)
        return out, (valid if need_valid else None)

    def _warp_inset(self, t: torch.Tensor, mask: Optional[torch.Tensor], need_valid: bool, padding: list):
        """'t' warp of the WHOLE target frame by this flow sitting at (top, left) of it, zero elsewhere
        (padding without the cut), with apply_flow's zero-flow early-out."""
        top, left = padding[0], padding[2]
        fh, fw = self.shape[1], self.shape[2]
        n = max(t.shape[0], self.shape[0])
        if self._all_zero(thresholded=True, masked=False):           # utils.py:497-498: the target comes back unchanged
            out = t if t.shape[0] == n else t.expand(n, -1, -1, -1)
            valid = None
            if need_valid:                                            # flow_class.py:924-932
                valid = torch.zeros((n,) + tuple(t.shape[2:]), dtype=torch.bool, device=self._device)
                inner = mask[..., top:top + fh, left:left + fw] if mask is not None else True
                valid[:, top:top + fh, left:left + fw] = self._mask & inner
            return out, valid
        out, valid = _ops.warp_t_inset(self._vecs, t, (top, left), mask if need_valid else None,
                                       self._mask if need_valid else None, want_valid=need_valid)
        return out, (valid if need_valid else None)

    def _splat_window(self, t: torch.Tensor, mask: Optional[torch.Tensor], need_valid: bool, padding: list):
        """'s' apply of the (top, left) window of a larger target, with apply_flow's zero-flow early-out."""
        top, left = padding[0], padding[2]
        fh, fw = self.shape[1], self.shape[2]
        th, tw = t.shape[-2:]
        tw_ = t[..., top:top + fh, left:left + fw]
        payload_mask = None
        if need_valid:                                               # flow_class.py:880-895 restricted to the window
            payload_mask = self._mask if mask is None else (mask[..., top:top + fh, left:left + fw] & self._mask)
        if self._all_zero(thresholded=True, masked=False):           # utils.py:497-498, then the cut
            n = max(t.shape[0], self.shape[0])
            valid = None
            if need_valid:
                valid = payload_mask if payload_mask.shape[0] == n else payload_mask.expand(n, -1, -1)
            return (tw_ if tw_.shape[0] == n else tw_.expand(n, -1, -1, -1)), valid
        same = payload_mask is self._mask
        out, _d, _w, valid = _ops.splat_flow(self._vecs, tw_, None if (same or not need_valid) else payload_mask,
                                             self._mask, sign=1.0, occlude_zero=True,
                                             valid_mode=VALID_GT if need_valid else VALID_NONE,
                                             window=(top, left, th, tw))
        return out, (valid if need_valid else None)

    def _finish_apply(self, warped, valid, flow_target, return_dtype, return_2d, return_3d, return_valid_area):
        """Common tail of apply: wrap / round / squeeze (reference flow_class.py:937-959)."""
        if flow_target is not None:
            return Flow._make(warped[:, :2].contiguous(), flow_target._ref, valid.contiguous())
        if not return_dtype.is_floating_point:                      # flow_class.py:943-949
            warped = torch.round(warped.float())
            if return_dtype == torch.uint8:
                warped = torch.clamp(warped, 0, 255)
            if not get_pure_pytorch():
                warped = warped.to(return_dtype)
        else:
            warped = warped.to(return_dtype)
        if return_2d and warped.shape[0] == 1:
            warped = warped.squeeze(0).squeeze(0)
        if return_3d and warped.shape[0] == 1:
            warped = warped.squeeze(0)
        if return_valid_area:
            return warped, valid
        return warped

    def track(self, pts: torch.Tensor, int_out: bool = None, get_valid_status: bool = None):
        """Warp points with the flow (reference flow_class.py:961-1020).  With `get_valid_status` also returns, per
        point, whether it is tracked by a valid vector and lands inside the field (= `valid_source` at the point)."""
        input_2d = pts.dim() == 2
        get_valid_status = False if get_valid_status is None else get_valid_status
        if not isinstance(get_valid_status, bool):
            raise TypeError("Error tracking points: Get_tracked needs to be a boolean")
        warped_pts = track_pts(flow=self._vecs, ref=self._ref, pts=pts, int_out=int_out)
        if input_2d and self.shape[0] == 1:
            warped_pts = warped_pts.squeeze(0)
        if not get_valid_status:
            return warped_pts
        if pts.dtype.is_floating_point:
            pts = torch.round(pts)
        pts2 = (pts.unsqueeze(0) if input_2d else pts).to(self._device)
        if pts2.shape[0] != self.shape[0]:
            pts2 = pts2.expand(self.shape[0], -1, -1)
        valid_source = self.valid_source().view(self.shape[0], -1)
        idx = (pts2[..., 0] * self.shape[-1] + pts2[..., 1]).long()
        status = torch.gather(valid_source, 1, idx)
        if warped_pts.dim() == 2:
            status = status.squeeze(0)
        return warped_pts, status

    # ------------------------------------------------------------------ ref switching and inversion
    def switch_ref(self, mode: str = None) -> FlowAlias:
        """reference flow_class.py:1022-1062.  'valid' mode is ONE splat whose payload is the flow itself:
        s->t: S(v, [v|m], m);  t->s: S(-v, [v|m], m)  (SURVEY.md section 3.4)."""
        mode = 'valid' if mode is None else mode
        other = 't' if self._ref == 's' else 's'
        if mode == 'invalid':
            return Flow._make(self._vecs, other, self._mask)
        if mode != 'valid':
            raise ValueError("Error switching flow reference: Mode not recognised, should be 'valid' or 'invalid'")
        if self._all_zero(thresholded=False, masked=True):
            return Flow._make(self._vecs, other, self._mask)
        sign = 1.0 if self._ref == 's' else -1.0
        return self._splat_self(self._vecs, sign, other)

    def _splat_self(self, payload: torch.Tensor, sign: float, out_ref: str) -> FlowAlias:
        """S(sign*v, [payload | m], m) with apply_flow's early-out (thresholded, unmasked; utils.py:497)."""
        if self._all_zero(thresholded=True, masked=False):
            return Flow._make(payload, out_ref, self._mask)
        # payload mask == flow mask here, so every contributing source has mask channel 1 (no "bad weight" plane)
        out, _d, _w, valid = _ops.splat_flow(self._vecs, payload, None, self._mask, sign=sign,
                                             occlude_zero=True, valid_mode=VALID_GT)
        return Flow._make(out, out_ref, valid)

    def invert(self, ref: str = None) -> FlowAlias:
        """reference flow_class.py:1064-1086"""
        ref = self._ref if ref is None else get_valid_ref(ref)
        if ref != self._ref:
            return -Flow._make(self._vecs, ref, self._mask)          # flow_class.py:1081,1084
        if self._ref == 's':
            return self._splat_self(-self._vecs, 1.0, 's')           # self.apply(-self), :1079
        # 't' -> 't': Flow(-v,'s',m).switch_ref()  (:1086) = S(-v, [-v|m], m) relabelled 't'
        if self._all_zero(thresholded=False, masked=True):
            return Flow._make(-self._vecs, 't', self._mask)
        return self._splat_self(-self._vecs, -1.0, 't')

    # ------------------------------------------------------------------ valid areas (flow_class.py:1088-1172)
    def _valid_area(self, sign: float, splat: bool, consider_mask: bool) -> torch.Tensor:
        if not splat and self._deferred():
            _o, valid = _ops.warp_t(self._vecs, None, None, self._mask, sign=sign, thresh=0.9999, want_valid=True,
                                    fflags=self._get_dflags(),
                                    early=_ops.early_rules(_ops.early_rule('f', STAT_NONZERO_T, 1)))
            return valid
        if self._all_zero(thresholded=True, masked=False):           # apply_flow's early-out returns its input
            return self._mask.clone()
        if splat:
            _o, _d, _w, valid = _ops.splat_flow(self._vecs, None, None if consider_mask else self._mask,
                                                self._mask if consider_mask else None,
                                                sign=sign, occlude_zero=True, valid_mode=VALID_EQ1)
            return valid
        _o, valid = _ops.warp_t(self._vecs, None, None, self._mask, sign=sign, thresh=0.9999, want_valid=True)
        return valid

    def valid_target(self, consider_mask: bool = None) -> torch.Tensor:
        consider_mask = True if consider_mask is None else consider_mask
        if not isinstance(consider_mask, bool):
            raise TypeError("Error applying flow: Consider_mask needs to be a boolean")
        return self._valid_area(1.0, self._ref == 's', consider_mask)

    def valid_source(self, consider_mask: bool = None) -> torch.Tensor:
        consider_mask = True if consider_mask is None else consider_mask
        if not isinstance(consider_mask, bool):
            raise TypeError("Error applying flow: Consider_mask needs to be a boolean")
        return self._valid_area(-1.0, self._ref == 't', consider_mask)

    # ------------------------------------------------------------------ composition (flow_class.py:1648-1810)
    def combine_with(self, flow: FlowAlias, mode: int, thresholded: bool = None) -> FlowAlias:
        """flow_1 (+) flow_2 = flow_3, solved for the flow named by `mode` (reference flow_class.py:1648-1810).
        Mode 3 is a single fused kernel: warp + add + mask AND."""
        if not isinstance(flow, Flow):
            raise TypeError("Error combining flows: Flow need to be of type 'Flow'")
        if self.shape != flow.shape:
            raise ValueError("Error combining flows: Flow fields need to have the same shape, including batch size")
        if self.ref != flow.ref:
            raise ValueError("Error combining flows: Flow fields need to have the same reference")
        if self._device != flow._device:
            flow = flow.to_device(self._device)
        if mode not in [1, 2, 3]:
            raise ValueError("Error combining flows: Mode needs to be 1, 2 or 3")
        thresholded = False if thresholded is None else thresholded
        if not isinstance(thresholded, bool):
            raise TypeError("Error combining flows: Thresholded needs to be a boolean")

        if mode == 3 and self._deferred():
            # flow_class.py:1729-1744 and utils.py:497-498 as rules for the kernel (first match wins, like the
            # reference's if-ladder): self zero -> `flow`; flow zero -> `self`; warping flow zero -> plain sum
            bit = STAT_NONZERO_TM if thresholded else STAT_NONZERO_M
            if self._ref == 's':        # v1 + W(-v1, [v2|m2]): self warps (f), flow is the payload (p)
                rules = _ops.early_rules(_ops.early_rule('f', bit, 2), _ops.early_rule('p', bit, 3),
                                         _ops.early_rule('f', STAT_NONZERO_T, 1))
                return self._fused_warp_add(flow, sign=-1.0, rules=rules)
            rules = _ops.early_rules(_ops.early_rule('p', bit, 3), _ops.early_rule('f', bit, 2),
                                     _ops.early_rule('f', STAT_NONZERO_T, 1))
            return flow._fused_warp_add(self, sign=1.0, rules=rules)
        if self._all_zero(thresholded, masked=True):                 # flow_class.py:1729-1744
            return flow
        if flow._all_zero(thresholded, masked=True):
            return self if mode == 3 else self.invert()

        s = self._ref == 's'
        if mode == 3:
            if s:   # self + self.invert('t').apply(flow)  = v1 + W(-v1, [v2|m2]),  mask theta & m1
                return self._fused_warp_add(flow, sign=-1.0)
            # flow + flow.apply(self) = v2 + W(v2, [v1|m1]),  mask theta & m2
            return flow._fused_warp_add(self, sign=1.0)
        if mode == 2:
            if s:
                return self.apply(flow - self)                       # flow_class.py:1768
            return flow - flow.apply(self.invert().apply(self))      # flow_class.py:1773
        if s:                                                        # mode 1, flow_class.py:1759-1760
            flow_inv_t = flow.invert('t')
            return flow - (flow_inv_t + flow_inv_t.apply(self.switch_ref())).apply(self)
        return self.invert().apply(flow - self)                      # flow_class.py:1763

    def combine(self, other: FlowAlias, mode: int, ref: str = None) -> FlowAlias:
        """flow_1 (+) flow_2 = flow_3 for flows of ANY reference, result in reference `ref` (reference
        flow_class.py:1812-1939; the successor of `combine_with`).  Pure composition of apply / invert / switch_ref,
        so it runs entirely on the fused kernels.

        mode 1: self = flow_2, other = flow_3, result flow_1;  mode 2: self = flow_1, other = flow_3, result flow_2;
        mode 3: self = flow_1, other = flow_2, result flow_3."""
        if not isinstance(other, Flow):
            raise TypeError("Error combining flows: Flow need to be of type 'Flow'")
        if self.shape != other.shape:
            raise ValueError("Error combining flows: Flow fields need to have the same shape, including batch size")
        if self._device != other._device:
            other = other.to_device(self._device)
        if mode not in [1, 2, 3]:
            raise ValueError("Error combining flows: Mode needs to be 1, 2 or 3")
        ref = self._ref if ref is None else get_valid_ref(ref)

        # "times" 0,1,2 = the three images the flows connect: flow_1: 0->1, flow_2: 1->2, flow_3: 0->2
        span = {1: (0, 1), 2: (1, 2), 3: (0, 2)}                     # flow index -> (source time, target time)
        sign = [[0, 1, 1], [-1, 0, 1], [-1, -1, 0]]                  # direction of travel between two times
        unknown = mode
        known_self, known_other = {1: (2, 3), 2: (1, 3), 3: (1, 2)}[mode]
        s_time, t_time = span[unknown]
        goal = s_time if ref == 's' else t_time                      # time frame the result must live in
        rest = ({0, 1, 2} - {s_time, t_time}).pop()                  # the third time

        def frame(flow, idx):                                        # time frame a flow's vectors are anchored in
            return span[idx][0] if flow._ref == 's' else span[idx][1]

        # the "close" input touches the goal time, the "far" one does not
        if (mode == 1 and goal == s_time) or (mode in (2, 3) and goal == t_time):
            close, close_idx, far, far_idx = other, known_other, self, known_self
        else:
            close, close_idx, far, far_idx = self, known_self, other, known_other
        close_frame, far_frame = frame(close, close_idx), frame(far, far_idx)

        if far_frame in (s_time, t_time):                            # anchored "around the corner": bring it adjacent
            far = far.switch_ref()
        if close_frame == goal:                                      # move far from the third time to the goal time
            far = close.apply(far) if sign[rest][goal] == 1 else close.invert(ref='t').apply(far)
        if goal == t_time:
            result = far * sign[s_time][rest] + close * sign[rest][t_time]
        else:
            result = close * sign[s_time][rest] + far * sign[rest][t_time]
        if close_frame != goal:                                      # the sum lives at the third time: move it
            result = close.apply(result) if sign[rest][goal] == 1 else close.invert(ref='s').apply(result)
        result._ref = ref
        return result

    def _fused_warp_add(self, payload: FlowAlias, sign: float, rules: int = 0) -> FlowAlias:
        """v_self + W(sign*v_self, [v_payload | m_payload]) with mask (theta(mask channel) & m_self).
        Restates  self + Flow(sign*v_self,'t',m_self).apply(payload)  including apply_flow's early-out.
        With `rules` (deferred validation) every early-out is decided by the kernel and the result carries its own
        device flags, so a chain of combines never touches the host."""
        if rules:
            out, valid, fl = _ops.warp_t(self._vecs, payload._vecs, payload._mask, self._mask, sign=sign, add_flow=1.0,
                                         thresh=0.99999, want_valid=True, fflags=self._get_dflags(),
                                         pflags=payload._get_dflags(), early=rules, want_flags=True)
            res = Flow._make(out, self._ref, valid)
            res._dflags = fl
            return res
        if self._all_zero(thresholded=True, masked=False):
            return Flow._make(self._vecs + payload._vecs, self._ref, payload._mask & self._mask)
        out, valid = _ops.warp_t(self._vecs, payload._vecs, payload._mask, self._mask, sign=sign, add_flow=1.0,
                                 thresh=0.99999, want_valid=True)
        return Flow._make(out, self._ref, valid)
