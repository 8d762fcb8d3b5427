"""Host-buffer entry point for the 't' warp: streams a host-resident batch through the GPU in chunks with the
host->device copies, the kernels and the device->host copies overlapped on three CUDA streams.

The reference accepts host data (numpy / CPU tensors) and moves it with `.to(device)` inside the call
(`utils.py:515-519`, `flow_class.py:225-226`), i.e. copy -> compute -> copy, serialised.  On a B200 the kernels take
~2 % of that wall time and PCIe the rest, so the only lever is to keep both PCIe directions busy at once: chunk k+1
uploads while chunk k computes while chunk k-1 downloads (double-buffered device inputs, CUDA events between streams,
no host synchronisation except the one validation sync per chunk that `Flow(...)` owns).
"""
from typing import Optional, Tuple

import torch

from .flow_class import Flow


class HostWarpPipeline:
    """Chunked `Flow(flow, ref, mask).apply(img, tmask, return_valid_area=True)` (+ optional backward with an
    upstream gradient) on pinned host buffers.  Results land in caller-provided (or internally allocated) pinned host
    tensors.  One instance per device; reuse it across steps so the device buffers are allocated once."""

    def __init__(self, device: torch.device, chunk: int = 4):
        self.device = torch.device(device)
        self.chunk = int(chunk)
        self.s_h2d = torch.cuda.Stream(self.device)
        self.s_d2h = torch.cuda.Stream(self.device)
        self._bufs = None

    def _device_buffers(self, flow_h, img_h, mask_h, tmask_h, gout_h):
        key = (tuple(flow_h.shape[1:]), tuple(img_h.shape[1:]), mask_h is not None, tmask_h is not None,
               gout_h is not None, self.chunk)
        if self._bufs is None or self._bufs[0] != key:
            def mk(t):
                return None if t is None else torch.empty((self.chunk,) + tuple(t.shape[1:]), dtype=t.dtype,
                                                          device=self.device)
            sets = [[mk(flow_h), mk(img_h), mk(mask_h), mk(tmask_h), mk(gout_h)] for _ in range(2)]
            self._bufs = (key, sets)
        return self._bufs[1]

    @staticmethod
    def pinned_like(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory()

    def run(self, flow_h: torch.Tensor, img_h: torch.Tensor, ref: str = 't', mask_h: Optional[torch.Tensor] = None,
            tmask_h: Optional[torch.Tensor] = None, gout_h: Optional[torch.Tensor] = None,
            out_h: Optional[torch.Tensor] = None, valid_h: Optional[torch.Tensor] = None,
            gflow_h: Optional[torch.Tensor] = None, gimg_h: Optional[torch.Tensor] = None
            ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
        """All `*_h` are pinned CPU tensors with the same leading batch size N.  With `gout_h` the backward pass runs
        too and the gradients w.r.t. flow and image are returned.  Returns (out_h, valid_h, gflow_h, gimg_h)."""
        n = flow_h.shape[0]
        want_grad = gout_h is not None
        out_h = self.pinned_like(img_h.shape, torch.float32) if out_h is None else out_h
        valid_h = self.pinned_like((n,) + tuple(flow_h.shape[2:]), torch.bool) if valid_h is None else valid_h
        if want_grad:
            gflow_h = self.pinned_like(flow_h.shape, torch.float32) if gflow_h is None else gflow_h
            gimg_h = self.pinned_like(img_h.shape, torch.float32) if gimg_h is None else gimg_h
        sets = self._device_buffers(flow_h, img_h, mask_h, tmask_h, gout_h)
        compute = torch.cuda.current_stream(self.device)
        starts = list(range(0, n, self.chunk))
        up_done = [None] * len(starts)          # event: inputs of chunk k are on the device
        free = [None, None]                     # event: buffer set b may be overwritten (its chunk finished computing)
        hosts = (flow_h, img_h, mask_h, tmask_h, gout_h)

        def upload(k):
            a = starts[k]
            b = min(a + self.chunk, n)
            bufset = sets[k % 2]
            with torch.cuda.stream(self.s_h2d):
                if free[k % 2] is not None:
                    self.s_h2d.wait_event(free[k % 2])
                for dst, src in zip(bufset, hosts):
                    if src is not None:
                        dst[:b - a].copy_(src[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_h2d)
            up_done[k] = ev

        with torch.cuda.device(self.device):
            upload(0)
            for k, a in enumerate(starts):
                b = min(a + self.chunk, n)
                m = b - a
                if k + 1 < len(starts):
                    upload(k + 1)                # enqueued BEFORE this chunk's validation sync
                compute.wait_event(up_done[k])
                d_flow, d_img, d_mask, d_tmask, d_gout = sets[k % 2]
                f = d_flow[:m].detach().requires_grad_(want_grad)
                i = d_img[:m].detach().requires_grad_(want_grad)
                out, valid = Flow(f, ref, None if d_mask is None else d_mask[:m]).apply(
                    i, None if d_tmask is None else d_tmask[:m], return_valid_area=True)
                gf = gi = None
                if want_grad:
                    # an all-zero chunk takes apply's early-out: `out` is the input image, the flow is unused
                    gf, gi = torch.autograd.grad(out, (f, i), d_gout[:m], allow_unused=True)
                    gf = torch.zeros_like(f) if gf is None else gf
                    gi = torch.zeros_like(i) if gi is None else gi
                done = torch.cuda.Event()
                done.record(compute)
                with torch.cuda.stream(self.s_d2h):
                    self.s_d2h.wait_event(done)
                    for dst, src in ((out_h, out), (valid_h, valid), (gflow_h, gf), (gimg_h, gi)):
                        if src is not None:
                            src.record_stream(self.s_d2h)
                            dst[a:b].copy_(src.detach(), non_blocking=True)
                    # the early-out can hand back views of this chunk's input buffers: they may only be overwritten
                    # by the upload of chunk k+2 once the copies above have read them
                    drained = torch.cuda.Event()
                    drained.record(self.s_d2h)
                free[k % 2] = drained
            compute.wait_stream(self.s_d2h)
            self.s_d2h.synchronize()            # the returned host tensors are complete when run() returns
        return out_h, valid_h, gflow_h, gimg_h
