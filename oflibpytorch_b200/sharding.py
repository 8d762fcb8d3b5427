"""Batch sharding of the hot path across the GPUs of one node.

Every hot-path operation is independent per batch item, so the batch (or frame) dimension is split into contiguous
shards, one per rank / GPU, and each rank runs the ordinary single-GPU calls on its shard: there is NO collective on
the data path.  The only cross-batch coupling in the reference is its zero-flow early-out, which looks at the WHOLE
batch passed to one call (``all(is_zero_flow(...))``, utils.py:497, flow_class.py:1046,1729,1738); a driver that
wants bit-identical behaviour to the unsharded call ANDs the per-shard flags first (`agree_all_zero`, one 4-byte
all-reduce, off the data path) and, if the global answer differs from the local one, avoids the early-out by
calling the kernels through `Flow` objects whose flags it overrides (`force_nonzero`).
"""
from typing import List, Sequence, Tuple

import torch


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of `n_items` for `rank` (frames [256g/G, 256(g+1)/G))."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank %d for world size %d" % (rank, world_size))
    return (n_items * rank) // world_size, (n_items * (rank + 1)) // world_size


def shard_batch(rank: int, world_size: int, *tensors):
    """Slice every tensor along dim 0 to this rank's shard; tensors with batch size 1 (broadcast) and None pass
    through unchanged."""
    out = []
    n = max((t.shape[0] for t in tensors if t is not None), default=0)
    a, b = shard_range(n, world_size, rank)
    for t in tensors:
        out.append(t if (t is None or t.shape[0] == 1) else t[a:b])
    return out if len(out) != 1 else out[0]


def agree_all_zero(local_all_zero: bool, group=None) -> bool:
    """Logical AND of the per-shard "whole shard is zero flow" flags over all ranks (what the unsharded call's
    early-out would have seen).  Works with any initialised process group (nccl or gloo); without one it is the
    identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return bool(local_all_zero)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    t = torch.tensor([1 if local_all_zero else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def force_nonzero(flow) -> None:
    """Mark a Flow as "not all zero" so that the hot-path methods skip their early-out even when this shard's
    vectors happen to be zero (used when `agree_all_zero` says some OTHER shard is not)."""
    stats = flow._get_stats()
    flow._stats = [s | 2 | 4 | 8 | 16 for s in stats]


# ----------------------------------------------------------------------------------------------------------------
# one oversized field: row bands with halo exchange ("single oversized field")
# ----------------------------------------------------------------------------------------------------------------

def exchange_row_halos(band: torch.Tensor, k: int, rank: int, world_size: int, group=None):
    """`band` (..., hb, W) holds this rank's rows of a field split into contiguous row bands.  Returns
    (extended, k_top): `extended` = [last k rows of rank-1 | band | first k rows of rank+1] (missing neighbours at
    the field's top / bottom contribute nothing) and k_top = number of halo rows prepended.  Point-to-point
    send/recv (NCCL over NVLink on GPUs, gloo on CPU); works for any dtype."""
    import torch.distributed as dist
    if k <= 0 or world_size == 1:
        return band, 0
    if band.shape[-2] < k:
        raise ValueError("halo of %d rows exceeds this rank's band of %d rows; use fewer ranks" % (k, band.shape[-2]))
    up, down = rank - 1, rank + 1
    ops, top, bot = [], None, None
    send_top = band[..., :k, :].contiguous()
    send_bot = band[..., -k:, :].contiguous()
    if up >= 0:
        top = torch.empty_like(send_top)
        ops += [dist.P2POp(dist.isend, send_top, up, group), dist.P2POp(dist.irecv, top, up, group)]
    if down < world_size:
        bot = torch.empty_like(send_bot)
        ops += [dist.P2POp(dist.isend, send_bot, down, group), dist.P2POp(dist.irecv, bot, down, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = [t for t in (top, band, bot) if t is not None]
    return (torch.cat(parts, dim=-2) if len(parts) > 1 else band), (k if top is not None else 0)


def apply_t_rowbands(flow_band: torch.Tensor, payload_band: torch.Tensor, *, h_total: int, rank: int, world_size: int,
                     flow_mask_band: torch.Tensor = None, payload_mask_band: torch.Tensor = None, halo: int = None,
                     want_valid: bool = True, group=None):
    """Backward ('t') warp of ONE field too large for one GPU, split into `world_size` contiguous row bands
    (`shard_range(h_total, world_size, rank)`): every rank holds its rows of the flow (N,2,hb,W), of the payload
    (N,C,hb,W) and of the masks, exchanges `halo` payload rows with its two neighbours and warps its own band in the
    GLOBAL coordinate frame (windowed kernel), so the result equals the corresponding rows of the single-GPU call bit
    for bit.  `halo` defaults to ceil(max |flow_y|) + 2 over all ranks (one scalar all-reduce).
    Returns (out_band (N,C,hb,W), valid_band (N,hb,W) or None).  Differentiable w.r.t. flow_band and payload_band
    (gradients of halo rows flow back to the neighbour only if the caller exchanges them; not done here)."""
    import math
    import torch.distributed as dist
    from . import _ops
    a, b = shard_range(h_total, world_size, rank)
    hb, w = flow_band.shape[-2:]
    if hb != b - a:
        raise ValueError("flow band has %d rows, expected rows [%d, %d) of %d" % (hb, a, b, h_total))
    if halo is None:
        k = flow_band[:, 1].detach().abs().amax().reshape(1)
        if dist.is_available() and dist.is_initialized() and world_size > 1:
            dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
        halo = int(math.ceil(float(k.item()))) + 2
    ext, k_top = exchange_row_halos(payload_band, halo, rank, world_size, group)
    ext_mask = None
    if payload_mask_band is not None:
        ext_mask, _ = exchange_row_halos(payload_mask_band.to(torch.uint8), halo, rank, world_size, group)
        ext_mask = ext_mask.to(torch.bool)
    window = [a, 0, h_total, w, a - k_top, 0, ext.shape[-2], w]
    out, valid = _ops.warp_t_win(flow_band, ext, window, ext_mask if want_valid else None,
                                 flow_mask_band if want_valid else None, want_valid=want_valid)
    return out, (valid if want_valid else None)


def exchange_add_accumulator_halos(acc: torch.Tensor, k_top: int, k_bot: int, rank: int, world_size: int, group=None):
    """`acc` (N, k_top + hb + k_bot, W, KP) is this rank's splat accumulator extended by halo rows that belong to its
    neighbours.  Sends the halo rows to their owners, receives the neighbours' halos for its own first / last rows and
    adds them in a FIXED order (own + lower neighbour + upper neighbour), so the result does not depend on arrival
    order.  Returns the accumulator restricted to this rank's own hb rows."""
    import torch.distributed as dist
    hb = acc.shape[1] - k_top - k_bot
    own = acc[:, k_top:k_top + hb]
    if world_size == 1:
        return own.contiguous()
    up, down = rank - 1, rank + 1
    ops, from_up, from_down = [], None, None
    # what the neighbours send has the height of THEIR halo on my side; by construction (same k everywhere, clipped
    # at the field's borders only) that is k rows, bounded by my band
    if up >= 0:
        send = acc[:, :k_top].contiguous()                       # my contributions to the band above
        from_up = torch.empty((acc.shape[0], min(k_top, hb)) + tuple(acc.shape[2:]), dtype=acc.dtype, device=acc.device)
        ops += [dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, from_up, up, group)]
    if down < world_size:
        send = acc[:, k_top + hb:].contiguous()                   # my contributions to the band below
        from_down = torch.empty((acc.shape[0], min(k_bot, hb)) + tuple(acc.shape[2:]), dtype=acc.dtype, device=acc.device)
        ops += [dist.P2POp(dist.isend, send, down, group), dist.P2POp(dist.irecv, from_down, down, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    own = own.clone()
    if from_up is not None:
        own[:, :from_up.shape[1]] += from_up
    if from_down is not None:
        own[:, hb - from_down.shape[1]:] += from_down
    return own


def apply_s_rowbands(flow_band: torch.Tensor, payload_band: torch.Tensor, *, h_total: int, rank: int, world_size: int,
                     flow_mask_band: torch.Tensor = None, payload_mask_band: torch.Tensor = None, halo: int = None,
                     consider_mask: bool = True, want_valid: bool = True, group=None):
    """Forward ('s') warp of ONE field split into contiguous row bands over `world_size` GPUs -- the splat counterpart
    of `apply_t_rowbands`, i.e. `Flow(flow,'s',mask).apply(payload, payload_mask, return_valid_area=True)` for a field
    that does not fit one GPU.  Every rank splats its source rows into a local accumulator extended by `halo` rows on
    both sides, the halos are exchanged with the two neighbours (P2P over NVLink) and added in a fixed order, and each
    rank normalises its own rows.  Differentiable with respect to the flow and the payload (the backward pass exchanges
    halo rows with the neighbours too, so every rank has to run it).  `halo` defaults to ceil(max |flow_y|) + 2 over
    all ranks.
    Returns (out_band (N,C,hb,W), valid_band (N,hb,W) or None)."""
    import math
    import torch.distributed as dist
    from . import _ops
    a, b = shard_range(h_total, world_size, rank)
    hb, w = flow_band.shape[-2:]
    if hb != b - a:
        raise ValueError("flow band has %d rows, expected rows [%d, %d) of %d" % (hb, a, b, h_total))
    if halo is None:
        k = flow_band[:, 1].detach().abs().amax().reshape(1)
        if dist.is_available() and dist.is_initialized() and world_size > 1:
            dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
        halo = int(math.ceil(float(k.item()))) + 2
    # every rank must take the same decision (bands differ by one row): test against the smallest band
    if world_size > 1 and halo > h_total // world_size:
        raise ValueError("halo of %d rows exceeds the smallest band of %d rows; use fewer ranks"
                         % (halo, h_total // world_size))
    # apply_flow's early-out (utils.py:497-498) looks at the WHOLE field: agree on it across the bands
    local_zero = bool((flow_band.detach().abs() < 1e-3).all().item())
    if agree_all_zero(local_zero, group):
        valid = None
        if want_valid:
            valid = payload_mask_band if payload_mask_band is not None else torch.ones(
                (payload_band.shape[0], hb, w), dtype=torch.bool, device=payload_band.device)
            if flow_mask_band is not None:                           # flow_class.py:880-895
                valid = valid & flow_mask_band
        return payload_band, valid
    lo, hi = max(a - halo, 0), min(b + halo, h_total)
    fm = flow_mask_band
    dm = payload_mask_band
    if want_valid:                                                   # flow_class.py:880-895: target mask AND flow mask
        if fm is not None:
            dm = fm if dm is None else (dm & fm)
        if consider_mask and dm is fm:
            dm = None                                                # "bad weight" slot provably zero
    else:
        dm = None
    flow_mask_used = fm if consider_mask else None
    out, valid = _SplatRowBands.apply(flow_band, payload_band, dm, flow_mask_used, (a, b, h_total, lo, hi, halo),
                                      rank, world_size, want_valid, group)
    return out, (valid if want_valid else None)


class _SplatRowBands(torch.autograd.Function):
    """Row-band splat with its adjoint.  Forward: accumulate own sources into the halo-extended accumulator, add the
    neighbours' halos, normalise own rows.  Backward: per-target prepass on own rows, COPY the neighbours' halo rows
    of it in, gather for own sources.  Both directions exchange with the two neighbours, so every rank must run the
    backward pass (an absent upstream gradient counts as zero)."""

    @staticmethod
    def forward(ctx, flow_band, payload_band, dm, flow_mask_used, geom, rank, world_size, want_valid, group):
        from . import _ops
        a, b, h_total, lo, hi, halo = geom
        hb = b - a
        acc = _ops.splat_band_accumulate(flow_band, payload_band, dm, flow_mask_used, 1.0, True, (a, h_total, lo, hi - lo))
        own = exchange_add_accumulator_halos(acc, a - lo, hi - b, rank, world_size, group)
        out, valid, density, unocc = _ops.splat_band_normalise(
            own, flow_band, payload_band, dm, flow_mask_used, True,
            _ops.VALID_GT if want_valid else _ops.VALID_NONE, (a, h_total, a, hb), want_aux=True)
        ctx.save_for_backward(flow_band, payload_band, flow_mask_used, out, density, unocc)
        ctx.geom, ctx.rank, ctx.world_size, ctx.group = geom, rank, world_size, group
        ctx.mark_non_differentiable(valid)
        ctx.set_materialize_grads(False)
        return out, valid

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout, _gvalid):
        from . import _ops
        flow_band, payload_band, flow_mask_used, out, density, unocc = ctx.saved_tensors
        a, b, h_total, lo, hi, halo = ctx.geom
        need_gflow, need_gdata = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if gout is None:
            gout = torch.zeros_like(out)
        g = _ops.splat_band_bwd_prepass(out, density, unocc, gout)                  # (N, hb, W, KG), own rows
        n, hb, w, kg = g.shape
        g_ext, k_top = exchange_row_halos(g.view(n, hb, w * kg), halo, ctx.rank, ctx.world_size, ctx.group)
        acc_h = g_ext.shape[-2]
        gflow, gdata = _ops.splat_band_bwd_gather(flow_band, payload_band, flow_mask_used, True,
                                                  g_ext.reshape(n, acc_h, w, kg), unocc, gout,
                                                  (a, h_total, a - k_top, acc_h), need_gflow, need_gdata)
        return (gflow if need_gflow else None), (gdata if need_gdata else None), None, None, None, None, None, None, None


def gather_shards(local: torch.Tensor, group=None) -> List[torch.Tensor]:
    """Optional epilogue (off the timed path): all-gather equally sized per-rank results."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    out = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, local.contiguous(), group=group)
    return out
