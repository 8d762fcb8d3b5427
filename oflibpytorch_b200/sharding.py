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

# ---- peer memory (NVLink P2P through torch symmetric memory) -------------------------------------------------
_PEER_CACHE = {}


def peer_memory_available(t: torch.Tensor, group=None) -> bool:
    """True when the halos of `t` can travel as direct peer reads: CUDA tensor, NCCL process group, torch symmetric
    memory importable, not switched off (OFB_ROWBAND_P2P=0)."""
    import os
    import torch.distributed as dist
    if os.environ.get('OFB_ROWBAND_P2P', '1') == '0' or not t.is_cuda:
        return False
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(group) != 'nccl':
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:       # noqa: BLE001
        return False
    return True


def _peer_buffer(nbytes: int, device, group, tag: str):
    """A symmetric (peer-mapped) byte buffer of at least `nbytes`, and its rendezvous handle.  Collective: every rank
    of the group must ask for the same size under the same tag.  Cached (allocation + rendezvous are slow)."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    grp = dist.group.WORLD if group is None else group
    key = (tag, grp.group_name, str(device))
    ent = _PEER_CACHE.get(key)
    if ent is None or ent[0].numel() < nbytes:
        size = max(int(nbytes), 1 << 20)
        buf = symm_mem.empty(size, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, grp.group_name)
        ent = (buf, hdl)
        _PEER_CACHE[key] = ent
    return ent


def _peer_view(hdl, rank: int, shape, dtype, byte_offset: int = 0):
    """Rank `rank`'s buffer as a tensor on THIS device: loads from it are P2P reads over NVLink."""
    elem = torch.empty((), dtype=dtype).element_size()
    return hdl.get_buffer(rank, tuple(shape), dtype, byte_offset // elem)


def exchange_row_halos_p2p(band: torch.Tensor, k: int, rank: int, world_size: int, group=None):
    """exchange_row_halos over peer memory: every rank stages its first and last k rows in its symmetric buffer, and
    each rank copies its neighbours' staged rows straight into the extended tensor (P2P reads over NVLink, no
    send/recv, no host synchronisation; two device-side barriers bracket the reads)."""
    if k <= 0 or world_size == 1:
        return band, 0
    if band.shape[-2] < k:
        raise ValueError("halo of %d rows exceeds this rank's band of %d rows; use fewer ranks" % (k, band.shape[-2]))
    lead = tuple(band.shape[:-2])
    hb, w = band.shape[-2:]
    rows = (lead + (k, w))
    nbytes = band[..., :k, :].numel() * band.element_size()
    buf, hdl = _peer_buffer(2 * nbytes, band.device, group, 'rows')
    mine_top = buf[:nbytes].view(band.dtype).view(rows)
    mine_bot = buf[nbytes:2 * nbytes].view(band.dtype).view(rows)
    mine_top.copy_(band[..., :k, :])
    mine_bot.copy_(band[..., -k:, :])
    hdl.barrier(channel=0)                                   # every rank's rows are staged
    up, down = rank - 1, rank + 1
    k_top = k if up >= 0 else 0
    k_bot = k if down < world_size else 0
    ext = torch.empty(lead + (k_top + hb + k_bot, w), dtype=band.dtype, device=band.device)
    ext[..., k_top:k_top + hb, :].copy_(band)
    if up >= 0:
        ext[..., :k, :].copy_(_peer_view(hdl, up, rows, band.dtype, nbytes))          # the upper rank's LAST rows
    if down < world_size:
        ext[..., k_top + hb:, :].copy_(_peer_view(hdl, down, rows, band.dtype, 0))   # the lower rank's FIRST rows
    hdl.barrier(channel=1)                                   # all reads done: the staging buffer may be reused
    return ext, k_top


def return_row_halos_p2p(ext: torch.Tensor, k_top: int, k_bot: int, rank: int, world_size: int, group=None):
    """Adjoint of exchange_row_halos: `ext` (..., k_top + hb + k_bot, W) holds gradients for this rank's own rows and
    for the halo rows that belong to its neighbours.  The halo rows are staged in peer memory; every rank then adds
    what its neighbours hold for ITS rows (ofb_add_rows reading peer memory, upper neighbour first, then lower: a
    fixed order).  Returns the gradient of the rank's own hb rows."""
    from . import _ops
    hb = ext.shape[-2] - k_top - k_bot
    w = ext.shape[-1]
    lead = tuple(ext.shape[:-2])
    own = ext[..., k_top:k_top + hb, :]
    if world_size == 1:
        return own.contiguous()
    k = max(k_top, k_bot)
    rows = lead + (k, w)
    nbytes = 4 * k * w
    for d in lead:
        nbytes *= d
    buf, hdl = _peer_buffer(2 * nbytes, ext.device, group, 'grad_rows')
    mine_top = buf[:nbytes].view(torch.float32).view(rows)            # what I hold for the rank ABOVE me
    mine_bot = buf[nbytes:2 * nbytes].view(torch.float32).view(rows)  # what I hold for the rank BELOW me
    if k_top:
        mine_top.copy_(ext[..., :k_top, :])
    if k_bot:
        mine_bot.copy_(ext[..., k_top + hb:, :])
    hdl.barrier(channel=0)
    own = own.contiguous()
    flat = own.view(-1, hb, w)
    if rank - 1 >= 0:                                        # the upper rank's bottom halo covers my first k rows
        _ops.add_rows(flat[:, :k], _peer_view(hdl, rank - 1, rows, torch.float32, nbytes).view(-1, k, w))
    if rank + 1 < world_size:                                # the lower rank's top halo covers my last k rows
        _ops.add_rows(flat[:, hb - k:], _peer_view(hdl, rank + 1, rows, torch.float32, 0).view(-1, k, w))
    hdl.barrier(channel=1)
    return own


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


def _exchange(band, k, rank, world_size, group):
    """Halo rows from the two neighbours: peer reads over NVLink when possible, send/recv otherwise (gloo, CPU)."""
    if peer_memory_available(band, group):
        return exchange_row_halos_p2p(band, k, rank, world_size, group)
    return exchange_row_halos(band, k, rank, world_size, group)


class _WarpRowBands(torch.autograd.Function):
    """Row-band 't' warp with its adjoint.  Forward: payload (and mask) halo rows come from the neighbours, the band is
    warped in the global frame (windowed kernel).  Backward: the windowed adjoint yields gradients for the own rows AND
    for the halo rows; the latter go back to the neighbours (`return_row_halos_p2p`), so the payload gradient of a
    band equals the corresponding rows of the single-GPU gradient.  Every rank must run the backward pass."""

    @staticmethod
    def forward(ctx, flow_band, payload_band, payload_mask_band, flow_mask_band, geom, rank, world_size, want_valid,
                group):
        from . import _ops
        a, b, h_total, halo = geom
        w = flow_band.shape[-1]
        ext, k_top = _exchange(payload_band.detach(), halo, rank, world_size, group)
        ext_mask = None
        if payload_mask_band is not None and want_valid:
            em, _ = _exchange(payload_mask_band.view(torch.uint8), halo, rank, world_size, group)
            ext_mask = em.view(torch.bool)
        window = [a, 0, h_total, w, a - k_top, 0, ext.shape[-2], w]
        out, valid = torch.ops.oflib_b200.warp_t_win_fwd(flow_band.detach(), 1.0, ext, ext_mask,
                                                        flow_mask_band if want_valid else None, 0.99999,
                                                        want_valid, window)
        ctx.save_for_backward(flow_band, ext)
        ctx.window, ctx.k_top, ctx.k_bot = window, k_top, ext.shape[-2] - k_top - (b - a)
        ctx.rank, ctx.world_size, ctx.group = rank, world_size, group
        ctx.mark_non_differentiable(valid)
        ctx.set_materialize_grads(False)
        return out, valid

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout, _gvalid):
        flow_band, ext = ctx.saved_tensors
        need_gflow, need_gsrc = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if gout is None:
            gout = torch.zeros((flow_band.shape[0], ext.shape[1]) + tuple(flow_band.shape[-2:]), device=ext.device)
        # the payload gradient is needed on every rank as soon as one rank needs it (the halos are exchanged)
        gf, gs_ext = torch.ops.oflib_b200.warp_t_win_bwd(flow_band, 1.0, ext, gout.contiguous(), need_gflow, True,
                                                         ctx.window)
        gsrc = None
        if peer_memory_available(gs_ext, ctx.group) or ctx.world_size == 1:
            gsrc = return_row_halos_p2p(gs_ext, ctx.k_top, ctx.k_bot, ctx.rank, ctx.world_size, ctx.group)
        else:
            gsrc = exchange_add_accumulator_halos(gs_ext.permute(0, 2, 3, 1).contiguous(), ctx.k_top, ctx.k_bot,
                                                  ctx.rank, ctx.world_size, ctx.group).permute(0, 3, 1, 2).contiguous()
        return (gf if need_gflow else None), (gsrc if need_gsrc else None), None, None, None, None, None, None, None


def apply_t_rowbands(flow_band: torch.Tensor, payload_band: torch.Tensor, *, h_total: int, rank: int, world_size: int,
                     flow_mask_band: torch.Tensor = None, payload_mask_band: torch.Tensor = None, halo: int = None,
                     want_valid: bool = True, group=None):
    """Backward ('t') warp of ONE field too large for one GPU, split into `world_size` contiguous row bands
    (`shard_range(h_total, world_size, rank)`): every rank holds its rows of the flow (N,2,hb,W), of the payload
    (N,C,hb,W) and of the masks, obtains `halo` payload rows from its two neighbours (direct peer reads over NVLink)
    and warps its own band in the GLOBAL coordinate frame (windowed kernel), so the result equals the corresponding
    rows of the single-GPU call bit for bit.  `halo` defaults to ceil(max |flow_y|) + 2 over all ranks (one scalar
    all-reduce).  Differentiable w.r.t. flow_band and payload_band: the gradients of the halo rows are handed back to
    the neighbours in the backward pass (every rank has to run it).
    Returns (out_band (N,C,hb,W), valid_band (N,hb,W) or None)."""
    import math
    import torch.distributed as dist
    a, b = shard_range(h_total, world_size, rank)
    hb, w = flow_band.shape[-2:]
    if hb != b - a:
        raise ValueError("flow band has %d rows, expected rows [%d, %d) of %d" % (hb, a, b, h_total))
    if halo is None:
        k = flow_band[:, 1].detach().abs().amax().reshape(1)
        if dist.is_available() and dist.is_initialized() and world_size > 1:
            dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
        halo = int(math.ceil(float(k.item()))) + 2
    if world_size > 1 and halo > h_total // world_size:
        raise ValueError("halo of %d rows exceeds the smallest band of %d rows; use fewer ranks"
                         % (halo, h_total // world_size))
    out, valid = _WarpRowBands.apply(flow_band, payload_band, payload_mask_band, flow_mask_band,
                                     (a, b, h_total, halo), rank, world_size, want_valid, group)
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
        if world_size > 1 and peer_memory_available(flow_band, group):
            # accumulate straight into peer-mapped memory; the neighbours' halo rows are then READ in place over
            # NVLink and added to the own rows (ofb_add_rows, fixed order: upper first, then lower)
            n = max(flow_band.shape[0], payload_band.shape[0])
            w = flow_band.shape[-1]
            kp = int(_ops._native.lib().ofb_splat_s_kp(payload_band.shape[1], int(dm is not None)))
            hmax = (h_total + world_size - 1) // world_size + 2 * halo        # the tallest extended band of any rank
            buf, hdl = _peer_buffer(4 * n * hmax * w * kp, flow_band.device, group, 'acc')
            acc_h = hi - lo
            acc = buf[:4 * n * acc_h * w * kp].view(torch.float32).view(n, acc_h, w, kp)
            _ops.splat_band_accumulate(flow_band, payload_band, dm, flow_mask_used, 1.0, True, (a, h_total, lo, acc_h),
                                       out=acc)
            hdl.barrier(channel=0)
            k_top, k_bot = a - lo, hi - b
            own = acc[:, k_top:k_top + hb].view(n, hb, w * kp)
            for nb in (rank - 1, rank + 1):
                if nb < 0 or nb >= world_size:
                    continue
                na, nbb = shard_range(h_total, world_size, nb)
                nlo, nhi = max(na - halo, 0), min(nbb + halo, h_total)
                peer = _peer_view(hdl, nb, (n, nhi - nlo, w * kp), torch.float32)
                if nb < rank:      # its rows [nbb, nhi) are my rows [a, a + nhi - nbb)
                    _ops.add_rows(own[:, :nhi - nbb], peer[:, nbb - nlo:nhi - nlo])
                else:              # its rows [nlo, na) are my rows [nlo, b)
                    _ops.add_rows(own[:, nlo - a:], peer[:, :na - nlo])
            out, valid, density, unocc = _ops.splat_band_normalise(
                acc, flow_band, payload_band, dm, flow_mask_used, True,
                _ops.VALID_GT if want_valid else _ops.VALID_NONE, (a, h_total, lo, acc_h), want_aux=True)
            hdl.barrier(channel=1)              # the neighbours have read my halo rows, I have normalised: reusable
        else:
            acc = _ops.splat_band_accumulate(flow_band, payload_band, dm, flow_mask_used, 1.0, True,
                                             (a, h_total, lo, hi - lo))
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
        g_ext, k_top = _exchange(g.view(n, hb, w * kg), halo, ctx.rank, ctx.world_size, ctx.group)
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
