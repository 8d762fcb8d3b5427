"""The multi-GPU path's host logic on CPU: world_size-2 gloo process groups (one process per "GPU").
Covers: contiguous shard ranges, shard_batch broadcasting rules, the zero-flow agreement all-reduce, and that the
sharded computation (each rank on its shard, no data-path collective) equals the unsharded one -- checked with the
CPU oracle, since the product kernels need a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from oflibpytorch_b200.sharding import shard_range
    for n in (1, 7, 64, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(256, 8, 3) == (96, 128)          # frames [256g/G, 256(g+1)/G)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oflibpytorch_b200.sharding import agree_all_zero, gather_shards, shard_batch
        from helpers import holed_mask, smooth_flow, gen
        from oracle import oracle as orc
        torch.set_num_threads(1)
        n, c, h, w = 4, 2, 24, 40
        flow = smooth_flow(n, h, w, 3.0, 700)
        flow[:2] = 0.0                                    # rank 0's shard is all zero, rank 1's is not
        img = torch.rand(n, c, h, w, generator=gen(701))
        mask = holed_mask(n, h, w, 702)
        tmask = holed_mask(1, h, w, 703)                  # broadcast over the batch
        f_s, i_s, m_s, t_s = shard_batch(rank, world, flow, img, mask, tmask)
        assert f_s.shape[0] == n // world and t_s.shape[0] == 1
        local_zero = bool(orc.is_zero_flow(f_s, True).all())
        assert local_zero == (rank == 0)
        global_zero = agree_all_zero(local_zero)
        assert global_zero is False                       # the unsharded call would NOT take the early-out
        # each rank warps its own shard; rank 0 must not take the early-out the unsharded call would not take
        fl = orc.OFlow(f_s, 't', m_s)
        if local_zero and not global_zero:
            out = orc.warp_t(fl.vecs, torch.cat((i_s, (t_s.expand(f_s.shape[0], -1, -1)).unsqueeze(1).float()), 1))
            res, valid = out[:, :-1], (out[:, -1] > 0.99999) & m_s
        else:
            res, valid = fl.apply(i_s, t_s.expand(f_s.shape[0], -1, -1), return_valid_area=True)
        parts = gather_shards(res.contiguous())
        vparts = gather_shards(valid.to(torch.uint8).contiguous())
        if rank == 0:
            full, fvalid = orc.OFlow(flow, 't', mask).apply(img, tmask.expand(n, -1, -1), return_valid_area=True)
            ok = torch.equal(torch.cat(parts), full) and torch.equal(torch.cat(vparts).bool(), fvalid)
            q.put(('ok' if ok else 'mismatch between sharded and unsharded result'))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put('rank 0 failed: %r' % (e,))
        raise
    finally:
        dist.destroy_process_group()


def test_sharded_equals_unsharded_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == 'ok'


def test_agree_all_zero_without_process_group_is_identity():
    from oflibpytorch_b200.sharding import agree_all_zero
    assert agree_all_zero(True) is True and agree_all_zero(False) is False


def _halo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oflibpytorch_b200.sharding import exchange_row_halos, shard_range
        h, w, k = 23, 7, 3
        full = torch.arange(2 * 3 * h * w, dtype=torch.float32).view(2, 3, h, w)
        a, b = shard_range(h, world, rank)
        ext, k_top = exchange_row_halos(full[:, :, a:b].contiguous(), k, rank, world)
        lo, hi = max(a - k, 0), min(b + k, h)
        ok = torch.equal(ext, full[:, :, lo:hi]) and k_top == a - lo
        m = (full[0, 0] % 2 == 0)
        em, _ = exchange_row_halos(m[a:b].to(torch.uint8).contiguous(), k, rank, world)
        ok = ok and torch.equal(em.bool(), m[lo:hi])
        # accumulator halos: every rank holds partial sums for rows [a-k, b+k); after the exchange-add each rank's own
        # rows must equal the sum over ALL ranks' partials (integers, so the sum is exact in fp32)
        from oflibpytorch_b200.sharding import exchange_add_accumulator_halos
        kp = 4
        def partial(r):
            aa, bb = shard_range(h, world, r)
            l2, h2 = max(aa - k, 0), min(bb + k, h)
            g = torch.Generator().manual_seed(100 + r)
            t = torch.zeros(2, h, w, kp)
            t[:, l2:h2] = torch.randint(0, 50, (2, h2 - l2, w, kp), generator=g).float()
            return t, l2, h2
        mine, l2, h2 = partial(rank)
        own = exchange_add_accumulator_halos(mine[:, l2:h2].contiguous(), a - l2, h2 - b, rank, world)
        total = sum(partial(r)[0] for r in range(world))
        ok = ok and torch.equal(own, total[:, a:b])
        res = [None] * world
        dist.all_gather_object(res, bool(ok))
        if rank == 0:
            q.put('ok' if all(res) else 'halo mismatch %r' % (res,))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_exchange_row_halos_gloo(world):
    """Halo rows of a row-band split field: every rank ends up with exactly rows [a-k, b+k) of the full field."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == 'ok'
