"""One OVERSIZED field (default 16K: 8640 x 15360, 3 channels; argv: H W) split into row bands over the ranks of a torchrun job
(SURVEY.md 8e): forward and backward of the 't' warp and of the 's' splat, halos exchanged as direct peer reads over
NVLink (torch symmetric memory).  Each result is checked against the single-GPU computation of the whole field, and
the time (CUDA events, max over ranks) is printed as one JSON object by rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29513 \
        tools/rowband_bench.py > profiles/r02_rowband_gG.json
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oflibpytorch_b200 as of                                                   # noqa: E402
from oflibpytorch_b200 import sharding                                           # noqa: E402
from oflibpytorch_b200.sharding import apply_s_rowbands, apply_t_rowbands, shard_range   # noqa: E402
from helpers import gen, holed_mask, smooth_flow                                 # noqa: E402

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n, c = 1, 3
h = int(sys.argv[1]) if len(sys.argv) > 1 else 8640
w = int(sys.argv[2]) if len(sys.argv) > 2 else 15360
flow = smooth_flow(n, h, w, 8.0, 1000, cell=64)
img = torch.rand(n, c, h, w, generator=gen(1001))
gout = torch.rand(n, c, h, w, generator=gen(1004))
fmask, tmask = holed_mask(n, h, w, 1002), holed_mask(n, h, w, 1003)
a, b = shard_range(h, world, rank)
fb, ib, gb = (t[:, :, a:b].contiguous().to(dev) for t in (flow, img, gout))
mb, tb = fmask[:, a:b].contiguous().to(dev), tmask[:, a:b].contiguous().to(dev)
halo = int(torch.ceil(flow[:, 1].abs().max()).item()) + 2      # agreed up front: no per-call all-reduce in the timing


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ok(flag):
    res = [None] * world
    dist.all_gather_object(res, bool(flag))
    return all(res)


def close(x, y):
    scale = float(y.abs().max())
    return bool(((x - y).abs() <= 1e-4 * max(1.0, scale) + 1e-4 * y.abs()).all())


kw = dict(h_total=h, rank=rank, world_size=world, flow_mask_band=mb, payload_mask_band=tb, halo=halo)
res = {'field': '%dx%dx%dx%d fp32' % (n, c, h, w), 'n_gpus': world, 'halo_rows': halo,
       'peer_memory': bool(world > 1 and sharding.peer_memory_available(fb)), 'ms': {}, 'parity': {}}

# ---- single-GPU references (every rank computes them for its own check; 8K fits one GPU, it is the scheme that is
# being exercised, at a size where the answer can still be checked)
F, I = flow.to(dev).requires_grad_(), img.to(dev).requires_grad_()
ot, vt = of.Flow(F, 't', fmask.to(dev)).apply(I, tmask.to(dev), return_valid_area=True)
gFt, gIt = torch.autograd.grad(ot, (F, I), gout.to(dev))
os_, vs = of.Flow(F, 's', fmask.to(dev)).apply(I, tmask.to(dev), return_valid_area=True)
gFs, gIs = torch.autograd.grad(os_, (F, I), gout.to(dev))
ot, os_ = ot.detach(), os_.detach()

# ---- 't'
f2, i2 = fb.clone().requires_grad_(), ib.clone().requires_grad_()
o2, v2 = apply_t_rowbands(f2, i2, **kw)
gf2, gi2 = torch.autograd.grad(o2, (f2, i2), gb)
res['parity']['t forward bit-identical'] = all_ok(torch.equal(o2.detach(), ot[:, :, a:b]) and torch.equal(v2, vt[:, a:b]))
res['parity']['t gradients (flow, payload incl. halo rows) within 1e-4'] = all_ok(
    close(gf2, gFt[:, :, a:b]) and close(gi2, gIt[:, :, a:b]))


def t_fwd():
    with torch.no_grad():
        apply_t_rowbands(fb, ib, **kw)


def t_fwdbwd():
    f, i = fb.detach().requires_grad_(), ib.detach().requires_grad_()
    o, _ = apply_t_rowbands(f, i, **kw)
    torch.autograd.grad(o, (f, i), gb)


res['ms']['t forward'] = round(timed(t_fwd), 3)
res['ms']['t forward + backward'] = round(timed(t_fwdbwd), 3)

# ---- 's'
f3, i3 = fb.clone().requires_grad_(), ib.clone().requires_grad_()
o3, v3 = apply_s_rowbands(f3, i3, **kw)
gf3, gi3 = torch.autograd.grad(o3, (f3, i3), gb)
err = (o3.detach() - os_[:, :, a:b]).abs()
res['parity']['s forward: masks bit-exact, floats within 1e-4/1e-3'] = all_ok(
    bool((err <= 1e-3 + 1e-4 * os_[:, :, a:b].abs()).all()) and torch.equal(v3, vs[:, a:b]))
res['parity']['s gradients within 1e-4'] = all_ok(close(gf3, gFs[:, :, a:b]) and close(gi3, gIs[:, :, a:b]))


def s_fwd():
    with torch.no_grad():
        apply_s_rowbands(fb, ib, **kw)


def s_fwdbwd():
    f, i = fb.detach().requires_grad_(), ib.detach().requires_grad_()
    o, _ = apply_s_rowbands(f, i, **kw)
    torch.autograd.grad(o, (f, i), gb)


res['ms']['s forward'] = round(timed(s_fwd), 3)
res['ms']['s forward + backward'] = round(timed(s_fwdbwd), 3)

# ---- the same field on ONE GPU, for the scaling ratio
if rank == 0:
    Fd, Id, Md, Td, Gd = flow.to(dev), img.to(dev), fmask.to(dev), tmask.to(dev), gout.to(dev)

    def one(ref, bwd):
        def fn():
            f, i = Fd.detach().requires_grad_(bwd), Id.detach().requires_grad_(bwd)
            o, _ = of.Flow(f, ref, Md).apply(i, Td, return_valid_area=True)
            if bwd:
                torch.autograd.grad(o, (f, i), Gd)
        return fn

    def t1(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res['ms_single_gpu'] = {'t forward': round(t1(one('t', False)), 3), 't forward + backward': round(t1(one('t', True)), 3),
                            's forward': round(t1(one('s', False)), 3), 's forward + backward': round(t1(one('s', True)), 3)}
dist.barrier()
if rank == 0:
    res['speedup_vs_single_gpu'] = {k: round(res['ms_single_gpu'][k] / v, 2) for k, v in res['ms'].items()}
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
sys.exit(0 if all(res['parity'].values()) else 1)
