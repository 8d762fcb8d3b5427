"""Row-band spatial sharding of ONE field over the ranks of a torchrun job: each rank warps its band after a halo
exchange over NCCL (NVLink P2P) and the result is compared bit for bit with the single-GPU warp of the whole field.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/spatial_2gpu.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import oflibpytorch_b200 as of
from oflibpytorch_b200.sharding import apply_s_rowbands, apply_t_rowbands, shard_range
from helpers import smooth_flow, holed_mask, gen

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
n, c, h, w = 1, 3, 4320, 7680                               # one 8K field
flow = smooth_flow(n, h, w, 8.0, 1000, cell=64)
img = torch.rand(n, c, h, w, generator=gen(1001))
fmask, tmask = holed_mask(n, h, w, 1002), holed_mask(n, h, w, 1003)
a, b = shard_range(h, world, rank)
fb, ib = flow[:, :, a:b].contiguous().to(dev), img[:, :, a:b].contiguous().to(dev)
mb, tb = fmask[:, a:b].contiguous().to(dev), tmask[:, a:b].contiguous().to(dev)
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    out, valid = apply_t_rowbands(fb, ib, h_total=h, rank=rank, world_size=world, flow_mask_band=mb, payload_mask_band=tb)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
full, vfull = of.Flow(flow.to(dev), 't', fmask.to(dev)).apply(img.to(dev), tmask.to(dev), return_valid_area=True)
ok = torch.equal(out, full[:, :, a:b]) and torch.equal(valid, vfull[:, a:b])
res = [None] * world
dist.all_gather_object(res, bool(ok))
if rank == 0:
    print('spatial row bands on %d GPUs: bit-identical to the single-GPU result: %s; %.2f ms per 8K field (incl. halo exchange)'
          % (world, all(res), (t1 - t0) * 1e3))
# ---- the splat counterpart: accumulator halos exchanged and added over NVLink P2P
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    outs, valids = apply_s_rowbands(fb, ib, h_total=h, rank=rank, world_size=world, flow_mask_band=mb, payload_mask_band=tb)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
fulls, vfulls = of.Flow(flow.to(dev), 's', fmask.to(dev)).apply(img.to(dev), tmask.to(dev), return_valid_area=True)
err = (outs - fulls[:, :, a:b]).abs()
tol = 1e-3 + 1e-4 * fulls[:, :, a:b].abs()
ok_s = bool((err <= tol).all()) and torch.equal(valids, vfulls[:, a:b])
res_s = [None] * world
dist.all_gather_object(res_s, bool(ok_s))
if rank == 0:
    print('spatial row-band SPLAT on %d GPUs: masks bit-exact and floats within 1e-4/1e-3 of the single-GPU result: %s; '
          '%.2f ms per 8K field (incl. accumulator halo exchange), max abs err %.2e'
          % (world, all(res_s), (t1 - t0) * 1e3, float(err.max())))
# ---- and its adjoint: gradients of the band's flow and image through accumulate / halo add / normalise, with the
# per-target prepass rows of the neighbours copied in over P2P in the backward pass
gout = torch.rand(n, c, h, w, generator=gen(1004))
gb = gout[:, :, a:b].contiguous().to(dev)
for it in range(3):
    f2, i2 = fb.clone().requires_grad_(), ib.clone().requires_grad_()
    o2, _v2 = apply_s_rowbands(f2, i2, h_total=h, rank=rank, world_size=world, flow_mask_band=mb, payload_mask_band=tb)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    o2.backward(gb)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
ff, fi = flow.to(dev).requires_grad_(), img.to(dev).requires_grad_()
of.Flow(ff, 's', fmask.to(dev)).apply(fi, tmask.to(dev), return_valid_area=True)[0].backward(gout.to(dev))
def close(x, y):
    scale = float(y.abs().max())
    return bool(((x - y).abs() <= 1e-4 * max(1.0, scale) + 1e-4 * y.abs()).all())
ok_g = close(f2.grad, ff.grad[:, :, a:b]) and close(i2.grad, fi.grad[:, :, a:b])
res_g = [None] * world
dist.all_gather_object(res_g, bool(ok_g))
if rank == 0:
    print('spatial row-band SPLAT ADJOINT on %d GPUs: band gradients within 1e-4 of the single-GPU gradients: %s; '
          '%.2f ms per 8K field (incl. halo exchange of the prepass rows)' % (world, all(res_g), (t1 - t0) * 1e3))
dist.destroy_process_group()
sys.exit(0 if (all(res) and all(res_s) and all(res_g)) else 1)
