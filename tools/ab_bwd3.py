"""A/B of the W adjoint implementations on one box (tuning aid): register-merge (ofb_set_option(1, 0)) against the
tile-owned fixed-point accumulation (7, the default).  Prints ms per launch at the bench size (incl. the zero-fill of
gsrc) and the deviation of gsrc / gflow from the register-merge kernel.
usage: python tools/ab_bwd3.py [batch] [sigma]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _native, _ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda", sigma=sigma)
flow[0, :, :40, :] = 300.0          # samples outside the field
flow[1, 0, 500:540, :] += torch.arange(1920, device="cuda") * 0.9     # violent compression: footprint fallback
lib = _native.lib()


def run(impl, c, gflow=True, steps=8):
    lib.ofb_set_option(1, impl)
    src = img[:, :c].contiguous() if c <= 3 else torch.cat([img, img[:, :2]], 1).contiguous()
    go = gout[:, :c].contiguous() if c <= 3 else torch.cat([gout, gout[:, :2]], 1).contiguous()
    for _ in range(2):
        gf, gs = _ops._warp_t_bwd(flow, 1.0, src, go, 0.0, gflow, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gf, gs = _ops._warp_t_bwd(flow, 1.0, src, go, 0.0, gflow, True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, gf, gs


for c in (3, 2, 1, 5):
    for gflow in (True, False):
        t0, gf0, gs0 = run(0, c, gflow)
        t7, gf7, gs7 = run(7, c, gflow)
        scale = float(gs0.abs().max())
        dgf = float((gf7 - gf0).abs().max()) if gflow else 0.0
        print("C=%d gflow=%d  merge %.3f ms  tile %.3f ms   max|dgsrc| %.3e (scale %.2f, rel %.2e)  max|dgflow| %.3e" % (
            c, gflow, t0, t7, float((gs7 - gs0).abs().max()), scale, float((gs7 - gs0).abs().max()) / scale, dgf), flush=True)
