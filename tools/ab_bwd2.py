"""A/B of register-merge W adjoint options: python tools/ab_bwd2.py [batch]   (key 3: red.v4 quads on/off)"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _native, _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
lib = _native.lib()
lib.ofb_set_option(1, 0)


def run(steps=8):
    for _ in range(2):
        gf, gs = _ops._warp_t_bwd(flow, 1.0, img, gout, 0.0, True, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gf, gs = _ops._warp_t_bwd(flow, 1.0, img, gout, 0.0, True, True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, gf, gs


lib.ofb_set_option(3, 0)
t0, gf0, gs0 = run()
print("scalar REDs: %.3f ms" % t0, flush=True)
for rep in range(2):
    lib.ofb_set_option(3, 1)
    t1, gf1, gs1 = run()
    print("v4 quads   : %.3f ms   max|dgsrc| %.3e  max|dgflow| %.3e" % (
        t1, float((gs1 - gs0).abs().max()), float((gf1 - gf0).abs().max())), flush=True)
    lib.ofb_set_option(3, 0)
    t0, _, _ = run()
    print("scalar REDs: %.3f ms" % t0, flush=True)
