"""A/B of the W adjoint implementations on one box (tuning aid): register-merge (0), ring + RED flush (1),
ring + TMA reduce flush (2), tile-owned accumulation (7, the default), swept over the ring's rows-per-task knob.  Prints ms per launch at the bench size and the
max deviation of gsrc / gflow from implementation 0.
usage: python tools/ab_bwd.py [batch] [sigma]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _native, _ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda", sigma=sigma)
lib = _native.lib()


def run(impl, rows, steps=6):
    lib.ofb_set_option(1, impl)
    lib.ofb_set_option(2, rows)
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


t0, gf0, gs0 = run(0, 24)
print("impl 0 (register merge): %.3f ms" % t0, flush=True)
scale = float(gs0.abs().max())
for impl in (1, 2, 7):
    for rows in (16, 24, 48):
        try:
            t, gf, gs = run(impl, rows)
        except Exception as e:     # noqa
            print("impl %d rows %d FAILED %s" % (impl, rows, e), flush=True)
            continue
        print("impl %d rows %3d: %.3f ms   max|dgsrc| %.3e (scale %.2f)  max|dgflow| %.3e" % (
            impl, rows, t, float((gs - gs0).abs().max()), scale, float((gf - gf0).abs().max())), flush=True)
