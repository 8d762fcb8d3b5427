"""ms per call of combine_with mode 3 (fused warp + add + result flags) at 16 x 4K in 'deferred' mode, forward only and
forward + backward (tuning aid).  usage: python tools/combine_time.py [fwd_impl]"""
import sys, torch
sys.path.insert(0, ".")
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _native
if len(sys.argv) > 1: _native.lib().ofb_set_option(5, int(sys.argv[1]))
of.set_validation('deferred')
n, h, w = 16, 2160, 3840
f1 = bench.smooth_flow(n, h, w, 4.0, 3, 'cuda'); f2 = bench.smooth_flow(n, h, w, 4.0, 4, 'cuda')
m = torch.rand(n, h, w, device='cuda') > 0.05
def t(fn, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
def fwd():
    return of.Flow(f1, 't', m).combine_with(of.Flow(f2, 't', m), 3)
def both():
    a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_()
    r = of.Flow(a, 't', m).combine_with(of.Flow(b, 't', m), 3)
    r.vecs.backward(f1)
print("combine_with mode 3, 16x4K: fwd %.3f ms   fwd+bwd %.3f ms" % (t(fwd), t(both)))
