"""Per-call timing of the headline step (Flow.apply 't' fwd+bwd, batch 32) with and without saved derivatives and
for the adjoint implementations (tuning aid).  usage: python tools/ab_step.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops, _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
of.set_validation('deferred')
def run(steps=8):
    res = None
    for it in range(3 + steps):
        if it == 3: _ops.PROFILE = []; torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        f = flow.detach().requires_grad_(); i = img.detach().requires_grad_()
        out, valid = of.Flow(f, "t", mask).apply(i, tmask, return_valid_area=True)
        out.backward(gout)
        res = (out.detach(), valid, f.grad, i.grad)
    e1.record(); torch.cuda.synchronize()
    prof, _ops.PROFILE = _ops.PROFILE, None
    per = {}
    for nme, a, b in prof: per.setdefault(nme, []).append(a.elapsed_time(b))
    return e0.elapsed_time(e1) / steps, {k: round(sum(v) / len(v), 4) for k, v in per.items()}, res
lib = _native.lib()
base = None
for impl, save in ((0, False), (7, False), (7, True)):
    lib.ofb_set_option(1, impl); _ops.set_save_derivatives(save)
    ms, per, res = run()
    if base is None: base = res
    d = [float((a.float() - b.float()).abs().max()) for a, b in zip(res, base)]
    print("impl %d save_derivatives %d: step %.3f ms  %s  max dev from first (out, valid, gflow, gimg) %s" % (impl, save, ms, per, d), flush=True)
