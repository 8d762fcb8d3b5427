"""Times the fused W adjoint with only one of its two outputs requested (flow gradient / payload gradient)."""
import sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops
n = 32
flow, img, mask, tmask, gout = bench.make_inputs(n, 'cuda')
def run(fg, ig, steps=5):
    for it in range(3 + steps):
        if it == 3:
            _ops.PROFILE = []
        f = flow.detach().requires_grad_(fg); i = img.detach().requires_grad_(ig)
        out, valid = of.Flow(f, 't', mask).apply(i, tmask, return_valid_area=True)
        out.backward(gout)
    torch.cuda.synchronize()
    prof, _ops.PROFILE = _ops.PROFILE, None
    ts = [a.elapsed_time(b) for nme, a, b in prof if nme == 'ofb_warp_t_bwd']
    return sum(ts) / len(ts)
print('both  %.3f ms' % run(True, True))
print('gflow %.3f ms' % run(True, False))
print('gsrc  %.3f ms' % run(False, True))
