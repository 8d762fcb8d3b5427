"""ms per launch of the default W adjoint at the bench size (tuning aid): python tools/bwd_time.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
def run(c, gflow, steps=10):
    src, go = img[:, :c].contiguous(), gout[:, :c].contiguous()
    for _ in range(3): _ops._warp_t_bwd(flow, 1.0, src, go, 0.0, gflow, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): _ops._warp_t_bwd(flow, 1.0, src, go, 0.0, gflow, True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
print("C=3 both %.3f  gsrc-only %.3f   C=2 both %.3f" % (run(3, True), run(3, False), run(2, True)))
