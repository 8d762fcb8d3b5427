"""One configuration of the W adjoint for ncu: python tools/bwd_probe.py impl rows [batch] [launches]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _native, _ops
impl, rows = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
k = int(sys.argv[4]) if len(sys.argv) > 4 else 2
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
lib = _native.lib()
lib.ofb_set_option(1, impl)
lib.ofb_set_option(2, rows)
for _ in range(k):
    gf, gs = _ops._warp_t_bwd(flow, 1.0, img, gout, 0.0, True, True)
torch.cuda.synchronize()
print("ok", float(gs.sum()))
