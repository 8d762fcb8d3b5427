"""cfg-5 chain at reduced length in 'deferred' validation mode: ms per frame (tuning aid).  usage: chain_probe.py [frames]"""
import sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
of.set_validation('deferred')
ms = bench.workload_cfg5(of, torch.device('cuda'), 6546.6, 0, 1, frames_total=F, steps=3)
print(ms)
