"""Where does the step spend the time that is not kernel time? Gaps between consecutive C-ABI calls (CUDA events)."""
import sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops
flow, img, mask, tmask, gout = bench.make_inputs(64, 'cuda')
for it in range(8):
    if it == 3:
        _ops.PROFILE = []
    bench.gpu_step(of, flow, img, mask, tmask, gout)
torch.cuda.synchronize()
pr, _ops.PROFILE = _ops.PROFILE, None
for i in range(len(pr) - 1):
    print('%-18s %.3f ms   gap to next %.3f ms' % (pr[i][0], pr[i][1].elapsed_time(pr[i][2]), pr[i][2].elapsed_time(pr[i + 1][1])))
