"""W adjoint over the batch in chunks small enough for the payload gradient to stay L2-resident between its zero-fill
and its REDs (tuning probe): total ms for 64 items when called on slices of k items."""
import sys, torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _ops
n = 64
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
def run(k, steps=5):
    def once():
        for a in range(0, n, k):
            _ops._warp_t_bwd(flow[a:a + k], 1.0, img[a:a + k], gout[a:a + k], 0.0, True, True)
    g = None
    once(); torch.cuda.synchronize()
    # replay as a CUDA graph so that Python dispatch does not count
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): once()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): once()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for k in (64, 16, 8, 4, 2, 1):
    print("chunk %2d items (%4d MB of gradient): %.3f ms" % (k, k * 3 * 1080 * 1920 * 4 >> 20, run(k)), flush=True)
