"""How much would an L2-resident gradient image buy the W adjoint?  One 1080p item (gsrc = 25 MB): the adjoint kernel
timed alone (CUDA events) right after its zero-fill, with and without an L2 flush (a 512 MB write) in between."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _native, _ops
import ctypes

flow, img, mask, tmask, gout = bench.make_inputs(8, "cuda")
lib = _native.lib()
junk = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
n, c, h, w = img.shape
hw = h * w
gflow = torch.empty(n, 2, h, w, device="cuda")
gsrc = torch.empty_like(img)
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def item_call(k):
    # one item through the C ABI (the call zero-fills gsrc itself, then launches the kernel)
    lib.ofb_warp_t_bwd(p(flow[k]), 2 * hw, 1.0, p(img[k]), c * hw, p(gout[k]), 0.0, p(gflow[k]), p(gsrc[k]), c * hw,
                       1, c, h, w, st)


def run(flush, reps=6):
    ts = []
    for r in range(reps):
        for k in range(n):
            if flush:
                junk.fill_(r)          # evicts whatever the previous call left in the L2 (inputs included)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            item_call(k)
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


print("per-item call (memset + kernel), L2 warm from the previous item's traffic: %.1f us" % (run(False) * 1e3))
print("per-item call (memset + kernel), L2 flushed before each call            : %.1f us" % (run(True) * 1e3))
print("whole-batch call / 8                                                      :", end=" ")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib.ofb_warp_t_bwd(p(flow), 2 * hw, 1.0, p(img), c * hw, p(gout), 0.0, p(gflow), p(gsrc), c * hw, n, c, h, w, st)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    lib.ofb_warp_t_bwd(p(flow), 2 * hw, 1.0, p(img), c * hw, p(gout), 0.0, p(gflow), p(gsrc), c * hw, n, c, h, w, st)
e1.record()
torch.cuda.synchronize()
print("%.1f us" % (e0.elapsed_time(e1) / 5 / n * 1e3))
