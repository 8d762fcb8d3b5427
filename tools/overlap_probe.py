"""Can the zero-fill of the payload gradient hide behind the forward kernel?  Times, at batch 32 x 3 x 1080p: the forward
alone, a memset of the payload gradient alone, the two back to back on one stream, and the two on two streams
(tuning aid).  usage: python tools/overlap_probe.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
import bench, oflibpytorch_b200 as of
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
of.set_validation('deferred')
buf = torch.empty_like(img)
side = torch.cuda.Stream()
def fwd():
    with torch.no_grad():
        return of.Flow(flow, "t", mask).apply(img, tmask, return_valid_area=True)
def serial():
    fwd(); buf.zero_()
def overlapped():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        buf.zero_()
    fwd()
    cur.wait_stream(side)
def t(fn, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
print("batch %d: fwd %.3f ms  memset %.3f ms  serial %.3f ms  two streams %.3f ms" % (n, t(fwd), t(lambda: buf.zero_()), t(serial), t(overlapped)))
