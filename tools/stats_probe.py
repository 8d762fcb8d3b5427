"""Why is flow_stats slower inside the bench step than alone? Times it back to back, after a sync, and after a
large scatter kernel."""
import sys, torch
sys.path.insert(0, '.')
import bench
from oflibpytorch_b200 import _ops
flow, img, mask, tmask, gout = bench.make_inputs(64, 'cuda')
def t(fn, n=10):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    e[0].record()
    for i in range(n):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return [round(e[i].elapsed_time(e[i + 1]), 3) for i in range(n)]
print('back to back      ', t(lambda: _ops.flow_stats(flow, mask)))
print('no mask           ', t(lambda: _ops.flow_stats(flow, None)))
def with_sync():
    _ops.flow_stats(flow, mask).tolist()
print('with tolist sync  ', t(with_sync))
_ops.PROFILE = []
for _ in range(6):
    _ops.flow_stats(flow, mask).tolist()
torch.cuda.synchronize()
print('PROFILE hook      ', [round(a.elapsed_time(b), 3) for _, a, b in _ops.PROFILE]); _ops.PROFILE = None
big = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
def after_fill():
    big.fill_(1)
    _ops.PROFILE = []
    _ops.flow_stats(flow, mask)
for _ in range(4):
    after_fill(); torch.cuda.synchronize()
    print('after 1 GiB fill  ', [round(a.elapsed_time(b), 3) for _, a, b in _ops.PROFILE])
_ops.PROFILE = None
x = flow.view(-1)
print('torch.sum(flow)   ', t(lambda: x.sum()), 'GB', x.numel() * 4 / 1e9)
print('torch.max(flow)   ', t(lambda: x.max()))
y = torch.empty_like(x)
print('copy_             ', t(lambda: y.copy_(x)))
h = x.view(torch.int32)
print('bitwise_or reduce?', t(lambda: torch.count_nonzero(h)))
import ctypes
from oflibpytorch_b200 import _native
lib = _native.lib()
flags = torch.zeros(64, dtype=torch.int32, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t_: ctypes.c_void_p(t_.data_ptr())
for reps in (1, 50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        lib.ofb_flow_stats(P(flow), P(mask), P(flags), 64, 1080, 1920, st)
    e1.record(); torch.cuda.synchronize()
    print('direct C calls x%d: %.3f ms each' % (reps, e0.elapsed_time(e1) / reps))
