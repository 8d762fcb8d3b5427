import ctypes, sys, torch
sys.path.insert(0, '.')
from oflibpytorch_b200 import _native
lib = _native.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for d in (1919.0, 1079.0, 3.0):
    tot = 0
    for e in range(0, 152):                  # biased exponent of x
        lo, hi = e << 23, (e + 1) << 23
        bad = torch.zeros(1, dtype=torch.int64, device='cuda')
        lib.ofb_selftest_div(d, lo, hi, ctypes.c_void_p(bad.data_ptr()), st)
        torch.cuda.synchronize()
        b = int(bad.item())
        tot += b
        if b:
            print('d=%g exponent %d (2^%d): %d mismatches' % (d, e, e - 127, b))
    print('d=%g total %d' % (d, tot))
