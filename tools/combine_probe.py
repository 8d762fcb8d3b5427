"""Time combine_with modes 1/2/3 for both references on n 4K fields; per-call breakdown by C entry point."""
import sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops, _native
n, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 2160, 3840
dev = 'cuda'
fa = bench.smooth_flow(n, H, W, 8.0, 8, dev); fb = bench.smooth_flow(n, H, W, 8.0, 9, dev)
m = torch.ones(n, H, W, dtype=torch.bool, device=dev); m[:, 600:1000, 1400:2200] = False; m[:, :, ::97] = False
for ref in 'ts':
    A, B = of.Flow(fa, ref, m), of.Flow(fb, ref, m.roll(211, dims=2))
    for mode in (1, 2, 3):
        with torch.no_grad():
            for _ in range(2): A.combine_with(B, mode)
            torch.cuda.synchronize()
            _ops.PROFILE = []; _native.launch_count_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): A.combine_with(B, mode)
            e1.record(); torch.cuda.synchronize()
        per = {}
        for nme, a, b in _ops.PROFILE: per[nme.replace('ofb_', '')] = per.get(nme.replace('ofb_', ''), 0.0) + a.elapsed_time(b) / 3
        _ops.PROFILE = None
        tot = e0.elapsed_time(e1) / 3
        print('ref %s mode %d: %.2f ms (%.1f Gpix/s), kernels %.2f ms, launches %d: %s' % (
            ref, mode, tot, n * H * W / tot / 1e6, sum(per.values()), _native.launch_count() // 3,
            {k: round(v, 2) for k, v in per.items()}), flush=True)
