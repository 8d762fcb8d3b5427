"""Times the UNMODIFIED reference (oflibpytorch 2.1.1, /root/reference/src) on the host cores of the build container
for the bench's workloads, as context for the `cpu_baseline` of bench.py (which times the C/OpenMP oracle port: the
Python reference cannot travel to the GPU box).  SURVEY.md 8(d): cfg 2 at N in {1, 4}, cfg 3 at N = 1.

    python baseline/time_reference_cpu.py > baseline/reference_cpu_timing.json
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, '/root/reference/src')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oflibpytorch as ref            # noqa: E402  (the reference)
import bench                          # noqa: E402  (input recipes only)

cores = os.cpu_count() or 1
torch.set_num_threads(cores)
H, W = bench.H, bench.W
out = {'reference': 'oflibpytorch 2.1.1 (unmodified, /root/reference/src), torch %s CPU' % torch.__version__,
       'cores': cores, 'where': 'build container (no GPU)', 'results': {}}


def timeit(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts)


for n in (1, 4):
    flow, img, mask, tmask, gout = bench.make_inputs(n, 'cpu')

    def cfg2():
        f = flow.clone().requires_grad_()
        i = img.clone().requires_grad_()
        o, v = ref.Flow(f, 't', mask).apply(i, tmask, return_valid_area=True)
        o.backward(gout)

    best, mean = timeit(cfg2, 3)
    out['results']['cfg2 Flow.apply t + valid area fwd+bwd, %dx3x1080x1920' % n] = {
        'ms_best': round(best * 1e3, 1), 'ms_mean': round(mean * 1e3, 1), 'Mpix/s': round(n * H * W / best / 1e6, 2)}

n = 1
flow = bench.smooth_flow(n, H, W, 8.0, 5, 'cpu')
flow[:, :, 200:440, 300:700] = 0.0
g = torch.Generator().manual_seed(6)
img = torch.rand(n, 3, H, W, generator=g)
mask = bench.holed(n, H, W, 'cpu')
tmask = mask.roll(131, dims=2)


def cfg3():
    with torch.no_grad():
        ref.Flow(flow, 's', mask).apply(img, tmask, return_valid_area=True)


best, mean = timeit(cfg3, 3)
out['results']['cfg3 Flow.apply s + valid area fwd, 1x3x1080x1920'] = {
    'ms_best': round(best * 1e3, 1), 'ms_mean': round(mean * 1e3, 1), 'Mpix/s': round(n * H * W / best / 1e6, 2)}

fa, fb = bench.smooth_flow(1, H, W, 8.0, 20, 'cpu'), bench.smooth_flow(1, H, W, 8.0, 21, 'cpu')
gout2 = torch.rand(1, 2, H, W, generator=g)


def comb3():
    a = fa.clone().requires_grad_()
    b = fb.clone().requires_grad_()
    o = ref.combine_flows(a, b, 3, 't')
    o.backward(gout2)


best, mean = timeit(comb3, 3)
out['results']['combine_flows mode 3 ref t fwd+bwd, 1x1080x1920'] = {
    'ms_best': round(best * 1e3, 1), 'ms_mean': round(mean * 1e3, 1), 'Mpix/s': round(H * W / best / 1e6, 2)}
print(json.dumps(out, indent=1))
