"""Secondary workloads (BASELINE.json configs 3 and 4) timed through the public API with CUDA events.
Not the bench line -- used to steer the optimisation of the splat kernels."""
import json, sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops

def timeit(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    _ops.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    prof, _ops.PROFILE = _ops.PROFILE, None
    per = {}
    for name, a, b in prof: per.setdefault(name, []).append(a.elapsed_time(b))
    return e0.elapsed_time(e1) / steps, {k: round(sum(v) / steps, 3) for k, v in per.items()}

def main():
    dev = 'cuda'
    res = {}
    # cfg 3: forward 's' warp via the splat, 1080p batch 32 (+5% exactly-zero flow blocks), fwd and fwd+bwd
    n, H, W, C = 32, 1080, 1920, 3
    flow = bench.smooth_flow(n, H, W, 8.0, 5, dev)
    flow[:, :, 200:440, 300:700] = 0.0
    g = torch.Generator(device=dev).manual_seed(6)
    img = torch.rand(n, C, H, W, generator=g, device=dev); gout = torch.rand(n, C, H, W, generator=g, device=dev)
    mask = torch.ones(n, H, W, dtype=torch.bool, device=dev); mask[:, 300:500, 700:1100] = False; mask[:, :, ::97] = False
    tmask = mask.roll(131, dims=2)
    px = n * H * W
    def s_fwd():
        with torch.no_grad():
            return of.Flow(flow, 's', mask).apply(img, tmask, return_valid_area=True)
    def s_fwdbwd():
        f = flow.detach().requires_grad_(); i = img.detach().requires_grad_()
        o, v = of.Flow(f, 's', mask).apply(i, tmask, return_valid_area=True)
        o.backward(gout)
    t, k = timeit(s_fwd); res['cfg3_s_apply_fwd'] = {'ms': round(t, 3), 'Gpix/s': round(px / t / 1e6, 2), 'frac_of_35B': round(35 * px / t / 1e6 / 6546.6, 3), 'calls': k}
    t, k = timeit(s_fwdbwd); res['cfg3_s_apply_fwdbwd'] = {'ms': round(t, 3), 'Gpix/s': round(px / t / 1e6, 2), 'frac_of_104B': round(104 * px / t / 1e6 / 6546.6, 3), 'calls': k}
    del img, gout, flow
    # cfg 4: combine mode 3 + switch_ref on 4K batch 16 with masks, both refs
    n, H, W = 16, 2160, 3840
    fa = bench.smooth_flow(n, H, W, 8.0, 8, dev); fb = bench.smooth_flow(n, H, W, 8.0, 9, dev)
    m = torch.ones(n, H, W, dtype=torch.bool, device=dev); m[:, 600:1000, 1400:2200] = False; m[:, :, ::97] = False
    px = n * H * W
    for ref in 'ts':
        A, B = of.Flow(fa, ref, m), of.Flow(fb, ref, m)
        def comb():
            with torch.no_grad(): return A.combine_with(B, 3)
        def sw():
            with torch.no_grad(): return A.switch_ref()
        t, k = timeit(comb); res['cfg4_combine3_' + ref] = {'ms': round(t, 3), 'Gpix/s': round(px / t / 1e6, 2), 'frac_of_27B': round(27 * px / t / 1e6 / 6546.6, 3), 'calls': k}
        t, k = timeit(sw); res['cfg4_switch_ref_' + ref] = {'ms': round(t, 3), 'Gpix/s': round(px / t / 1e6, 2), 'frac_of_18B': round(18 * px / t / 1e6 / 6546.6, 3), 'calls': k}
        def vt():
            with torch.no_grad(): return A.valid_target(), A.valid_source()
        t, k = timeit(vt); res['cfg4_valid_both_' + ref] = {'ms': round(t, 3), 'Gpix/s': round(2 * px / t / 1e6, 2), 'calls': k}
    for k, v in res.items(): print(k, json.dumps(v))

main()
