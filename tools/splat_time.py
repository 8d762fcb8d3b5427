"""ms per call of the S forward paths at the bench sizes (tuning aid): Flow.apply 's' + valid area (cfg 3, 32 x 3 x 1080p)
and switch_ref (cfg 4, 16 x 4K), forward only, 'deferred' validation.  usage: python tools/splat_time.py"""
import sys, torch
sys.path.insert(0, ".")
import bench, oflibpytorch_b200 as of
of.set_validation('deferred')
def t(fn, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, r
flow, img, mask, tmask, gout = bench.make_inputs(32, "cuda")
with torch.no_grad():
    a, ra = t(lambda: of.Flow(flow, 's', mask).apply(img, tmask, return_valid_area=True))
    f4 = bench.smooth_flow(16, 2160, 3840, 8.0, 3, 'cuda'); m4 = torch.rand(16, 2160, 3840, device='cuda') > 0.05
    b, rb = t(lambda: of.Flow(f4, 't', m4).switch_ref())
print("apply s 32x3x1080p fwd %.3f ms   switch_ref 16x4K %.3f ms   checks %.6f %.6f %d %d"
      % (a, b, float(ra[0].double().mean()), float(rb.vecs.double().abs().mean()), int(ra[1].sum()), int(rb.mask.sum())))
