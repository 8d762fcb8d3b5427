"""cProfile of the host side of one bench step (where do the ~0.1 ms between the statistics sync and the forward
launch go?)."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
flow, img, mask, tmask, gout = bench.make_inputs(8, 'cuda')
for _ in range(5):
    bench.gpu_step(of, flow, img, mask, tmask, gout)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    f = flow.requires_grad_(); i = img.requires_grad_()
    fl = of.Flow(f, 't', mask)
torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(50):
    with torch.no_grad():
        out, valid = fl.apply(img, tmask, return_valid_area=True)
torch.cuda.synchronize()
t2 = time.perf_counter()
print('Flow() %.1f us per call (incl. stats kernel + sync), apply() %.1f us per call (incl. kernel)' % ((t1 - t0) / 50 * 1e6, (t2 - t1) / 50 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    bench.gpu_step(of, flow, img, mask, tmask, gout)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
