"""BASELINE.json config 5 on one GPU at reduced length: F frames of 4K flow, chained combine_with(mode 3) + apply with
valid area, one backward through the chain.  Prints ms per frame and Mpix/s (flow pixels per frame)."""
import sys, time, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W, C = 2160, 3840, 3
dev = 'cuda'
flows = [bench.smooth_flow(1, H, W, 4.0, 100 + i, dev) for i in range(F)]
g = torch.Generator(device=dev).manual_seed(7)
img = torch.rand(1, C, H, W, generator=g, device=dev)
def step():
    fl = [f.detach().requires_grad_() for f in flows]
    im = img.detach().requires_grad_()
    acc, loss = None, 0.0
    for i in range(F):
        cur = of.Flow(fl[i], 't')
        acc = cur if acc is None else acc.combine_with(cur, 3)
        out, valid = acc.apply(im, return_valid_area=True)
        loss = loss + out.sum()
    loss.backward()
    return loss
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
ms = e0.elapsed_time(e1) / 3
print('chain of %d 4K frames: %.2f ms per chain, %.3f ms per frame, %.1f Mpix/s (wall %.2f ms per chain); 155 B/px -> %.2f of HBM peak'
      % (F, ms, ms / F, F * H * W / ms / 1e3, (t1 - t0) / 3 * 1e3, 155 * F * H * W / (ms * 1e-3) / 1e9 / 6546.6))
