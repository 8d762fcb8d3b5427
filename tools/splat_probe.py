"""Small driver for profiling the S kernels: Flow.apply 's' with masks (fwd + bwd) and switch_ref on n 1080p fields.
Per iteration it launches, in order: accumulate, normalise (apply), bwd prepass, bwd gather, accumulate, normalise
(switch_ref)."""
import sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
n, H, W, C = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 1080, 1920, 3
dev = 'cuda'
flow = bench.smooth_flow(n, H, W, 8.0, 5, dev); flow[:, :, 200:440, 300:700] = 0.0
g = torch.Generator(device=dev).manual_seed(6)
img = torch.rand(n, C, H, W, generator=g, device=dev)
gout = torch.rand(n, C, H, W, generator=g, device=dev)
mask = torch.ones(n, H, W, dtype=torch.bool, device=dev); mask[:, 300:500, 700:1100] = False; mask[:, :, ::97] = False
tmask = mask.roll(131, dims=2)
for _ in range(3):
    f = flow.detach().requires_grad_(); i = img.detach().requires_grad_()
    fl = of.Flow(f, 's', mask)
    o, v = fl.apply(i, tmask, return_valid_area=True)
    o.backward(gout)
    with torch.no_grad():
        r = fl.switch_ref()
torch.cuda.synchronize()
