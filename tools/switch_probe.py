import sys, torch
sys.path.insert(0, '.')
import bench, oflibpytorch_b200 as of
n, H, W = 8, 2160, 3840
fa = bench.smooth_flow(n, H, W, 8.0, 8, 'cuda')
m = torch.ones(n, H, W, dtype=torch.bool, device='cuda'); m[:, 600:1000, 1400:2200] = False; m[:, :, ::97] = False
A = of.Flow(fa, 't', m)
for _ in range(3):
    with torch.no_grad():
        r = A.switch_ref()
torch.cuda.synchronize()
