import sys, torch
sys.path.insert(0, '.')
import oflibpytorch_b200 as of
torch.manual_seed(0)
n, c, h, w = 1, 3, 64, 128
flow = torch.randn(n, 2, h, w, device='cuda')
img = torch.rand(n, c, h, w, device='cuda')
out = of.apply_flow(flow, img, 't')
torch.cuda.synchronize()
print('ok', float(out.sum()))
