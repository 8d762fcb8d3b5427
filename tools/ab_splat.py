"""A/B timing of the S (splat) kernels of several builds of the library on the same box (tuning aid).
usage: python tools/ab_splat.py [name=lib.so ...]   (the current build is always measured as "new")
Workloads: cfg3 = Flow.apply 's' + valid area on 32x3x1080x1920 (fwd and fwd+bwd), cfg4 = switch_ref on 16x2160x3840."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops
dev = "cuda"
def prof(fn, steps=5):
    for it in range(2 + steps):
        if it == 2: _ops.PROFILE = []
        fn()
    torch.cuda.synchronize()
    pr, _ops.PROFILE = _ops.PROFILE, None
    per = {}
    for nme, a, b in pr: per.setdefault(nme.replace("ofb_", ""), []).append(a.elapsed_time(b))
    return {k: round(sum(v) / steps, 4) for k, v in per.items()}
n, H, W, C = 32, 1080, 1920, 3
flow = bench.smooth_flow(n, H, W, 8.0, 5, dev)
mask = torch.ones(n, H, W, dtype=torch.bool, device=dev); mask[:, 300:500, 700:1100] = False; mask[:, :, ::97] = False
g = torch.Generator(device=dev).manual_seed(6)
img = torch.rand(n, C, H, W, generator=g, device=dev); gout = torch.rand(n, C, H, W, generator=g, device=dev)
tmask = mask.roll(131, dims=2)
def cfg3():
    f = flow.detach().requires_grad_(); i = img.detach().requires_grad_()
    out, valid = of.Flow(f, "s", mask).apply(i, tmask, return_valid_area=True)
    out.backward(gout)
print("cfg3", prof(cfg3))
del img, gout, flow
n, H, W = 16, 2160, 3840
fa = bench.smooth_flow(n, H, W, 8.0, 8, dev)
m = torch.ones(n, H, W, dtype=torch.bool, device=dev); m[:, 600:1000, 1400:2200] = False; m[:, :, ::97] = False
A = of.Flow(fa, "t", m)
def cfg4():
    with torch.no_grad():
        A.switch_ref()
print("cfg4", prof(cfg4))
'''
libs = [("new", "")] + [tuple(a.split("=", 1)) for a in sys.argv[1:]]
for rep in range(2):
    for name, lib in libs:
        env = dict(os.environ)
        env["OFB_SKIP_ABI_TABLE"] = "1"
        if lib:
            env["OFB_LIB_OVERRIDE"] = os.path.abspath(lib)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(name, " | ".join(out.stdout.strip().splitlines()[-2:]) if out.stdout.strip() else out.stderr[-600:])
