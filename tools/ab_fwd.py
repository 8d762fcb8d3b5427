"""A/B timing of the W forward/adjoint kernels of several builds of the library on the same box (tuning aid).
usage: python tools/ab_fwd.py [name=]lib.so ...   (each build must export the ABI subset used by the 't' warp; the
current build is measured as "new")"""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
import bench, oflibpytorch_b200 as of
from oflibpytorch_b200 import _ops
flow, img, mask, tmask, gout = bench.make_inputs(32, "cuda")
def run(steps=8):
    for it in range(3 + steps):
        if it == 3: _ops.PROFILE = []
        f = flow.detach().requires_grad_(); i = img.detach().requires_grad_()
        out, valid = of.Flow(f, "t", mask).apply(i, tmask, return_valid_area=True)
        out.backward(gout)
    torch.cuda.synchronize()
    prof, _ops.PROFILE = _ops.PROFILE, None
    per = {}
    for nme, a, b in prof: per.setdefault(nme, []).append(a.elapsed_time(b))
    return {k: round(sum(v) / len(v), 4) for k, v in per.items()}
print(run())
'''
libs = [(a.split("=", 1) if "=" in a else ("old", a)) for a in sys.argv[1:]]
libs = [(n, os.path.abspath(l)) for n, l in libs] + [("new", "")]
for rep in range(2):
    for name, lib in libs:
        env = dict(os.environ)
        env["OFB_SKIP_ABI_TABLE"] = "1"
        if lib:
            env["OFB_LIB_OVERRIDE"] = lib
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(name, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
