"""A/B of the W forward implementations (ofb_set_option key 5: 0 lean, 1 direct) on one box, for each library build
given on the command line (tuning aid).  Checks that the outputs are bit-identical and prints ms per launch.
usage: python tools/ab_fwd2.py [name=lib.so ...]"""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _native, _ops
lib = _native.lib()
flow, img, mask, tmask, gout = bench.make_inputs(32, "cuda")
flow[0, :, :40, :] = 300.0      # a region sampling outside the field
flow[1, :, 500:540, :] = -2000.0
def run(impl, c, steps=10):
    lib.ofb_set_option(5, impl)
    src = img[:, :c].contiguous() if c <= 3 else torch.cat([img, img[:, :1]], 1).contiguous()
    for _ in range(3):
        out, valid, *_ = _ops._warp_t_fwd(flow, 1.0, src, tmask, mask, 0.0, 1.0 - 1e-3, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out, valid, *_ = _ops._warp_t_fwd(flow, 1.0, src, tmask, mask, 0.0, 1.0 - 1e-3, True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out, valid
for c in (3, 2, 1, 4):
    t1, o1, v1 = run(1, c)
    t0, o0, v0 = run(0, c)
    print("C=%d direct %.4f ms  lean %.4f ms  equal out %s valid %s" % (c, t1, t0, torch.equal(o0, o1), torch.equal(v0, v1)), flush=True)
'''
libs = [a.split("=", 1) for a in sys.argv[1:]] or [("tree", "")]
for name, lib in libs:
    env = dict(os.environ)
    if lib:
        env["OFB_LIB_OVERRIDE"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("==", name)
    print(out.stdout.strip() if out.stdout.strip() else out.stderr[-800:])
