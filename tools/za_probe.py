"""Zero-ahead mode of the tile adjoint (ofb_set_option key 6) against the memset path: time and results at the bench size,
plus a violent flow whose footprints reach far down (deferred tiles).  usage: python tools/za_probe.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
import bench
from oflibpytorch_b200 import _ops, _native
lib = _native.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flow, img, mask, tmask, gout = bench.make_inputs(n, "cuda")
def run(za, fl, c, gflow, steps=10):
    lib.ofb_set_option(6, za)
    src, go = img[:, :c].contiguous(), gout[:, :c].contiguous()
    for _ in range(3): r = _ops._warp_t_bwd(fl, 1.0, src, go, 0.0, gflow, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): r = _ops._warp_t_bwd(fl, 1.0, src, go, 0.0, gflow, True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, r
wild = flow.clone()
wild[:, 1, 200:260, 300:900] = -400.0       # samples 400 rows further down: footprints far below the tile
wild[:, 1, 700:760, 300:900] = 500.0        # and far above
wild[:, 0, 400:420, :] = 3000.0             # outside
for name, fl in (("bench flow", flow), ("violent flow", wild)):
    for c, gf in ((3, True), (3, False), (2, True), (1, True), (4, True)):
        t0, r0 = run(0, fl, c, gf)
        t1, r1 = run(1, fl, c, gf)
        scale = float(r0[1].abs().max())
        dev = float((r0[1] - r1[1]).abs().max()) / scale
        gfe = True if not gf else torch.equal(r0[0], r1[0])
        print("%s C=%d gflow=%d: memset path %.3f ms  zero-ahead %.3f ms  gsrc max dev %.2e of max  gflow equal %s"
              % (name, c, gf, t0, t1, dev, gfe), flush=True)
