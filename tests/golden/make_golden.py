"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oflibpytorch 2.1.1).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Every array written here is an input or an output of the reference's own public functions on CPU,
fp32, PURE_PYTORCH=True.  The oracle (oracle/) and the CUDA path are both tested against these files.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/reference/src')
import oflibpytorch as of  # noqa: E402
from oflibpytorch.utils import (apply_s_flow, get_flow_endpoints,  # noqa: E402
                                grid_from_unstructured_data)

HERE = os.path.dirname(os.path.abspath(__file__))
of.set_pure_pytorch()
torch.set_num_threads(1)


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def smooth_flow(n, h, w, sigma, seed):
    """bicubic upsampling of a coarse Gaussian field."""
    lo = torch.randn(n, 2, h // 16 + 2, w // 16 + 2, generator=gen(seed)) * sigma
    return torch.nn.functional.interpolate(lo, size=(h, w), mode='bicubic', align_corners=True).contiguous()


def holed_mask(n, h, w, seed):
    m = torch.rand(n, h, w, generator=gen(seed)) > 0.05
    m[:, h // 4:h // 2, w // 3:w // 2] = False
    m[:, :, ::11] = False
    return m


def npy(t):
    return t.detach().cpu().numpy()


out = {}


def put(case, **kw):
    for k, v in kw.items():
        out[case + '/' + k] = npy(v) if isinstance(v, torch.Tensor) else np.asarray(v)


N, C, H, W = 2, 3, 37, 53

# ---------------------------------------------------------------- W: apply_flow 't' + gradients
flow = smooth_flow(N, H, W, 4.0, 1).requires_grad_()
img = torch.rand(N, C, H, W, generator=gen(2)).requires_grad_()
gout = torch.rand(N, C, H, W, generator=gen(3))
res = of.apply_flow(flow, img, 't')
res.backward(gout)
put('warp_t', flow=flow, img=img, gout=gout, out=res, gflow=flow.grad, gimg=img.grad)

# large-magnitude flow: many samples leave the image (zero padding, partially in-bounds corners)
flow = (torch.randn(N, 2, H, W, generator=gen(4)) * 20).requires_grad_()
img = torch.rand(N, C, H, W, generator=gen(5)).requires_grad_()
res = of.apply_flow(flow, img, 't')
res.backward(gout)
put('warp_t_big', flow=flow, img=img, gout=gout, out=res, gflow=flow.grad, gimg=img.grad)

# ---------------------------------------------------------------- S core: grid_from_unstructured_data
x = (torch.rand(N, H, W, generator=gen(6)) * (W + 6) - 3).requires_grad_()
y = (torch.rand(N, H, W, generator=gen(7)) * (H + 6) - 3).requires_grad_()
data = torch.randn(N, C, H, W, generator=gen(8)).requires_grad_()
mask = torch.rand(N, H, W, generator=gen(9)) > 0.2
gd = torch.rand(N, H, W, generator=gen(10))
o, d = grid_from_unstructured_data(x, y, data, mask)
(o * gout).sum().add((d * gd).sum()).backward()
put('gfud', x=x, y=y, data=data, mask=mask, gout=gout, gdens=gd, out=o, dens=d,
    gx=x.grad, gy=y.grad, gdata=data.grad)
x2 = x.detach().clone().requires_grad_()
y2 = y.detach().clone().requires_grad_()
data2 = data.detach().clone().requires_grad_()
o, d = grid_from_unstructured_data(x2, y2, data2)
(o * gout).sum().backward()
put('gfud_nomask', out=o, dens=d, gx=x2.grad, gy=y2.grad, gdata=data2.grad)

# ---------------------------------------------------------------- endpoints (utils.py:1045)
fl = smooth_flow(N, H, W, 3.0, 11)
xs, ys = get_flow_endpoints(fl, 's')
xt, yt = get_flow_endpoints(fl, 't')
put('endpoints', flow=fl, xs=xs, ys=ys, xt=xt, yt=yt)

# ---------------------------------------------------------------- S: apply_s_flow / apply_flow 's'
flow = smooth_flow(N, H, W, 4.0, 12)
flow[:, :, 5:15, 8:20] = 0.0                      # zero-flow block -> occlusion rule
flow[:, :, 20:24, 30:40] = 5e-4                   # below threshold -> also "zero"
flow = flow.requires_grad_()
img = torch.rand(N, C, H, W, generator=gen(13)).requires_grad_()
mask = holed_mask(N, H, W, 14)
for occl in (True, False):
    f2 = flow.detach().clone().requires_grad_()
    i2 = img.detach().clone().requires_grad_()
    o, wm = apply_s_flow(f2, i2, mask, occlude_zero_flow=occl)
    o.backward(gout)
    put('apply_s_flow_occ%d' % occl, flow=f2, img=i2, mask=mask, gout=gout, out=o, wmask=wm,
        gflow=f2.grad, gimg=i2.grad)
f2 = flow.detach().clone().requires_grad_()
i2 = img.detach().clone().requires_grad_()
o = of.apply_flow(f2, i2, 's')
o.backward(gout)
put('apply_flow_s_nomask', out=o, gflow=f2.grad, gimg=i2.grad)

# ---------------------------------------------------------------- Flow.apply with validity
for ref in ('t', 's'):
    flow = smooth_flow(N, H, W, 4.0, 20)
    flow[:, :, 25:30, 10:25] = 0.0
    fmask = holed_mask(N, H, W, 21)
    tmask = holed_mask(N, H, W, 22)
    img = torch.rand(N, C, H, W, generator=gen(23))
    for cm in (True, False):
        f2 = flow.clone().requires_grad_()
        i2 = img.clone().requires_grad_()
        fl = of.Flow(f2, ref, fmask)
        o, v = fl.apply(i2, tmask, return_valid_area=True, consider_mask=cm)
        o.backward(gout)
        put('flow_apply_%s_cm%d' % (ref, cm), flow=flow, fmask=fmask, tmask=tmask, img=img, gout=gout,
            out=o, valid=v, gflow=f2.grad, gimg=i2.grad)
    # no target mask, no valid area
    o = of.Flow(flow, ref, fmask).apply(img)
    put('flow_apply_%s_plain' % ref, out=o)
    # batch broadcasting: single flow, batched target; batched flow, single (C,H,W) / (H,W) target
    o, v = of.Flow(flow[:1], ref, fmask[:1]).apply(img, return_valid_area=True)
    put('flow_apply_%s_bc1' % ref, out=o, valid=v)
    o, v = of.Flow(flow, ref, fmask).apply(img[0], tmask[0], return_valid_area=True)
    put('flow_apply_%s_bc2' % ref, out=o, valid=v)
    o = of.Flow(flow, ref, fmask).apply(img[0, 0])
    put('flow_apply_%s_bc3' % ref, out=o)
    # uint8 target (rounding rule, flow_class.py:943-949)
    img8 = (img * 255).to(torch.uint8)
    o = of.Flow(flow, ref, fmask).apply(img8)
    put('flow_apply_%s_u8' % ref, img8=img8, out=o)
    o8 = of.apply_flow(flow, img8, ref, fmask)
    put('apply_flow_%s_u8' % ref, out=o8)
    # Flow target
    flow_b = smooth_flow(N, H, W, 3.0, 24)
    mask_b = holed_mask(N, H, W, 25)
    r = of.Flow(flow, ref, fmask).apply(of.Flow(flow_b, ref, mask_b))
    put('flow_apply_%s_flow' % ref, flow_b=flow_b, mask_b=mask_b, vecs=r.vecs, mask=r.mask)
    # padding / cut
    pad = [3, 5, 2, 7]
    big = torch.rand(N, C, H + pad[0] + pad[1], W + pad[2] + pad[3], generator=gen(26))
    bigm = torch.rand(N, H + pad[0] + pad[1], W + pad[2] + pad[3], generator=gen(27)) > 0.1
    for cut in (True, False):
        o, v = of.Flow(flow, ref, fmask).apply(big, bigm.clone(), return_valid_area=True, padding=pad, cut=cut)
        put('flow_apply_%s_pad_cut%d' % (ref, cut), big=big, bigm=bigm, pad=np.array(pad), out=o, valid=v)

# ---------------------------------------------------------------- switch_ref / invert / valid areas
flow = smooth_flow(N, H, W, 3.0, 30)
flow[:, :, 0:6, 0:9] = 0.0
fmask = holed_mask(N, H, W, 31)
gv = torch.rand(N, 2, H, W, generator=gen(32))
for ref in ('t', 's'):
    f2 = flow.clone().requires_grad_()
    r = of.Flow(f2, ref, fmask).switch_ref()
    r.vecs.backward(gv)
    put('switch_ref_%s' % ref, flow=flow, fmask=fmask, gv=gv, vecs=r.vecs, mask=r.mask, gflow=f2.grad)
    f2 = flow.clone().requires_grad_()
    r = of.Flow(f2, ref, fmask).invert()
    r.vecs.backward(gv)
    put('invert_%s' % ref, vecs=r.vecs, mask=r.mask, gflow=f2.grad)
    for cm in (True, False):
        put('valid_%s_cm%d' % (ref, cm),
            vt=of.Flow(flow, ref, fmask).valid_target(consider_mask=cm),
            vs=of.Flow(flow, ref, fmask).valid_source(consider_mask=cm))
    put('tensor_api_%s' % ref, sw=of.switch_flow_ref(flow, ref), inv=of.invert_flow(flow, ref),
        vt=of.valid_target(flow, ref), vs=of.valid_source(flow, ref))

# ---------------------------------------------------------------- combine_with, all modes and refs
fa = smooth_flow(N, H, W, 3.0, 40)
fb = smooth_flow(N, H, W, 3.0, 41)
ma = holed_mask(N, H, W, 42)
mb = holed_mask(N, H, W, 43)
put('combine_in', fa=fa, fb=fb, ma=ma, mb=mb, gv=gv)
for ref in ('t', 's'):
    for mode in (1, 2, 3):
        a = fa.clone().requires_grad_()
        b = fb.clone().requires_grad_()
        r = of.Flow(a, ref, ma).combine_with(of.Flow(b, ref, mb), mode)
        r.vecs.backward(gv)
        put('combine_%s_%d' % (ref, mode), vecs=r.vecs, mask=r.mask, ga=a.grad, gb=b.grad)
        put('combine_flows_%s_%d' % (ref, mode), vecs=of.combine_flows(fa, fb, mode, ref))

# ---------------------------------------------------------------- config 1: bundled sintel.flo (10x20)
sf = of.Flow.from_sintel('/root/reference/test/sintel.flo')
img = torch.rand(3, 10, 20, generator=gen(50))
put('sintel', flow=sf.vecs, img=img, out_t=of.apply_flow(sf.vecs, img, 't'),
    out_s=of.apply_flow(sf.vecs, img, 's'))

# ---------------------------------------------------------------- adversarial mask cases at W=1920
# (SURVEY.md finding 1: the fp32 coordinate round trip decides these bits).  Only the packed
# validity bits are stored; inputs are re-derived from the recipe below by the tests.
Hh, Ww = 48, 1920
tm = torch.ones(1, Hh, Ww, dtype=torch.bool)
tm[:, 10:30, 700:1100] = False
tm[:, :, ::97] = False
im1 = torch.rand(1, 1, Hh, Ww, generator=gen(60))
cases = {
    'int': torch.tensor([7.0, -3.0]).view(1, 2, 1, 1).expand(1, 2, Hh, Ww).contiguous(),
    'half': (torch.round(smooth_flow(1, Hh, Ww, 6.0, 61) * 2) / 2),
    'smooth': smooth_flow(1, Hh, Ww, 6.0, 62),
}
for name, fl in cases.items():
    for ref in ('t', 's'):
        o, v = of.Flow(fl, ref, tm.clone()).apply(im1, tm.clone(), return_valid_area=True)
        put('adv_%s_%s' % (name, ref), valid=np.packbits(npy(v)), out_sum=o.double().sum())
        put('adv_%s_%s_v' % (name, ref),
            vt=np.packbits(npy(of.Flow(fl, ref, tm).valid_target())),
            vs=np.packbits(npy(of.Flow(fl, ref, tm).valid_source())))

np.savez_compressed(os.path.join(HERE, 'hotpath_v1.npz'), **out)
print('wrote', len(out), 'arrays,', os.path.getsize(os.path.join(HERE, 'hotpath_v1.npz')) // 1024, 'KiB')
