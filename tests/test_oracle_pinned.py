"""Pins the CPU oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) -- runs without a GPU."""
import numpy as np
import pytest
import torch

from helpers import G, assert_close, assert_grad_close, assert_mask_equal, golden, smooth_flow, gen
from oracle import oracle as orc


def bit_equal(a, b):
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


@pytest.mark.parametrize('case', ['warp_t', 'warp_t_big'])
def test_warp_t_bit_exact_and_grads(case):
    flow = G(case, 'flow').requires_grad_()
    img = G(case, 'img').requires_grad_()
    out = orc.apply_flow(flow, img, 't')
    assert bit_equal(out.detach(), G(case, 'out')), 'W forward must be bit-identical to ATen CPU'
    out.backward(G(case, 'gout'))
    assert_grad_close(flow.grad, G(case, 'gflow'), 'gflow', rel=2e-6)
    assert_grad_close(img.grad, G(case, 'gimg'), 'gimg', rel=2e-6)


def test_grid_from_unstructured_bit_exact():
    x, y, data = G('gfud', 'x').requires_grad_(), G('gfud', 'y').requires_grad_(), G('gfud', 'data').requires_grad_()
    o, d = orc.grid_from_unstructured_data(x, y, data, G('gfud', 'mask'))
    assert bit_equal(o.detach(), G('gfud', 'out'))
    assert bit_equal(d.detach(), G('gfud', 'dens'))
    (o * G('gfud', 'gout')).sum().add((d * G('gfud', 'gdens')).sum()).backward()
    assert_grad_close(x.grad, G('gfud', 'gx'), 'gx', rel=1e-5)
    assert_grad_close(y.grad, G('gfud', 'gy'), 'gy', rel=1e-5)
    assert_grad_close(data.grad, G('gfud', 'gdata'), 'gdata', rel=1e-5)
    x, y, data = G('gfud', 'x').requires_grad_(), G('gfud', 'y').requires_grad_(), G('gfud', 'data').requires_grad_()
    o, d = orc.grid_from_unstructured_data(x, y, data)
    assert bit_equal(o.detach(), G('gfud_nomask', 'out'))
    assert bit_equal(d.detach(), G('gfud_nomask', 'dens'))
    (o * G('gfud', 'gout')).sum().backward()
    assert_grad_close(x.grad, G('gfud_nomask', 'gx'), 'gx', rel=1e-5)
    assert_grad_close(data.grad, G('gfud_nomask', 'gdata'), 'gdata', rel=1e-5)


def test_endpoints_exact():
    fl = G('endpoints', 'flow')
    for ref in 'st':
        x, y = orc.get_flow_endpoints(fl, ref)
        assert bit_equal(x, G('endpoints', 'x' + ref)) and bit_equal(y, G('endpoints', 'y' + ref))
    # the raw C restatement too
    n, _, h, w = fl.shape
    x = torch.empty(n, h, w)
    y = torch.empty(n, h, w)
    orc.lib().orc_flow_endpoints(orc._p(fl), 1.0, orc._p(x), orc._p(y), n, h, w)
    assert bit_equal(x, G('endpoints', 'xs')) and bit_equal(y, G('endpoints', 'ys'))


@pytest.mark.parametrize('occ', [1, 0])
def test_apply_s_flow(occ):
    c = 'apply_s_flow_occ%d' % occ
    flow, img = G(c, 'flow').requires_grad_(), G(c, 'img').requires_grad_()
    o, wm = orc.apply_s_flow(flow, img, G(c, 'mask'), bool(occ))
    assert bit_equal(o.detach(), G(c, 'out'))
    assert_mask_equal(wm, G(c, 'wmask'))
    o.backward(G(c, 'gout'))
    assert_grad_close(flow.grad, G(c, 'gflow'), 'gflow', rel=1e-5)
    assert_grad_close(img.grad, G(c, 'gimg'), 'gimg', rel=1e-5)


def test_apply_flow_s_nomask():
    flow, img = G('apply_s_flow_occ1', 'flow').requires_grad_(), G('apply_s_flow_occ1', 'img').requires_grad_()
    o = orc.apply_flow(flow, img, 's')
    assert bit_equal(o.detach(), G('apply_flow_s_nomask', 'out'))
    o.backward(G('apply_s_flow_occ1', 'gout'))
    assert_grad_close(flow.grad, G('apply_flow_s_nomask', 'gflow'), 'gflow', rel=1e-5)


@pytest.mark.parametrize('ref', ['t', 's'])
def test_flow_apply_variants(ref):
    c0 = 'flow_apply_%s_cm1' % ref
    flow, fmask, tmask, img = G(c0, 'flow'), G(c0, 'fmask'), G(c0, 'tmask'), G(c0, 'img')
    for cm in (1, 0):
        c = 'flow_apply_%s_cm%d' % (ref, cm)
        f2, i2 = flow.clone().requires_grad_(), img.clone().requires_grad_()
        o, v = orc.OFlow(f2, ref, fmask).apply(i2, tmask, return_valid_area=True, consider_mask=bool(cm))
        assert bit_equal(o.detach(), G(c, 'out'))
        assert_mask_equal(v, G(c, 'valid'))
        o.backward(G(c, 'gout'))
        assert_grad_close(f2.grad, G(c, 'gflow'), 'gflow', rel=1e-5)
        assert_grad_close(i2.grad, G(c, 'gimg'), 'gimg', rel=1e-5)
    fl = orc.OFlow(flow, ref, fmask)
    assert bit_equal(fl.apply(img), G('flow_apply_%s_plain' % ref, 'out'))
    o, v = orc.OFlow(flow[:1], ref, fmask[:1]).apply(img, return_valid_area=True)
    assert bit_equal(o, G('flow_apply_%s_bc1' % ref, 'out'))
    assert_mask_equal(v, G('flow_apply_%s_bc1' % ref, 'valid'))
    o, v = fl.apply(img[0], tmask[0], return_valid_area=True)
    assert bit_equal(o, G('flow_apply_%s_bc2' % ref, 'out'))
    assert_mask_equal(v, G('flow_apply_%s_bc2' % ref, 'valid'))
    assert bit_equal(fl.apply(img[0, 0]), G('flow_apply_%s_bc3' % ref, 'out'))
    img8 = G('flow_apply_%s_u8' % ref, 'img8')
    assert torch.equal(fl.apply(img8), G('flow_apply_%s_u8' % ref, 'out'))
    assert torch.equal(orc.apply_flow(flow, img8, ref, fmask), G('apply_flow_%s_u8' % ref, 'out'))
    r = fl.apply(orc.OFlow(G('flow_apply_%s_flow' % ref, 'flow_b'), ref, G('flow_apply_%s_flow' % ref, 'mask_b')))
    assert bit_equal(r.vecs, G('flow_apply_%s_flow' % ref, 'vecs'))
    assert_mask_equal(r.mask, G('flow_apply_%s_flow' % ref, 'mask'))
    for cut in (1, 0):
        c = 'flow_apply_%s_pad_cut%d' % (ref, cut)
        o, v = fl.apply(G(c, 'big'), G(c, 'bigm'), return_valid_area=True,
                        padding=[int(p) for p in G(c, 'pad')], cut=bool(cut))
        assert bit_equal(o, G(c, 'out'))
        assert_mask_equal(v, G(c, 'valid'))


@pytest.mark.parametrize('ref', ['t', 's'])
def test_switch_invert_valid(ref):
    flow, fmask, gv = G('switch_ref_t', 'flow'), G('switch_ref_t', 'fmask'), G('switch_ref_t', 'gv')
    f2 = flow.clone().requires_grad_()
    r = orc.OFlow(f2, ref, fmask).switch_ref()
    assert bit_equal(r.vecs.detach(), G('switch_ref_%s' % ref, 'vecs'))
    assert_mask_equal(r.mask, G('switch_ref_%s' % ref, 'mask'))
    assert r.ref != ref
    r.vecs.backward(gv)
    assert_grad_close(f2.grad, G('switch_ref_%s' % ref, 'gflow'), 'gflow', rel=1e-5)
    f2 = flow.clone().requires_grad_()
    r = orc.OFlow(f2, ref, fmask).invert()
    assert bit_equal(r.vecs.detach(), G('invert_%s' % ref, 'vecs'))
    assert_mask_equal(r.mask, G('invert_%s' % ref, 'mask'))
    r.vecs.backward(gv)
    assert_grad_close(f2.grad, G('invert_%s' % ref, 'gflow'), 'gflow', rel=1e-5)
    for cm in (1, 0):
        fl = orc.OFlow(flow, ref, fmask)
        assert_mask_equal(fl.valid_target(bool(cm)), G('valid_%s_cm%d' % (ref, cm), 'vt'))
        assert_mask_equal(fl.valid_source(bool(cm)), G('valid_%s_cm%d' % (ref, cm), 'vs'))
    fl = orc.OFlow(flow, ref)
    assert bit_equal(fl.switch_ref().vecs, G('tensor_api_%s' % ref, 'sw'))
    assert bit_equal(fl.invert().vecs, G('tensor_api_%s' % ref, 'inv'))
    assert_mask_equal(fl.valid_target(), G('tensor_api_%s' % ref, 'vt'))
    assert_mask_equal(fl.valid_source(), G('tensor_api_%s' % ref, 'vs'))


@pytest.mark.parametrize('ref', ['t', 's'])
@pytest.mark.parametrize('mode', [1, 2, 3])
def test_combine_with(ref, mode):
    fa, fb, ma, mb, gv = (G('combine_in', k) for k in ('fa', 'fb', 'ma', 'mb', 'gv'))
    a, b = fa.clone().requires_grad_(), fb.clone().requires_grad_()
    r = orc.OFlow(a, ref, ma).combine_with(orc.OFlow(b, ref, mb), mode)
    c = 'combine_%s_%d' % (ref, mode)
    assert bit_equal(r.vecs.detach(), G(c, 'vecs'))
    assert_mask_equal(r.mask, G(c, 'mask'))
    r.vecs.backward(gv)
    assert_grad_close(a.grad, G(c, 'ga'), 'ga', rel=2e-5)
    assert_grad_close(b.grad, G(c, 'gb'), 'gb', rel=2e-5)
    r = orc.OFlow(fa, ref).combine_with(orc.OFlow(fb, ref), mode)
    assert bit_equal(r.vecs, G('combine_flows_%s_%d' % (ref, mode), 'vecs'))


def test_sintel_config1():
    flow, img = G('sintel', 'flow'), G('sintel', 'img')
    # the bundled fixture: u = row*col, v = 0 (reference test/test_utils.py:828-833)
    h, w = flow.shape[-2:]
    rows = torch.arange(h, dtype=torch.float32)[:, None] * torch.arange(w, dtype=torch.float32)[None]
    assert torch.equal(flow[0, 0], rows) and bool((flow[0, 1] == 0).all())
    assert bit_equal(orc.apply_flow(flow, img, 't'), G('sintel', 'out_t'))
    assert bit_equal(orc.apply_flow(flow, img, 's'), G('sintel', 'out_s'))


@pytest.mark.parametrize('name', ['int', 'half', 'smooth'])
@pytest.mark.parametrize('ref', ['t', 's'])
def test_adversarial_masks_w1920(name, ref):
    """Mask bits at W=1920 depend on the fp32 coordinate round trip (SURVEY.md finding 1)."""
    from helpers import adversarial_case
    fl, tm, im1 = adversarial_case(name)
    o, v = orc.OFlow(fl, ref, tm.clone()).apply(im1, tm.clone(), return_valid_area=True)
    c = 'adv_%s_%s' % (name, ref)
    assert np.array_equal(np.packbits(v.numpy()), golden()[c + '/valid'])
    assert abs(float(o.double().sum()) - float(golden()[c + '/out_sum'])) < 1e-6
    f = orc.OFlow(fl, ref, tm)
    assert np.array_equal(np.packbits(f.valid_target().numpy()), golden()[c + '_v/vt'])
    assert np.array_equal(np.packbits(f.valid_source().numpy()), golden()[c + '_v/vs'])
