"""Parity of the CUDA path (public API -> torch.library op -> C ABI -> sm_100a kernels) with
(a) golden vectors produced by the unmodified reference and (b) the pinned CPU oracle on seeded inputs.

Bar (BASELINE.json north_star): boolean masks bit-exact; floats and gradients within 1e-4 relative /
1e-3 absolute in fp32.  W forward is additionally required to be bit-identical (same arithmetic)."""
import numpy as np
import pytest
import torch

from helpers import (G, adversarial_case, assert_close, assert_grad_close, assert_mask_equal, gen, golden,
                     holed_mask, smooth_flow)

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def of():
    import oflibpytorch_b200
    return oflibpytorch_b200


def cu(t):
    return t.to(DEV)


def bit_equal(a, b):
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    return a.shape == b.shape and torch.equal(a.view(torch.int32), b.view(torch.int32))


# ------------------------------------------------------------------------------------------ golden: W
@pytest.mark.parametrize('case', ['warp_t', 'warp_t_big'])
def test_warp_t_golden(case):
    flow = cu(G(case, 'flow')).requires_grad_()
    img = cu(G(case, 'img')).requires_grad_()
    out = of().apply_flow(flow, img, 't')
    assert bit_equal(out, G(case, 'out')), 'W forward must be bit-identical to the reference CPU result'
    out.backward(cu(G(case, 'gout')))
    assert_grad_close(flow.grad, G(case, 'gflow'), 'gflow')
    assert_grad_close(img.grad, G(case, 'gimg'), 'gimg')


# ------------------------------------------------------------------------------------------ golden: S
def test_grid_from_unstructured_golden():
    x, y, data = (cu(G('gfud', k)).requires_grad_() for k in ('x', 'y', 'data'))
    o, d = of().grid_from_unstructured_data(x, y, data, cu(G('gfud', 'mask')))
    assert_close(o, G('gfud', 'out'), 'out')
    assert_close(d, G('gfud', 'dens'), 'density')
    (o * cu(G('gfud', 'gout'))).sum().add((d * cu(G('gfud', 'gdens'))).sum()).backward()
    assert_grad_close(x.grad, G('gfud', 'gx'), 'gx')
    assert_grad_close(y.grad, G('gfud', 'gy'), 'gy')
    assert_grad_close(data.grad, G('gfud', 'gdata'), 'gdata')
    x, y, data = (cu(G('gfud', k)).requires_grad_() for k in ('x', 'y', 'data'))
    o, d = of().grid_from_unstructured_data(x, y, data)
    assert_close(o, G('gfud_nomask', 'out'), 'out')
    assert_close(d, G('gfud_nomask', 'dens'), 'density')
    (o * cu(G('gfud', 'gout'))).sum().backward()
    assert_grad_close(x.grad, G('gfud_nomask', 'gx'), 'gx')
    assert_grad_close(y.grad, G('gfud_nomask', 'gy'), 'gy')
    assert_grad_close(data.grad, G('gfud_nomask', 'gdata'), 'gdata')


def test_endpoints_golden():
    from oflibpytorch_b200.utils import get_flow_endpoints
    fl = cu(G('endpoints', 'flow'))
    for ref in 'st':
        x, y = get_flow_endpoints(fl, ref)
        assert bit_equal(x, G('endpoints', 'x' + ref)) and bit_equal(y, G('endpoints', 'y' + ref))


@pytest.mark.parametrize('occ', [1, 0])
def test_apply_s_flow_golden(occ):
    c = 'apply_s_flow_occ%d' % occ
    flow, img = cu(G(c, 'flow')).requires_grad_(), cu(G(c, 'img')).requires_grad_()
    o, wm = of().apply_s_flow(flow, img, cu(G(c, 'mask')), bool(occ))
    assert_close(o, G(c, 'out'), 'out')
    assert_mask_equal(wm, G(c, 'wmask'), 'warped mask')
    o.backward(cu(G(c, 'gout')))
    assert_grad_close(flow.grad, G(c, 'gflow'), 'gflow')
    assert_grad_close(img.grad, G(c, 'gimg'), 'gimg')


def test_apply_flow_s_nomask_golden():
    c = 'apply_s_flow_occ1'
    flow, img = cu(G(c, 'flow')).requires_grad_(), cu(G(c, 'img')).requires_grad_()
    o = of().apply_flow(flow, img, 's')
    assert_close(o, G('apply_flow_s_nomask', 'out'), 'out')
    o.backward(cu(G(c, 'gout')))
    assert_grad_close(flow.grad, G('apply_flow_s_nomask', 'gflow'), 'gflow')
    assert_grad_close(img.grad, G('apply_flow_s_nomask', 'gimg'), 'gimg')


# ------------------------------------------------------------------------------------------ golden: Flow.apply
@pytest.mark.parametrize('ref', ['t', 's'])
def test_flow_apply_golden(ref):
    Flow = of().Flow
    c0 = 'flow_apply_%s_cm1' % ref
    flow, fmask, tmask, img = (cu(G(c0, k)) for k in ('flow', 'fmask', 'tmask', 'img'))
    for cm in (1, 0):
        c = 'flow_apply_%s_cm%d' % (ref, cm)
        f2, i2 = flow.clone().requires_grad_(), img.clone().requires_grad_()
        o, v = Flow(f2, ref, fmask).apply(i2, tmask, return_valid_area=True, consider_mask=bool(cm))
        assert_close(o, G(c, 'out'), c + ' out')
        assert_mask_equal(v, G(c, 'valid'), c + ' valid')
        o.backward(cu(G(c, 'gout')))
        assert_grad_close(f2.grad, G(c, 'gflow'), c + ' gflow')
        assert_grad_close(i2.grad, G(c, 'gimg'), c + ' gimg')
    fl = Flow(flow, ref, fmask)
    assert_close(fl.apply(img), G('flow_apply_%s_plain' % ref, 'out'), 'plain')
    o, v = Flow(flow[:1], ref, fmask[:1]).apply(img, return_valid_area=True)
    assert_close(o, G('flow_apply_%s_bc1' % ref, 'out'), 'bc1')
    assert_mask_equal(v, G('flow_apply_%s_bc1' % ref, 'valid'), 'bc1')
    o, v = fl.apply(img[0], tmask[0], return_valid_area=True)
    assert_close(o, G('flow_apply_%s_bc2' % ref, 'out'), 'bc2')
    assert_mask_equal(v, G('flow_apply_%s_bc2' % ref, 'valid'), 'bc2')
    assert_close(fl.apply(img[0, 0]), G('flow_apply_%s_bc3' % ref, 'out'), 'bc3')
    img8 = cu(G('flow_apply_%s_u8' % ref, 'img8'))
    o8 = fl.apply(img8)
    assert o8.dtype == torch.float32                      # PURE_PYTORCH keeps floats (flow_class.py:947)
    assert float((o8.cpu() - G('flow_apply_%s_u8' % ref, 'out')).abs().max()) <= 1.0
    assert float((o8.cpu() != G('flow_apply_%s_u8' % ref, 'out')).float().mean()) < 1e-3
    o8 = of().apply_flow(flow, img8, ref, fmask)
    assert o8.dtype == torch.uint8
    assert float((o8.cpu().int() - G('apply_flow_%s_u8' % ref, 'out').int()).abs().max()) <= 1
    r = fl.apply(Flow(cu(G('flow_apply_%s_flow' % ref, 'flow_b')), ref, cu(G('flow_apply_%s_flow' % ref, 'mask_b'))))
    assert_close(r.vecs, G('flow_apply_%s_flow' % ref, 'vecs'), 'flow target vecs')
    assert_mask_equal(r.mask, G('flow_apply_%s_flow' % ref, 'mask'), 'flow target mask')
    assert r.ref == ref
    for cut in (1, 0):
        c = 'flow_apply_%s_pad_cut%d' % (ref, cut)
        bigm = cu(G(c, 'bigm'))
        keep = bigm.clone()
        o, v = fl.apply(cu(G(c, 'big')), bigm, return_valid_area=True,
                        padding=[int(p) for p in G(c, 'pad')], cut=bool(cut))
        assert_close(o, G(c, 'out'), c)
        assert_mask_equal(v, G(c, 'valid'), c)
        assert torch.equal(bigm, keep), 'the caller\'s target_mask must not be modified'


@pytest.mark.parametrize('ref', ['t', 's'])
def test_switch_invert_valid_golden(ref):
    m = of()
    flow, fmask, gv = (cu(G('switch_ref_t', k)) for k in ('flow', 'fmask', 'gv'))
    f2 = flow.clone().requires_grad_()
    r = m.Flow(f2, ref, fmask).switch_ref()
    assert_close(r.vecs, G('switch_ref_%s' % ref, 'vecs'), 'switch vecs')
    assert_mask_equal(r.mask, G('switch_ref_%s' % ref, 'mask'), 'switch mask')
    assert r.ref != ref
    r.vecs.backward(gv)
    assert_grad_close(f2.grad, G('switch_ref_%s' % ref, 'gflow'), 'switch gflow')
    f2 = flow.clone().requires_grad_()
    r = m.Flow(f2, ref, fmask).invert()
    assert_close(r.vecs, G('invert_%s' % ref, 'vecs'), 'invert vecs')
    assert_mask_equal(r.mask, G('invert_%s' % ref, 'mask'), 'invert mask')
    assert r.ref == ref
    r.vecs.backward(gv)
    assert_grad_close(f2.grad, G('invert_%s' % ref, 'gflow'), 'invert gflow')
    for cm in (1, 0):
        fl = m.Flow(flow, ref, fmask)
        assert_mask_equal(fl.valid_target(bool(cm)), G('valid_%s_cm%d' % (ref, cm), 'vt'), 'valid_target')
        assert_mask_equal(fl.valid_source(bool(cm)), G('valid_%s_cm%d' % (ref, cm), 'vs'), 'valid_source')
    assert_close(m.switch_flow_ref(flow, ref), G('tensor_api_%s' % ref, 'sw'), 'switch_flow_ref')
    assert_close(m.invert_flow(flow, ref), G('tensor_api_%s' % ref, 'inv'), 'invert_flow')
    assert_mask_equal(m.valid_target(flow, ref), G('tensor_api_%s' % ref, 'vt'), 'valid_target()')
    assert_mask_equal(m.valid_source(flow, ref), G('tensor_api_%s' % ref, 'vs'), 'valid_source()')
    assert m.switch_flow_ref(flow[0], ref).shape == flow[0].shape


@pytest.mark.parametrize('ref', ['t', 's'])
@pytest.mark.parametrize('mode', [1, 2, 3])
def test_combine_golden(ref, mode):
    m = of()
    fa, fb, ma, mb, gv = (cu(G('combine_in', k)) for k in ('fa', 'fb', 'ma', 'mb', 'gv'))
    a, b = fa.clone().requires_grad_(), fb.clone().requires_grad_()
    r = m.Flow(a, ref, ma).combine_with(m.Flow(b, ref, mb), mode)
    c = 'combine_%s_%d' % (ref, mode)
    assert_close(r.vecs, G(c, 'vecs'), c + ' vecs')
    assert_mask_equal(r.mask, G(c, 'mask'), c + ' mask')
    assert r.ref == ref
    r.vecs.backward(gv)
    assert_grad_close(a.grad, G(c, 'ga'), c + ' ga')
    assert_grad_close(b.grad, G(c, 'gb'), c + ' gb')
    assert_close(m.combine_flows(fa, fb, mode, ref), G('combine_flows_%s_%d' % (ref, mode), 'vecs'), 'combine_flows')


def test_sintel_config1_golden():
    flow, img = cu(G('sintel', 'flow')), cu(G('sintel', 'img'))
    assert bit_equal(of().apply_flow(flow, img, 't'), G('sintel', 'out_t'))
    assert_close(of().apply_flow(flow, img, 's'), G('sintel', 'out_s'), 'sintel s')


@pytest.mark.parametrize('name', ['int', 'half', 'smooth'])
@pytest.mark.parametrize('ref', ['t', 's'])
def test_adversarial_masks_w1920_golden(name, ref):
    Flow = of().Flow
    fl, tm, im1 = (cu(t) for t in adversarial_case(name))
    o, v = Flow(fl, ref, tm).apply(im1, tm, return_valid_area=True)
    c = 'adv_%s_%s' % (name, ref)
    assert np.array_equal(np.packbits(v.cpu().numpy()), golden()[c + '/valid'])
    assert abs(float(o.double().sum()) - float(golden()[c + '/out_sum'])) < 1e-2
    f = Flow(fl, ref, tm)
    assert np.array_equal(np.packbits(f.valid_target().cpu().numpy()), golden()[c + '_v/vt'])
    assert np.array_equal(np.packbits(f.valid_source().cpu().numpy()), golden()[c + '_v/vs'])


# ------------------------------------------------------------------------------------------ oracle, seeded, larger
def _oracle():
    from oracle import oracle as orc
    return orc


@pytest.mark.parametrize('shape', [(2, 3, 270, 480), (1, 4, 131, 77), (3, 1, 64, 1920)])
@pytest.mark.parametrize('ref', ['t', 's'])
def test_flow_apply_vs_oracle(shape, ref):
    n, c, h, w = shape
    orc = _oracle()
    flow = smooth_flow(n, h, w, 5.0, 100 + h)
    flow[:, :, h // 3:h // 3 + 9, w // 4:w // 4 + 20] = 0.0
    fmask, tmask = holed_mask(n, h, w, 101), holed_mask(n, h, w, 102)
    img = torch.rand(n, c, h, w, generator=gen(103))
    gout = torch.rand(n, c, h, w, generator=gen(104))
    fo, io = flow.clone().requires_grad_(), img.clone().requires_grad_()
    oo, vo = orc.OFlow(fo, ref, fmask).apply(io, tmask, return_valid_area=True)
    oo.backward(gout)
    fg, ig = cu(flow).requires_grad_(), cu(img).requires_grad_()
    og, vg = of().Flow(fg, ref, cu(fmask)).apply(ig, cu(tmask), return_valid_area=True)
    og.backward(cu(gout))
    if ref == 't':
        assert bit_equal(og, oo)
    assert_close(og, oo, 'out')
    assert_mask_equal(vg, vo, 'valid')
    assert_grad_close(fg.grad, fo.grad, 'gflow')
    assert_grad_close(ig.grad, io.grad, 'gimg')


@pytest.mark.parametrize('ref', ['t', 's'])
@pytest.mark.parametrize('mode', [1, 2, 3])
def test_combine_vs_oracle(ref, mode):
    n, h, w = 2, 150, 200
    orc = _oracle()
    fa, fb = smooth_flow(n, h, w, 4.0, 200), smooth_flow(n, h, w, 4.0, 201)
    ma, mb = holed_mask(n, h, w, 202), holed_mask(n, h, w, 203)
    gv = torch.rand(n, 2, h, w, generator=gen(204))
    ao, bo = fa.clone().requires_grad_(), fb.clone().requires_grad_()
    ro = orc.OFlow(ao, ref, ma).combine_with(orc.OFlow(bo, ref, mb), mode)
    ro.vecs.backward(gv)
    ag, bg = cu(fa).requires_grad_(), cu(fb).requires_grad_()
    rg = of().Flow(ag, ref, cu(ma)).combine_with(of().Flow(bg, ref, cu(mb)), mode)
    rg.vecs.backward(cu(gv))
    assert_close(rg.vecs, ro.vecs, 'vecs')
    assert_mask_equal(rg.mask, ro.mask, 'mask')
    assert_grad_close(ag.grad, ao.grad, 'ga')
    assert_grad_close(bg.grad, bo.grad, 'gb')


@pytest.mark.parametrize('ref', ['t', 's'])
def test_switch_ref_vs_oracle_4k_band(ref):
    """A 4K-wide band (W = 3840): the coordinate round trip error doubles (SURVEY.md finding 1)."""
    n, h, w = 1, 96, 3840
    orc = _oracle()
    fl = smooth_flow(n, h, w, 6.0, 300)
    m = holed_mask(n, h, w, 301)
    ro = orc.OFlow(fl, ref, m).switch_ref()
    rg = of().Flow(cu(fl), ref, cu(m)).switch_ref()
    assert_close(rg.vecs, ro.vecs, 'vecs')
    assert_mask_equal(rg.mask, ro.mask, 'mask')
    fo, fg = orc.OFlow(fl, ref, m), of().Flow(cu(fl), ref, cu(m))
    assert_mask_equal(fg.valid_target(), fo.valid_target(), 'valid_target')
    assert_mask_equal(fg.valid_source(), fo.valid_source(), 'valid_source')
    c3o = fo.combine_with(orc.OFlow(smooth_flow(n, h, w, 6.0, 302), ref, m), 3)
    c3g = fg.combine_with(of().Flow(cu(smooth_flow(n, h, w, 6.0, 302)), ref, cu(m)), 3)
    assert_close(c3g.vecs, c3o.vecs, 'combine3 vecs')
    assert_mask_equal(c3g.mask, c3o.mask, 'combine3 mask')


# ------------------------------------------------------------------------------------------ edge cases
def test_zero_flow_early_outs():
    m = of()
    img = torch.rand(2, 3, 16, 24, device=DEV)
    z = torch.zeros(2, 2, 16, 24, device=DEV)
    assert m.apply_flow(z, img, 't') is img                       # utils.py:497-498 returns the target itself
    assert m.apply_flow(z + 5e-4, img, 's') is img                # thresholded
    fl = m.Flow(z, 's')
    assert fl.switch_ref().ref == 't'
    other = m.Flow(torch.rand(2, 2, 16, 24, device=DEV), 's')
    assert fl.combine_with(other, 3) is other                      # flow_class.py:1737
    assert other.combine_with(fl, 3) is other                      # flow_class.py:1744
    assert bool(m.is_zero_flow(z).all()) and not bool(m.is_zero_flow(other.vecs).any())
    assert bool(fl.is_zero().all())
    masked = m.Flow(other.vecs, 's', torch.zeros(2, 16, 24, dtype=torch.bool, device=DEV))
    assert bool(masked.is_zero().all()) and not bool(masked.is_zero(masked=False).any())


def test_error_contract_cuda():
    m = of()
    bad = torch.rand(1, 2, 8, 8, device=DEV)
    bad[0, 0, 3, 3] = float('nan')
    with pytest.raises(ValueError):
        m.Flow(bad)
    with pytest.raises(ValueError):
        m.apply_flow(bad, torch.rand(8, 8, device=DEV), 't')
    fl = m.Flow(torch.rand(1, 2, 8, 8, device=DEV))
    with pytest.raises(TypeError):
        fl.apply(torch.rand(8, 8, device=DEV), target_mask=torch.ones(8, 8, device=DEV))   # not bool
    with pytest.raises(ValueError):
        fl.apply(torch.rand(9, 8, device=DEV))


def test_odd_width_scalar_path():
    """W % 4 != 0 takes the scalar kernels."""
    orc = _oracle()
    n, c, h, w = 2, 2, 33, 45
    flow, img = smooth_flow(n, h, w, 3.0, 400), torch.rand(n, c, h, w, generator=gen(401))
    for ref in 'ts':
        oo, vo = orc.OFlow(flow, ref).apply(img, return_valid_area=True)
        og, vg = of().Flow(cu(flow), ref).apply(cu(img), return_valid_area=True)
        assert_close(og, oo, 'out ' + ref)
        assert_mask_equal(vg, vo, 'valid ' + ref)


# ------------------------------------------------------------------------------------------ arithmetic self-tests
@pytest.mark.parametrize('d', [1919, 1079, 3839, 2159, 1023, 435, 52, 36, 19, 9, 479, 199, 3, 1, 7679, 4319])
def test_round_trip_exhaustive(d):
    """The kernels evaluate the reference's coordinate round trip ((2c/(W-1) - 1) + 1)/2*(W-1) with a precomputed
    reciprocal and two FMA correction steps instead of an IEEE division; the result must be bit-identical for EVERY
    float c (all 2^32 bit patterns), because the validity masks depend on it (SURVEY.md finding 1)."""
    import ctypes
    from oflibpytorch_b200 import _native
    bad = torch.zeros(1, dtype=torch.int64, device=DEV)
    _native.check(_native.lib().ofb_selftest_div(float(d), 0, 0x80000000, ctypes.c_void_p(bad.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'selftest')
    torch.cuda.synchronize()
    assert int(bad.item()) == 0, '%d mismatching round trips for size-1 = %d' % (int(bad.item()), d)


def test_host_pipeline_matches_direct_call():
    """oflibpytorch_b200.streaming.HostWarpPipeline (chunked, three streams) == one direct call on the whole batch."""
    from oflibpytorch_b200.streaming import HostWarpPipeline
    n, c, h, w = 5, 3, 72, 96                        # 5 items in chunks of 2: ragged last chunk
    flow = smooth_flow(n, h, w, 4.0, 800)
    img = torch.rand(n, c, h, w, generator=gen(801))
    gout = torch.rand(n, c, h, w, generator=gen(802))
    fmask, tmask = holed_mask(n, h, w, 803), holed_mask(n, h, w, 804)
    pipe = HostWarpPipeline(torch.device('cuda', 0), chunk=2)
    pin = [t.pin_memory() for t in (flow, img, fmask, tmask, gout)]
    for _ in range(2):                               # second run reuses the device buffers
        out_h, valid_h, gf_h, gi_h = pipe.run(pin[0], pin[1], 't', pin[2], pin[3], pin[4])
        torch.cuda.synchronize()
    fg, ig = cu(flow).requires_grad_(), cu(img).requires_grad_()
    og, vg = of().Flow(fg, 't', cu(fmask)).apply(ig, cu(tmask), return_valid_area=True)
    og.backward(cu(gout))
    assert bit_equal(out_h, og) and torch.equal(valid_h, vg.cpu())
    assert_grad_close(gf_h, fg.grad, 'gflow')
    assert_grad_close(gi_h, ig.grad, 'gimg')
    out2, valid2, g1, g2 = pipe.run(pin[0], pin[1], 't', pin[2], pin[3])      # forward only
    torch.cuda.synchronize()
    assert g1 is None and g2 is None and bit_equal(out2, og) and torch.equal(valid2, vg.cpu())


def test_host_pipeline_source_reference_flows():
    """the same pipeline with ref 's' (splat path): chunked result == one direct call; masks bit-exact"""
    from oflibpytorch_b200.streaming import HostWarpPipeline
    n, c, h, w = 5, 3, 64, 80
    flow = smooth_flow(n, h, w, 4.0, 810)
    img = torch.rand(n, c, h, w, generator=gen(811))
    gout = torch.rand(n, c, h, w, generator=gen(812))
    fmask, tmask = holed_mask(n, h, w, 813), holed_mask(n, h, w, 814)
    pipe = HostWarpPipeline(torch.device('cuda', 0), chunk=3)
    pin = [t.pin_memory() for t in (flow, img, fmask, tmask, gout)]
    out_h, valid_h, gf_h, gi_h = pipe.run(pin[0], pin[1], 's', pin[2], pin[3], pin[4])
    torch.cuda.synchronize()
    fg, ig = cu(flow).requires_grad_(), cu(img).requires_grad_()
    og, vg = of().Flow(fg, 's', cu(fmask)).apply(ig, cu(tmask), return_valid_area=True)
    og.backward(cu(gout))
    assert_close(out_h, og, 'out')
    assert torch.equal(valid_h, vg.cpu())
    assert_grad_close(gf_h, fg.grad, 'gflow')
    assert_grad_close(gi_h, ig.grad, 'gimg')
