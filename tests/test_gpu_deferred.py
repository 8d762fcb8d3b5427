"""'deferred' validation: the reference's zero-flow early-outs (utils.py:497-498, flow_class.py:1729-1744) and the
finite test are decided on the device.  Every result must equal what 'eager' (host-decided, reference semantics)
returns on the same inputs, including the all-zero corner cases, and the gradients must agree."""
import pytest
import torch

from helpers import assert_grad_close, assert_mask_equal, gen, holed_mask, smooth_flow

pytestmark = pytest.mark.gpu


def _flows(kind, n, h, w, seed):
    if kind == 'zero':
        return torch.zeros(n, 2, h, w)
    if kind == 'tiny':                       # below the 1e-3 threshold but not exactly zero
        return torch.rand(n, 2, h, w, generator=gen(seed)) * 5e-4
    if kind == 'masked_zero':                # zero wherever the mask is set, non-zero in the holes
        f = smooth_flow(n, h, w, 3.0, seed)
        return f * (~holed_mask(n, h, w, 77)).unsqueeze(1)
    return smooth_flow(n, h, w, 3.0, seed)


@pytest.mark.parametrize('kind', ['smooth', 'zero', 'tiny', 'masked_zero'])
@pytest.mark.parametrize('with_masks', [False, True])
def test_apply_t_deferred_equals_eager(kind, with_masks):
    import oflibpytorch_b200 as of
    n, h, w = 3, 40, 72
    flow = _flows(kind, n, h, w, 5).cuda()
    img = torch.rand(n, 3, h, w, generator=gen(6)).cuda()
    gout = torch.rand(n, 3, h, w, generator=gen(7)).cuda()
    m = holed_mask(n, h, w, 77).cuda() if with_masks else None
    tm = holed_mask(n, h, w, 78).cuda() if with_masks else None
    res = {}
    for mode in ('eager', 'deferred'):
        with of.validation(mode):
            f = flow.clone().requires_grad_()
            i = img.clone().requires_grad_()
            out, valid = of.Flow(f, 't', m).apply(i, tm, return_valid_area=True)
            gf, gi = torch.autograd.grad(out, (f, i), gout, allow_unused=True)
            res[mode] = (out.detach(), valid, torch.zeros_like(f) if gf is None else gf, gi)
    of.check_deferred()
    assert torch.equal(res['eager'][0], res['deferred'][0])
    assert_mask_equal(res['eager'][1], res['deferred'][1], kind)
    assert_grad_close(res['deferred'][2], res['eager'][2], 'gflow ' + kind)
    assert_grad_close(res['deferred'][3], res['eager'][3], 'gimg ' + kind)


@pytest.mark.parametrize('ref', ['t', 's'])
@pytest.mark.parametrize('kinds', [('smooth', 'smooth'), ('zero', 'smooth'), ('smooth', 'zero'), ('zero', 'zero'),
                                   ('masked_zero', 'smooth'), ('smooth', 'masked_zero'), ('tiny', 'smooth'),
                                   ('smooth', 'tiny')])
def test_combine3_deferred_equals_eager(ref, kinds):
    import oflibpytorch_b200 as of
    n, h, w = 2, 36, 64
    fa, fb = _flows(kinds[0], n, h, w, 11).cuda(), _flows(kinds[1], n, h, w, 12).cuda()
    ma, mb = holed_mask(n, h, w, 77).cuda(), holed_mask(n, h, w, 79).cuda()
    gout = torch.rand(n, 2, h, w, generator=gen(13)).cuda()
    res = {}
    for mode in ('eager', 'deferred'):
        with of.validation(mode):
            a = fa.clone().requires_grad_()
            b = fb.clone().requires_grad_()
            r = of.Flow(a, ref, ma).combine_with(of.Flow(b, ref, mb), 3)
            ga, gb = torch.autograd.grad(r.vecs, (a, b), gout, allow_unused=True)
            # chained once more, so that the flags the kernel attached to its result are exercised
            r2 = r.combine_with(of.Flow(fb, ref, mb), 3)
            res[mode] = (r.vecs.detach(), r.mask, torch.zeros_like(a) if ga is None else ga,
                         torch.zeros_like(b) if gb is None else gb, r2.vecs.detach(), r2.mask, r.is_zero(), r2.is_zero())
    of.check_deferred()
    e, d = res['eager'], res['deferred']
    assert torch.equal(e[0], d[0]), kinds
    assert_mask_equal(e[1], d[1], str(kinds))
    assert_grad_close(d[2], e[2], 'g_a ' + str(kinds))
    assert_grad_close(d[3], e[3], 'g_b ' + str(kinds))
    assert torch.equal(e[4], d[4]), kinds
    assert_mask_equal(e[5], d[5], str(kinds))
    assert torch.equal(e[6], d[6]) and torch.equal(e[7], d[7])


def test_valid_area_and_apply_flow_deferred():
    import oflibpytorch_b200 as of
    n, h, w = 2, 32, 48
    for kind in ('smooth', 'zero'):
        flow = _flows(kind, n, h, w, 21).cuda()
        m = holed_mask(n, h, w, 77).cuda()
        img = torch.rand(n, 3, h, w, generator=gen(22)).cuda()
        with of.validation('eager'):
            vt, vs, o = of.Flow(flow, 't', m).valid_target(), of.Flow(flow, 's', m).valid_source(), of.apply_flow(flow, img, 't')
        with of.validation('deferred'):
            vt2, vs2, o2 = of.Flow(flow, 't', m).valid_target(), of.Flow(flow, 's', m).valid_source(), of.apply_flow(flow, img, 't')
        assert_mask_equal(vt, vt2, kind)
        assert_mask_equal(vs, vs2, kind)
        assert torch.equal(o, o2)
    of.check_deferred()


def test_deferred_nonfinite_raises_at_check():
    import oflibpytorch_b200 as of
    flow = smooth_flow(1, 16, 24, 2.0, 3).cuda()
    flow[0, 1, 3, 4] = float('nan')
    img = torch.rand(1, 1, 16, 24).cuda()
    with of.validation('eager'):
        with pytest.raises(ValueError):
            of.Flow(flow, 't')
    with of.validation('deferred'):
        of.Flow(flow, 't').apply(img)           # no exception here ...
        with pytest.raises(ValueError):
            of.check_deferred()                 # ... but at the deferred check
    of.check_deferred()                         # the queue is empty again


def test_deferred_chain_does_not_synchronise():
    """cfg 5 in miniature: chained combine_with(3) + apply under 'deferred' issues no blocking host read."""
    import oflibpytorch_b200 as of
    h, w = 64, 96
    flows = [smooth_flow(1, h, w, 2.0, 100 + i).cuda() for i in range(4)]
    img = torch.rand(1, 3, h, w, generator=gen(5)).cuda()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode('error')
    try:
        with of.validation('deferred'):
            acc = None
            for f in flows:
                cur = of.Flow(f.requires_grad_(), 't')
                acc = cur if acc is None else acc.combine_with(cur, 3)
                out, valid = acc.apply(img, return_valid_area=True)
            out.sum().backward()
    finally:
        torch.cuda.set_sync_debug_mode('default')
    of.check_deferred()
