"""The reference's own known-answer tests, run against the CPU oracle (no GPU needed) and,
marked gpu, against the CUDA path.  Inputs come from tests/golden/kats_v1.npz (built by the reference's own
constructors, tests/golden/make_kats.py); expected tables are in tests/kat_tables.py."""
import os

import numpy as np
import pytest
import torch

import kat_tables as K

KATS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kats_v1.npz'))


def T(name):
    return torch.from_numpy(np.array(KATS[name]))


class OracleImpl:
    dev = 'cpu'

    def __init__(self):
        from oracle import oracle as orc
        self.o = orc

    def combine(self, a, b, mode, ref):
        return self.o.combine(a, b, mode, ref)

    def track_pts(self, flow, ref, pts, int_out=False):
        return self.o.track_pts(flow, ref, pts, int_out)

    def flow_track(self, fl, pts, status):
        return self.o.flow_track(fl, pts, False, status)

    def flow(self, vecs, ref, mask=None):
        return self.o.OFlow(vecs, ref, mask)

    def gfud(self, x, y, d):
        return self.o.grid_from_unstructured_data(x, y, d)

    def endpoints(self, f, ref):
        return self.o.get_flow_endpoints(f, ref)


class CudaImpl:
    dev = 'cuda'

    def __init__(self):
        import oflibpytorch_b200 as of
        self.of = of

    def combine(self, a, b, mode, ref):
        return a.combine(b, mode, ref)

    def track_pts(self, flow, ref, pts, int_out=False):
        return self.of.track_pts(flow, ref, pts, int_out)

    def flow_track(self, fl, pts, status):
        return fl.track(pts, get_valid_status=status)

    def flow(self, vecs, ref, mask=None):
        return self.of.Flow(vecs.cuda(), ref, None if mask is None else torch.as_tensor(mask).cuda())

    def gfud(self, x, y, d):
        return self.of.grid_from_unstructured_data(x, y, d)

    def endpoints(self, f, ref):
        return self.of.get_flow_endpoints(f, ref)


IMPLS = [pytest.param(OracleImpl, id='oracle'), pytest.param(CudaImpl, id='cuda', marks=pytest.mark.gpu)]


def eq(mask_tensor, table):
    np.testing.assert_array_equal(mask_tensor.cpu().numpy(), table)


@pytest.mark.parametrize('impl_cls', IMPLS)
def test_valid_target_7x7(impl_cls):
    """reference test/test_flow_class.py:1389-1486"""
    im = impl_cls()
    ms, mt = K.masks_7x7()
    fs, ft = T('rot45_s'), T('rot45_t')
    eq(im.flow(fs, 's').valid_target()[0], K.VT_S_PP)
    eq(im.flow(ft, 't').valid_target()[0], K.VT_T)
    eq(im.flow(fs, 's', ms).valid_target()[0], K.VT_S_MASKED_CM_PP)
    eq(im.flow(fs, 's', ms).valid_target(False)[0], K.VT_S_MASKED_PP)
    eq(im.flow(ft, 't', mt).valid_target()[0], K.VT_T_MASKED)
    # batched (test_flow_class.py:1483-1486)
    fb = im.flow(torch.cat((fs, fs)), 's', np.stack((np.ones((7, 7), bool), ms)))
    vt = fb.valid_target()
    eq(vt[0], K.VT_S_PP)
    eq(vt[1], K.VT_S_MASKED_CM_PP)
    fb = im.flow(torch.cat((ft, ft)), 't', np.stack((np.ones((7, 7), bool), mt)))
    vt = fb.valid_target()
    eq(vt[0], K.VT_T)
    eq(vt[1], K.VT_T_MASKED)


@pytest.mark.parametrize('impl_cls', IMPLS)
def test_valid_source_7x7(impl_cls):
    """reference test/test_flow_class.py:1505-1600"""
    im = impl_cls()
    ms, mt = K.masks_7x7()
    fs, ft = T('rot45_s'), T('rot45_t')
    eq(im.flow(fs, 's').valid_source()[0], K.VS_S)
    eq(im.flow(ft, 't').valid_source()[0], K.VS_T_PP)
    eq(im.flow(fs, 's', ms).valid_source()[0], K.VS_S_MASKED)
    eq(im.flow(ft, 't', mt).valid_source()[0], K.VS_T_MASKED_CM_PP)
    eq(im.flow(ft, 't', mt).valid_source(False)[0], K.VS_T_MASKED_PP)


@pytest.mark.parametrize('impl_cls', IMPLS)
def test_grid_from_unstructured_count(impl_cls):
    """reference test/test_utils.py:1104-1120: count_nonzero(density > 0.95) == 24664, inverse rotation within 5e-2"""
    im = impl_cls()
    v = torch.cat((T('rot20_s'), T('rot20_s'))).to(im.dev).requires_grad_()
    x, y = im.endpoints(v, 's')
    data, dens = im.gfud(x, y, v)
    assert data.grad_fn is not None and dens.grad_fn is not None
    m = (dens > 0.95).cpu().numpy()
    assert int(np.count_nonzero(m)) == int(KATS['rot20_count']) == 24664
    rev = torch.cat((T('rot20_rev_s'), T('rot20_rev_s'))).permute(0, 2, 3, 1).numpy()
    approx = (-data).detach().cpu().permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(rev[m], approx[m], atol=5e-2)


@pytest.mark.parametrize('impl_cls', IMPLS)
def test_flow_endpoints_kat(impl_cls):
    """reference test/test_utils.py:1082-1101"""
    im = impl_cls()
    f = T('ep_flow').to(im.dev)
    for ref in 'st':
        x, y = im.endpoints(f, ref)
        np.testing.assert_array_equal(x.cpu().numpy(), KATS['ep_x_' + ref])
        np.testing.assert_array_equal(y.cpu().numpy(), KATS['ep_y_' + ref])


def test_normalise_coords_kat():
    """reference test/test_utils.py:456-475 (host helper, CPU)"""
    from oflibpytorch_b200.utils import normalise_coords
    out = normalise_coords(T('norm_in'), (480, 640))
    np.testing.assert_array_equal(out.numpy(), KATS['norm_out'])
    with pytest.raises(ValueError):
        normalise_coords(T('norm_in'), (480, 640, 3))


@pytest.mark.parametrize('impl_cls', IMPLS)
@pytest.mark.parametrize('ref', ['s', 't'])
@pytest.mark.parametrize('mode', [1, 2, 3])
def test_combine_analytic(impl_cls, ref, mode):
    """reference test/test_flow_operations.py:41-116: f1 (+) f2 = f3 for a rotation followed by a scaling; the
    solved-for flow matches the analytic one within the combined mask (atol 0.8 for the splat-based modes), and
    equals the reference's own result within the port tolerance."""
    im = impl_cls()
    f = {k: T('cmb_f%d_%s' % (k, ref)) for k in (1, 2, 3)}
    a, b, want = {1: (2, 3, 1), 2: (1, 3, 2), 3: (1, 2, 3)}[mode]
    r = im.flow(f[a], ref).combine_with(im.flow(f[b], ref), mode)
    vecs, mask = r.vecs.detach().cpu(), r.mask.cpu()
    np.testing.assert_array_equal(mask.numpy(), KATS['cmb_%s_%d_mask' % (ref, mode)])
    np.testing.assert_allclose(vecs.numpy(), KATS['cmb_%s_%d_vecs' % (ref, mode)], rtol=1e-4, atol=1e-3)
    m = mask[0].numpy()
    assert m.sum() > 0.4 * m.size
    err = np.abs(vecs[0].numpy() - f[want][0].numpy())[:, m]
    assert float(err.max()) < 0.8


@pytest.mark.parametrize('impl_cls', IMPLS)
@pytest.mark.parametrize('mode', [1, 2, 3])
def test_combine_general_refs(impl_cls, mode):
    """Flow.combine (reference flow_class.py:1812-1939) for every combination of input and output
    references, against the reference's own results (bit-exact masks; vectors bit-exact for the oracle, within the
    port tolerance for the CUDA path)."""
    im = impl_cls()
    a, b, ma, mb = T('cmbx_a'), T('cmbx_b'), T('cmbx_ma'), T('cmbx_mb')
    for ra in 'st':
        for rb in 'st':
            for ro in 'st':
                r = im.combine(im.flow(a, ra, ma), im.flow(b, rb, mb), mode, ro)
                k = 'cmbx_%d%s%s%s' % (mode, ra, rb, ro)
                assert r.ref == ro
                np.testing.assert_array_equal(r.mask.cpu().numpy(), KATS[k + '_m'], err_msg=k)
                if im.dev == 'cpu':
                    np.testing.assert_array_equal(r.vecs.detach().numpy(), KATS[k + '_v'], err_msg=k)
                else:
                    np.testing.assert_allclose(r.vecs.detach().cpu().numpy(), KATS[k + '_v'], rtol=1e-4, atol=1e-3,
                                               err_msg=k)


@pytest.mark.parametrize('impl_cls', IMPLS)
@pytest.mark.parametrize('ref', ['s', 't'])
def test_track_pts(impl_cls, ref):
    """track_pts (reference utils.py:941-1042) and Flow.track (flow_class.py:961-1020): float points
    (interpolated), integer points (gathered), int_out, 2-D points, validity status, gradients to flow and points."""
    im = impl_cls()
    flow, pf, pi, gp = T('trk_flow').to(im.dev), T('trk_pf').to(im.dev), T('trk_pi').to(im.dev), T('trk_gp').to(im.dev)
    f2, p2 = flow.clone().requires_grad_(), pf.clone().requires_grad_()
    w = im.track_pts(f2, ref, p2)
    w.backward(gp)
    tol = dict(rtol=0, atol=0) if (im.dev == 'cpu' and ref == 's') else dict(rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(w.detach().cpu().numpy(), KATS['trk_%s_f' % ref], **tol)
    np.testing.assert_allclose(f2.grad.cpu().numpy(), KATS['trk_%s_f_gflow' % ref], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(p2.grad.cpu().numpy(), KATS['trk_%s_f_gpts' % ref], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(im.track_pts(flow, ref, pi).cpu().numpy(), KATS['trk_%s_i' % ref], rtol=1e-4, atol=1e-3)
    wi = im.track_pts(flow, ref, pf, True)
    assert wi.dtype == torch.int64
    assert int((wi.cpu() != T('trk_%s_f_int' % ref)).sum()) <= (0 if im.dev == 'cpu' else 1)
    np.testing.assert_allclose(im.track_pts(flow, ref, pf[0]).cpu().numpy(), KATS['trk_%s_f_2d' % ref], rtol=1e-4, atol=1e-3)
    wp, st = im.flow_track(im.flow(T('trk_flow'), ref, T('trk_mask')), pf, True)
    np.testing.assert_allclose(wp.detach().cpu().numpy(), KATS['trk_%s_flowtrack' % ref], rtol=1e-4, atol=1e-3)
    np.testing.assert_array_equal(st.cpu().numpy(), KATS['trk_%s_status' % ref])
