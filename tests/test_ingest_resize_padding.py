"""SURVEY.md section 8f rows f3 (get_padding) and f4 (Sintel / KITTI ingest, resize_flow / Flow.resize): the CPU
oracle against the reference's outputs in tests/golden/ingest_v1.npz (no GPU needed), and -- marked gpu -- the CUDA
path against the same golden arrays and against the oracle on fresh seeded inputs.

Bars: padding lists, masks, loaded flows and KITTI values bit-exact; resized vectors and gradients within
1e-5 absolute + 1e-5 relative; the CUDA kernel follows the oracle's operation order and is bit-identical to it."""
import os

import numpy as np
import pytest
import torch

from helpers import gen, smooth_flow

HERE = os.path.dirname(os.path.abspath(__file__))
ING = np.load(os.path.join(HERE, 'golden', 'ingest_v1.npz'))
FILES = os.path.join(HERE, 'golden', 'files')
# test/test_flow_class.py:1630-1633, the expected paddings of the reference's own test
KAT_PADDING = {'s': [5, 0, 0, 3], 't': [0, 3, 5, 0], 's_masked': [3, 0, 0, 1], 't_masked': [0, 1, 3, 0]}


def T(name):
    return torch.from_numpy(np.array(ING[name]))


def scales():
    return [tuple(float(v) for v in row) for row in ING['rsz_scales']]


# ------------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_padding_matches_reference():
    from oracle import oracle as orc
    for name, want in KAT_PADDING.items():
        fl = orc.OFlow(T('pad_rot45_%s_v' % name), name[0], T('pad_rot45_%s_m' % name))
        assert orc.get_padding(fl) == [want] == ING['pad_rot45_' + name].tolist()
    v, m = T('pad_rnd_v'), T('pad_rnd_m')
    for ref in 'st':
        assert orc.get_padding(orc.OFlow(v, ref, m)) == ING['pad_rnd_' + ref].tolist()
        assert orc.get_padding(orc.OFlow(v, ref)) == ING['pad_rnd_nomask_' + ref].tolist()
    assert ING['pad_rnd_s'][1].tolist() == [0, 0, 0, 0]          # the sub-threshold item needs no padding


def test_oracle_resize_matches_reference():
    """Vectors within 1e-5 (torch's CPU kernel contracts its FMAs differently in its vector body and its scalar tail,
    so no single operation order is bit-identical everywhere; on fields wider than one vector the restatement IS
    bit-identical, see rsz_big below); rounded masks bit-exact."""
    from oracle import oracle as orc
    v, m = T('rsz_v'), T('rsz_m')
    for i, s in enumerate(scales()):
        assert torch.allclose(orc.resize_flow(v, s), T('rsz_%d_out' % i), rtol=1e-5, atol=1e-5), 'scale %r' % (s,)
        assert torch.equal(orc.resize_mask(m, s), T('rsz_%d_mask' % i)), 'scale %r' % (s,)
        gin = orc.resize_flow_bwd(T('rsz_%d_gout' % i), s, v.shape[2], v.shape[3])
        assert torch.allclose(gin, T('rsz_%d_gin' % i), rtol=1e-5, atol=1e-5), 'scale %r' % (s,)
    assert torch.equal(orc.resize_flow(T('rsz_big_v'), 1 / 3), T('rsz_big_out'))
    assert torch.allclose(orc.resize_flow(T('rsz_v')[0], 1.5), T('rsz_3d_out'), rtol=1e-5, atol=1e-5)
    # test/test_utils.py:891-899: a field built at 1/3 of the size agrees with the resized one
    np.testing.assert_allclose(T('rsz_big_out').numpy(), T('rsz_small_v').numpy(), atol=1, rtol=.1)


def test_oracle_loaders_match_reference():
    from oracle import oracle as orc
    import cv2
    for k in range(2):
        assert torch.equal(orc.load_sintel(os.path.join(FILES, 'sample%d.flo' % k)), T('flo%d' % k))
        png = cv2.imread(os.path.join(FILES, 'sample%d_kitti.png' % k), cv2.IMREAD_UNCHANGED)
        assert torch.equal(orc.decode_kitti(png), T('kitti%d' % k))
    with pytest.raises(ValueError):
        orc.load_sintel(os.path.join(FILES, 'wrong_tag.flo'))


def test_host_validation_of_scale_and_paths():
    """the reference's argument errors (test/test_utils.py:820-853, 901-913) are raised before any kernel"""
    import oflibpytorch_b200 as of
    from oflibpytorch_b200.utils import _valid_scale
    assert _valid_scale(2) == [2.0, 2.0] and _valid_scale((2, .5)) == [2.0, 0.5]
    with pytest.raises(TypeError):
        _valid_scale('test')
    for bad in (['test', 0], [1, 2, 3], 0, -0.1):
        with pytest.raises(ValueError):
            _valid_scale(bad)
    with pytest.raises(TypeError):
        of.load_sintel(0)
    with pytest.raises(ValueError):
        of.load_sintel(os.path.join(FILES, 'wrong_tag.flo'))
    with pytest.raises(TypeError):
        of.load_sintel_mask(0)
    with pytest.raises(ValueError):
        of.load_sintel_mask('no_such_file.png')
    with pytest.raises(ValueError):
        of.load_kitti('no_such_file')
    with pytest.raises(ValueError):
        of.load_kitti(os.path.join(FILES, 'wrong_kitti.png'))
    with pytest.raises(TypeError):
        of.Flow.from_kitti(os.path.join(FILES, 'sample0_kitti.png'), load_valid='yes')


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
def test_cuda_padding_matches_reference_and_oracle():
    import oflibpytorch_b200 as of
    from oracle import oracle as orc
    for name, want in KAT_PADDING.items():
        fl = of.Flow(T('pad_rot45_%s_v' % name).cuda(), name[0], T('pad_rot45_%s_m' % name).cuda())
        assert fl.get_padding() == [want]
        assert fl.get_padding(0) == want
    v, m = T('pad_rnd_v').cuda(), T('pad_rnd_m').cuda()
    for ref in 'st':
        assert of.Flow(v, ref, m).get_padding() == ING['pad_rnd_' + ref].tolist()
        assert of.Flow(v, ref).get_padding() == ING['pad_rnd_nomask_' + ref].tolist()
        assert of.get_flow_padding(v, ref) == ING['pad_rnd_nomask_' + ref].tolist()
        assert of.get_flow_padding(v[0], ref) == ING['pad_rnd_nomask_' + ref][0].tolist()
        assert of.Flow(v, ref, m).get_padding(2) == ING['pad_rnd_' + ref][2].tolist()
    # larger, odd-sized, batched: against the oracle
    v = smooth_flow(5, 131, 257, 9.0, 41)
    m = torch.rand(5, 131, 257, generator=gen(42)) > 0.5
    for ref in 'st':
        assert of.Flow(v.cuda(), ref, m.cuda()).get_padding() == orc.get_padding(orc.OFlow(v, ref, m))
    # test/test_flow_class.py:1645-1647: sub-threshold vectors never ask for padding
    z = torch.zeros(1, 2, 7, 7)
    z[0, 0] = torch.rand(7, 7, generator=gen(43)) * 1e-4
    assert of.Flow(z.cuda()).get_padding() == [[0, 0, 0, 0]]
    with pytest.raises(RuntimeError):
        of.Flow(v.cuda(), 't', torch.zeros(5, 131, 257, dtype=torch.bool).cuda()).get_padding()


@pytest.mark.gpu
def test_cuda_resize_matches_reference():
    import oflibpytorch_b200 as of
    v, m = T('rsz_v'), T('rsz_m')
    for i, s in enumerate(scales()):
        arg = s[0] if s[0] == s[1] else list(s)
        x = v.cuda().requires_grad_()
        r = of.resize_flow(x, arg)
        assert r.grad_fn is not None
        want = T('rsz_%d_out' % i)
        assert tuple(r.shape) == tuple(want.shape)
        assert torch.allclose(r.detach().cpu(), want, rtol=1e-5, atol=1e-5), 'scale %r' % (s,)
        r.backward(T('rsz_%d_gout' % i).cuda())
        assert torch.allclose(x.grad.cpu(), T('rsz_%d_gin' % i), rtol=1e-5, atol=1e-5), 'scale %r' % (s,)
        fl = of.Flow(v.cuda(), 't', m.cuda()).resize(arg)
        assert torch.equal(fl.mask.cpu(), T('rsz_%d_mask' % i)), 'scale %r' % (s,)
        assert torch.allclose(fl.vecs.cpu(), want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(of.resize_flow(T('rsz_big_v').cuda(), 1 / 3).cpu(), T('rsz_big_out'), rtol=1e-5, atol=1e-5)
    r3 = of.resize_flow(T('rsz_v')[0].cuda(), 1.5)
    assert r3.dim() == 3 and torch.allclose(r3.cpu(), T('rsz_3d_out'), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_cuda_resize_matches_oracle_on_large_fields():
    import oflibpytorch_b200 as of
    from oracle import oracle as orc
    v = smooth_flow(3, 203, 311, 7.0, 51)
    m = torch.rand(3, 203, 311, generator=gen(52)) > 0.4
    for s in (0.5, [0.37, 1.91], 2, [3.3, 0.25]):
        r = of.resize_flow(v.cuda(), s).cpu()
        assert torch.equal(r, orc.resize_flow(v, s)), 'the CUDA kernel follows the oracle operation by operation'
        assert torch.equal(of.Flow(v.cuda(), 's', m.cuda()).resize(s).mask.cpu(), orc.resize_mask(m, s))
        go = torch.rand(r.shape, generator=gen(53))
        x = v.cuda().requires_grad_()
        of.resize_flow(x, s).backward(go.cuda())
        assert torch.allclose(x.grad.cpu(), orc.resize_flow_bwd(go, s, 203, 311), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_cuda_loaders_match_reference():
    import oflibpytorch_b200 as of
    paths = [os.path.join(FILES, 'sample%d.flo' % k) for k in range(2)]
    batch = of.load_sintel_batch(paths)
    assert batch.is_cuda and tuple(batch.shape) == (2, 2, 13, 17)
    for k in range(2):
        assert torch.equal(batch[k].cpu(), T('flo%d' % k))
        assert torch.equal(of.load_sintel(paths[k]).cpu(), T('flo%d' % k))
    assert torch.equal(of.load_sintel_mask(os.path.join(FILES, 'sample_invalid.png')).cpu(), T('flo_mask'))
    fl = of.Flow.from_sintel(paths[0], os.path.join(FILES, 'sample_invalid.png'))
    assert fl.ref == 's' and torch.equal(fl.vecs.cpu(), T('from_sintel_v')) and torch.equal(fl.mask.cpu(), T('from_sintel_m'))
    kp = [os.path.join(FILES, 'sample%d_kitti.png' % k) for k in range(2)]
    flow, valid = of.load_kitti_batch(kp)
    for k in range(2):
        want = T('kitti%d' % k)
        assert torch.equal(flow[k].cpu(), want[:2]) and torch.equal(valid[k].cpu(), want[2] > 0)
        assert torch.equal(of.load_kitti(kp[k]).cpu(), want)
    fl = of.Flow.from_kitti(kp[0])
    assert fl.ref == 's' and torch.equal(fl.vecs.cpu(), T('from_kitti_v')) and torch.equal(fl.mask.cpu(), T('from_kitti_m'))
    assert bool(of.Flow.from_kitti(kp[0], load_valid=False).mask.all())


@pytest.mark.gpu
def test_cuda_flo_deinterleave_odd_sizes():
    from oflibpytorch_b200 import _ops
    for n, h, w in ((1, 1, 1), (2, 3, 5), (3, 16, 20), (1, 7, 13)):
        x = torch.randn(n, h, w, 2, generator=gen(60 + h))
        assert torch.equal(_ops.flo_to_planar(x.cuda()).cpu(), x.permute(0, 3, 1, 2).contiguous())
