"""Shared test helpers: golden-vector access, seeded input recipes, parity checkers."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hotpath_v1.npz')
_cache = {}


def golden():
    if 'g' not in _cache:
        _cache['g'] = np.load(GOLDEN)
    return _cache['g']


def G(case, key, device=None):
    """Golden array `case/key` as a torch tensor."""
    t = torch.from_numpy(np.array(golden()[case + '/' + key]))
    return t if device is None else t.to(device)


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def smooth_flow(n, h, w, sigma, seed, cell=16):
    """SURVEY.md 8(d) smooth-random flow: bicubic upsampling of a coarse Gaussian field (CPU, seeded)."""
    lo = torch.randn(n, 2, h // cell + 2, w // cell + 2, generator=gen(seed)) * sigma
    return torch.nn.functional.interpolate(lo, size=(h, w), mode='bicubic', align_corners=True).contiguous()


def holed_mask(n, h, w, seed):
    m = torch.rand(n, h, w, generator=gen(seed)) > 0.05
    m[:, h // 4:h // 2, w // 3:w // 2] = False
    m[:, :, ::11] = False
    return m


# Tolerance of BASELINE.json north_star: floats within 1e-4 relative / 1e-3 absolute; masks bit-exact.
RTOL, ATOL = 1e-4, 1e-3


def assert_close(a, b, what='', rtol=RTOL, atol=ATOL):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    assert a.shape == b.shape, '%s: shape %s vs %s' % (what, tuple(a.shape), tuple(b.shape))
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), '%s: %d of %d outside tolerance, max err %.3e' % (
        what, int(bad.sum()), bad.numel(), float(err.max()))


def assert_grad_close(a, b, what='', rel=1e-4):
    """Gradients: 1e-4 relative to the gradient's own scale (gradient conditioning),
    i.e. |a-b| <= rel * max(1, max|b|) + rel*|b|."""
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    assert a.shape == b.shape, '%s: shape %s vs %s' % (what, tuple(a.shape), tuple(b.shape))
    scale = max(1.0, float(b.abs().max()))
    err = (a - b).abs()
    bad = err > (rel * scale + rel * b.abs())
    assert not bool(bad.any()), '%s: %d of %d outside tolerance, max err %.3e (scale %.3e)' % (
        what, int(bad.sum()), bad.numel(), float(err.max()), scale)


def assert_mask_equal(a, b, what=''):
    a = a.detach().cpu()
    b = b.detach().cpu()
    assert a.dtype == torch.bool and b.dtype == torch.bool, what + ': masks must be bool'
    assert a.shape == b.shape, '%s: shape %s vs %s' % (what, tuple(a.shape), tuple(b.shape))
    n = int((a != b).sum())
    assert n == 0, '%s: %d of %d mask bits differ' % (what, n, a.numel())


def adversarial_case(name, hh=48, ww=1920):
    """Inputs of the 'adv_*' golden cases (tests/golden/make_golden.py, same recipe)."""
    tm = torch.ones(1, hh, ww, dtype=torch.bool)
    tm[:, 10:30, 700:1100] = False
    tm[:, :, ::97] = False
    im1 = torch.rand(1, 1, hh, ww, generator=gen(60))
    if name == 'int':
        fl = torch.tensor([7.0, -3.0]).view(1, 2, 1, 1).expand(1, 2, hh, ww).contiguous()
    elif name == 'half':
        fl = torch.round(smooth_flow(1, hh, ww, 6.0, 61) * 2) / 2
    else:
        fl = smooth_flow(1, hh, ww, 6.0, 62)
    return fl, tm, im1
