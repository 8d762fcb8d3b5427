"""The kernel variants of W agree with each other and with the oracle: lean forward vs direct forward (bit-identical),
tile-owned fixed-point adjoint vs register-merge adjoint vs the oracle (reference: grid_sampler_2d_backward behind
utils.py:555), the adjoint's fall-backs (footprint larger than the window, non-finite upstream gradient, fields narrower
than a tile) and the saved-derivative mode (ofb_warp_t_fwd_d / ofb_warp_t_bwd_d)."""
import pytest
import torch

from helpers import assert_grad_close, gen, holed_mask, smooth_flow

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def ops():
    from oflibpytorch_b200 import _ops
    return _ops


def lib():
    from oflibpytorch_b200 import _native
    return _native.lib()


@pytest.fixture(autouse=True)
def _reset_knobs():
    yield
    lib().ofb_set_option(1, -1)
    lib().ofb_set_option(5, 0)
    lib().ofb_set_option(6, 0)
    ops().set_save_derivatives(False)


def case(n, c, h, w, seed, sigma=6.0):
    flow = smooth_flow(n, h, w, sigma, seed).to(DEV)
    img = torch.rand(n, c, h, w, generator=gen(seed + 1)).to(DEV)
    gout = torch.randn(n, c, h, w, generator=gen(seed + 2)).to(DEV)
    return flow, img, gout


@pytest.mark.parametrize('c,h,w', [(3, 70, 101), (1, 33, 64), (4, 64, 130), (2, 17, 31)])
def test_lean_forward_is_bit_identical_to_direct(c, h, w):
    flow, img, _ = case(2, c, h, w, 11)
    flow[0, :, :9, :] = 250.0                       # samples far outside
    flow[1, 0, :, : w // 2] = -3.5                  # border warps on the left edge
    m1, m2 = holed_mask(2, h, w, 5).to(DEV), holed_mask(2, h, w, 6).to(DEV)
    res = []
    for impl in (1, 0):
        lib().ofb_set_option(5, impl)
        out, valid, *_ = ops()._warp_t_fwd(flow, 1.0, img, m1, m2, 0.0, 0.999, True)
        res.append((out, valid))
    assert torch.equal(res[0][0].view(torch.int32), res[1][0].view(torch.int32))
    assert torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize('c,h,w', [(3, 70, 101), (1, 33, 64), (5, 40, 72), (2, 17, 31)])
def test_tile_adjoint_matches_register_merge_and_oracle(c, h, w):
    from oracle import oracle as orc
    flow, img, gout = case(2, c, h, w, 21)
    flow[0, :, :9, :] = 250.0
    flow[1, 0, :, : w // 2] = -3.5
    res = {}
    for impl in (0, 7):
        lib().ofb_set_option(1, impl)
        res[impl] = ops()._warp_t_bwd(flow, 1.0, img, gout, 0.0, True, True)
    gf_o, gs_o = orc.raw_warp_t_bwd(flow.cpu(), img.cpu(), gout.cpu())
    for impl in (0, 7):
        assert_grad_close(res[impl][0], gf_o, 'gflow impl %d' % impl)
        assert_grad_close(res[impl][1], gs_o, 'gsrc impl %d' % impl)
    assert torch.equal(res[0][0], res[7][0]), 'the flow gradient is the same arithmetic in both kernels'
    # fixed point: one contribution is rounded to 2^-19 of the tile's largest |gout|
    assert float((res[0][1] - res[7][1]).abs().max()) <= 16 * 2.0 ** -19 * float(gout.abs().max())


def test_tile_adjoint_fallbacks():
    from oracle import oracle as orc
    n, c, h, w = 2, 3, 96, 160
    flow, img, gout = case(n, c, h, w, 31)
    # violent compression: the footprint of a 32 x 32 tile exceeds the 64 x 64 window -> float REDs per tap
    flow[0, 0] += torch.arange(w, device=DEV, dtype=torch.float32)[None, :] * 2.5
    flow[0, 1] += torch.arange(h, device=DEV, dtype=torch.float32)[:, None] * 2.5
    lib().ofb_set_option(1, 7)
    gf, gs = ops()._warp_t_bwd(flow, 1.0, img, gout, 0.0, True, True)
    gf_o, gs_o = orc.raw_warp_t_bwd(flow.cpu(), img.cpu(), gout.cpu())
    assert_grad_close(gf, gf_o, 'gflow')
    assert_grad_close(gs, gs_o, 'gsrc')
    # a non-finite upstream gradient propagates like in the float path (the tile does not use fixed point)
    gout2 = gout.clone()
    gout2[1, 0, 40, 50] = float('inf')
    gf, gs = ops()._warp_t_bwd(flow, 1.0, img, gout2, 0.0, True, True)
    lib().ofb_set_option(1, 0)
    gf0, gs0 = ops()._warp_t_bwd(flow, 1.0, img, gout2, 0.0, True, True)
    assert torch.equal(torch.isfinite(gs), torch.isfinite(gs0))
    fin = torch.isfinite(gs0)
    assert_grad_close(gs[fin], gs0[fin], 'gsrc next to the non-finite entry')
    # zero upstream gradient: nothing is accumulated, the flow gradient is zero
    lib().ofb_set_option(1, 7)
    gf, gs = ops()._warp_t_bwd(flow, 1.0, img, torch.zeros_like(gout), 0.0, True, True)
    assert float(gs.abs().max()) == 0.0 and float(gf.abs().max()) == 0.0


def test_tile_adjoint_broadcast_payload_and_gsrc_only():
    from oracle import oracle as orc
    n, c, h, w = 3, 2, 50, 77
    flow, img, gout = case(n, c, h, w, 41)
    img1 = img[:1].contiguous()
    lib().ofb_set_option(1, 7)
    _, gs = ops()._warp_t_bwd(flow, 1.0, img1, gout, 0.0, False, True)
    assert gs.shape == img1.shape
    gs_o = sum(orc.raw_warp_t_bwd(flow[k:k + 1].cpu(), img1.cpu(), gout[k:k + 1].cpu())[1] for k in range(n))
    assert_grad_close(gs, gs_o, 'gsrc of a broadcast payload')


def test_saved_derivatives_give_the_same_gradients():
    import oflibpytorch_b200 as of
    n, c, h, w = 2, 3, 70, 101
    flow, img, gout = case(n, c, h, w, 51)
    flow[0, :, :9, :] = 250.0
    mask = holed_mask(n, h, w, 7).to(DEV)
    res = []
    for save in (False, True):
        ops().set_save_derivatives(save)
        f, i = flow.clone().requires_grad_(), img.clone().requires_grad_()
        out, valid = of.Flow(f, 't', mask).apply(i, mask, return_valid_area=True)
        out.backward(gout)
        res.append((out.detach(), valid, f.grad, i.grad))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2]), 'flow gradient: same arithmetic, split at the combined taps'
    assert_grad_close(res[1][3], res[0][3], 'image gradient')
    # combine mode 3 (C = 2, add_flow): the chain keeps its gradients
    for save in (False, True):
        ops().set_save_derivatives(save)
        a, b = flow.clone().requires_grad_(), (-flow.flip(0)).clone().requires_grad_()
        r = of.Flow(a, 't').combine_with(of.Flow(b, 't'), 3)
        r.vecs.backward(gout[:, :2].contiguous())
        res.append((a.grad, b.grad))
    assert torch.equal(res[2][0], res[3][0])
    assert_grad_close(res[3][1], res[2][1], 'payload flow gradient')


def test_broadcast_upstream_gradient_is_not_materialised():
    """out.sum().backward() hands the adjoint a stride-0 gradient; the default kernel reads the single value."""
    import oflibpytorch_b200 as of
    n, c, h, w = 2, 3, 70, 101
    flow, img, _ = case(n, c, h, w, 61)
    res = []
    for impl in (0, 7):       # 0: register-merge kernel (dense gout, the bcast request answers OFB_EINVAL), 7: tile kernel
        lib().ofb_set_option(1, impl)
        f, i = flow.clone().requires_grad_(), img.clone().requires_grad_()
        out, _v = of.Flow(f, 't').apply(i, return_valid_area=True)
        (out.sum() * 0.75).backward()
        res.append((f.grad, i.grad))
    assert torch.equal(res[0][0], res[1][0])
    assert_grad_close(res[1][1], res[0][1], 'image gradient')
    g1 = torch.full((1,), 0.75, device=DEV).expand(n, c, h, w)
    assert all(s == 0 for s in g1.stride())
    gf, gs = ops()._warp_t_bwd(flow, 1.0, img, g1, 0.0, True, True)
    gf2, gs2 = ops()._warp_t_bwd(flow, 1.0, img, g1.contiguous(), 0.0, True, True)
    assert torch.equal(gf, gf2)
    assert_grad_close(gs, gs2, 'image gradient, broadcast vs dense gout')


@pytest.mark.parametrize('mode', ['1', '2'])
def test_tma_staged_forward_variants_keep_parity(mode):
    """OFB_TMA=1 (TMA-staged tiles) and OFB_TMA=2 (persistent, double-buffered tile loop) are opt-in forward kernels
    (slower than the default); the golden / oracle parity tests of the warp must pass with either.  The
    switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OFB_TMA=mode)
    out = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_parity.py'), '-q', '-x',
                          '-k', 'warp_t or apply_flow or flow_apply_t or combine'],
                         env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-500:]


@pytest.mark.parametrize('c,gflow', [(3, True), (1, True), (4, False), (2, True)])
def test_tile_adjoint_zero_ahead_mode_matches_the_memset_path(c, gflow):
    """ofb_set_option(6, 1): the tile adjoint zero-fills the payload gradient itself, tile rows ahead of the REDs
    (warp_t_tile.cu).  Same results as the default path and the oracle, including tiles whose footprint lies far below
    them (deferred pass), far above, outside the field, and a field whose last tile row is partial."""
    from oracle import oracle as orc
    n, h, w = 3, 300, 200                              # 10 tile rows (the last one partial), 7 tile columns
    flow, img, gout = case(n, c, h, w, 41)
    flow[0, 1, 20:50, 30:150] = -220.0                 # samples 220 rows further down: deferred tiles
    flow[1, 1, 250:290, 10:100] = 240.0                # far above
    flow[2, 0, 100:120, :] = 900.0                     # outside
    res = []
    for za in (0, 1, 1):                               # twice: the counters are re-initialised per call
        lib().ofb_set_option(6, za)
        res.append(ops()._warp_t_bwd(flow, 1.0, img, gout, 0.0, gflow, True))
    gf_o, gs_o = orc.raw_warp_t_bwd(flow.cpu(), img.cpu(), gout.cpu())
    for gf, gs in res[1:]:
        assert_grad_close(gs, gs_o, 'gsrc (zero-ahead) vs oracle')
        assert_grad_close(gs, res[0][1], 'gsrc (zero-ahead) vs memset path')
        if gflow:
            assert torch.equal(gf, res[0][0])
            assert_grad_close(gf, gf_o, 'gflow')
