"""Edge cases of the CUDA path against the CPU oracle: smallest fields, wide payloads (planar fallback kernels),
channel-group splits, non-contiguous / non-fp32 inputs, batch broadcasting, flows far outside the image."""
import pytest
import torch

from helpers import assert_close, assert_grad_close, assert_mask_equal, gen, holed_mask, smooth_flow

pytestmark = pytest.mark.gpu


def of():
    import oflibpytorch_b200
    return oflibpytorch_b200


def orc():
    from oracle import oracle
    return oracle


def both(flow, ref, fmask, img, tmask, gout, consider_mask=True):
    o = orc()
    fo, io = flow.clone().requires_grad_(), img.clone().float().requires_grad_()
    oo, vo = o.OFlow(fo, ref, fmask).apply(io, tmask, return_valid_area=True, consider_mask=consider_mask)
    oo.backward(gout)
    fg, ig = flow.cuda().requires_grad_(), img.cuda().float().requires_grad_()
    og, vg = of().Flow(fg, ref, None if fmask is None else fmask.cuda()).apply(
        ig, None if tmask is None else tmask.cuda(), return_valid_area=True, consider_mask=consider_mask)
    og.backward(gout.cuda())
    assert_close(og, oo, 'out')
    assert_mask_equal(vg, vo, 'valid')
    assert_grad_close(fg.grad, fo.grad, 'gflow')
    assert_grad_close(ig.grad, io.grad, 'gimg')


@pytest.mark.parametrize('ref', ['t', 's'])
@pytest.mark.parametrize('shape', [(1, 1, 2, 2), (1, 2, 2, 7), (2, 1, 9, 2), (1, 3, 3, 3), (1, 1, 64, 65), (1, 2, 33, 128)])
def test_tiny_and_odd_fields(ref, shape):
    n, c, h, w = shape
    flow = torch.randn(n, 2, h, w, generator=gen(900 + h * w)) * 1.5
    both(flow, ref, holed_mask(n, h, w, 901), torch.rand(n, c, h, w, generator=gen(902)),
         holed_mask(n, h, w, 903), torch.rand(n, c, h, w, generator=gen(904)))


@pytest.mark.parametrize('ref', ['t', 's'])
@pytest.mark.parametrize('c', [4, 5, 7, 8, 9, 13])
def test_wide_payloads(ref, c):
    """C > 4 splits into channel groups (W) / takes KP = 8 or the planar splat kernels (S)."""
    n, h, w = 2, 40, 56
    flow = smooth_flow(n, h, w, 3.0, 910 + c)
    both(flow, ref, holed_mask(n, h, w, 911), torch.rand(n, c, h, w, generator=gen(912)),
         holed_mask(n, h, w, 913), torch.rand(n, c, h, w, generator=gen(914)))
    both(flow, ref, None, torch.rand(n, c, h, w, generator=gen(915)), None, torch.rand(n, c, h, w, generator=gen(916)))


@pytest.mark.parametrize('ref', ['t', 's'])
def test_flows_far_outside_the_image(ref):
    n, c, h, w = 2, 2, 24, 40
    flow = torch.randn(n, 2, h, w, generator=gen(920)) * 30
    flow[0, :, :5] = 1e6                    # nothing of these rows lands anywhere
    flow[1, :, 10:12] = -3e9
    both(flow, ref, holed_mask(n, h, w, 921), torch.rand(n, c, h, w, generator=gen(922)), holed_mask(n, h, w, 923),
         torch.rand(n, c, h, w, generator=gen(924)))


@pytest.mark.parametrize('ref', ['t', 's'])
def test_non_contiguous_and_other_dtypes(ref):
    m, o = of(), orc()
    n, c, h, w = 2, 3, 30, 44
    flow_hwc = smooth_flow(n, h, w, 3.0, 930).permute(0, 2, 3, 1).contiguous()        # (N,H,W,2) layout
    img64 = torch.rand(n, c, h, w, generator=gen(931), dtype=torch.float64)
    img_nc = torch.rand(n, h, w, c, generator=gen(932)).permute(0, 3, 1, 2)            # non-contiguous view
    fo = o.OFlow(flow_hwc.permute(0, 3, 1, 2).contiguous(), ref)
    fg = m.Flow(flow_hwc.cuda(), ref)
    assert fg.shape == (n, h, w)
    r = fg.apply(img64.cuda())
    assert r.dtype == torch.float64
    assert_close(r, fo.apply(img64), 'float64 target')
    assert_close(fg.apply(img_nc.cuda()), fo.apply(img_nc.contiguous()), 'non-contiguous target')
    half = torch.rand(n, c, h, w, generator=gen(933)).half()
    assert fg.apply(half.cuda()).dtype == torch.float16
    i16 = (torch.rand(n, c, h, w, generator=gen(934)) * 1000).to(torch.int16)
    assert torch.equal(m.apply_flow(fg.vecs, i16.cuda(), ref).cpu(), o.apply_flow(fo.vecs, i16, ref)) or \
        float((m.apply_flow(fg.vecs, i16.cuda(), ref).cpu() - o.apply_flow(fo.vecs, i16, ref)).abs().max()) <= 1


@pytest.mark.parametrize('ref', ['t', 's'])
def test_batch_broadcasting_with_gradients(ref):
    """one flow, many targets and many flows, one target: gradients of the broadcast operand are summed."""
    m, o = of(), orc()
    n, c, h, w = 3, 2, 26, 38
    flow1, flown = smooth_flow(1, h, w, 3.0, 940), smooth_flow(n, h, w, 3.0, 941)
    img1, imgn = torch.rand(1, c, h, w, generator=gen(942)), torch.rand(n, c, h, w, generator=gen(943))
    gout = torch.rand(n, c, h, w, generator=gen(944))
    for fl, im in ((flow1, imgn), (flown, img1)):
        fo, io = fl.clone().requires_grad_(), im.clone().requires_grad_()
        o.apply_flow(fo, io, ref).backward(gout)
        fg, ig = fl.cuda().requires_grad_(), im.cuda().requires_grad_()
        r = m.apply_flow(fg, ig, ref)
        r.backward(gout.cuda())
        assert r.shape == (n, c, h, w)
        assert_close(r, o.apply_flow(fl, im, ref), 'out')
        assert_grad_close(fg.grad, fo.grad, 'gflow')
        assert_grad_close(ig.grad, io.grad, 'gimg')


def test_grid_from_unstructured_density_gradient_and_validity_modes():
    m, o = of(), orc()
    n, c, h, w = 2, 2, 28, 36
    x = (torch.rand(n, h, w, generator=gen(950)) * (w + 4) - 2)
    y = (torch.rand(n, h, w, generator=gen(951)) * (h + 4) - 2)
    d = torch.randn(n, c, h, w, generator=gen(952))
    gd = torch.rand(n, h, w, generator=gen(953))
    xo, yo, do = (t.clone().requires_grad_() for t in (x, y, d))
    oo, deno = o.grid_from_unstructured_data(xo, yo, do)
    (deno * gd).sum().backward()                                     # gradient ONLY through the density output
    xg, yg, dg = (t.cuda().requires_grad_() for t in (x, y, d))
    og, deng = m.grid_from_unstructured_data(xg, yg, dg)
    (deng * gd.cuda()).sum().backward()
    assert_close(og, oo, 'data')
    assert_close(deng, deno, 'density')
    assert_grad_close(xg.grad, xo.grad, 'gx')
    assert_grad_close(yg.grad, yo.grad, 'gy')
    assert float(dg.grad.abs().max()) == 0.0 and float(do.grad.abs().max()) == 0.0


def test_apply_s_flow_returns_warped_mask_and_density_path():
    m, o = of(), orc()
    n, c, h, w = 2, 3, 31, 47
    flow = smooth_flow(n, h, w, 4.0, 960)
    flow[:, :, 8:14, 10:30] = 0.0
    img, mask = torch.rand(n, c, h, w, generator=gen(961)), holed_mask(n, h, w, 962)
    for occ in (True, False):
        with torch.no_grad():
            og, wg = m.apply_s_flow(flow.cuda(), img.cuda(), mask.cuda(), occ)
        oo, wo = o.apply_s_flow(flow, img, mask, occ)
        assert_close(og, oo, 'out')
        assert_mask_equal(wg, wo, 'warped mask')


def test_windowed_warp_equals_rows_of_the_full_warp():
    """ofb_warp_t_win_fwd/bwd: warping a row band in the global frame with a halo-extended source band gives exactly the
    corresponding rows of the full-field warp (this is what row-band spatial sharding relies on)."""
    from oflibpytorch_b200 import _ops
    n, c, h, w = 2, 3, 96, 80
    flow = smooth_flow(n, h, w, 4.0, 970)
    img = torch.rand(n, c, h, w, generator=gen(971))
    fmask, tmask = holed_mask(n, h, w, 972), holed_mask(n, h, w, 973)
    gout = torch.rand(n, c, h, w, generator=gen(974))
    fg, ig = flow.cuda().requires_grad_(), img.cuda().requires_grad_()
    full, vfull = of().Flow(fg, 't', fmask.cuda()).apply(ig, tmask.cuda(), return_valid_area=True)
    full.backward(gout.cuda())
    k = int(flow[:, 1].abs().max().ceil()) + 2
    for a, b in ((0, 40), (40, 71), (71, 96)):
        lo, hi = max(a - k, 0), min(b + k, h)
        fb = flow[:, :, a:b].contiguous().cuda().requires_grad_()
        sb = img[:, :, lo:hi].contiguous().cuda().requires_grad_()
        out, valid = _ops.warp_t_win(fb, sb, [a, 0, h, w, lo, 0, hi - lo, w], tmask[:, lo:hi].contiguous().cuda(),
                                     fmask[:, a:b].contiguous().cuda(), want_valid=True)
        assert torch.equal(out, full[:, :, a:b]) and torch.equal(valid, vfull[:, a:b])
        out.backward(gout[:, :, a:b].contiguous().cuda())
        assert_grad_close(fb.grad, fg.grad[:, :, a:b], 'gflow band')
        # the band's contribution to the image gradient is confined to its halo-extended rows
        assert sb.grad.shape == sb.shape


def test_fused_padding_is_bit_identical_to_the_padded_path():
    """Flow.apply(padding=..., cut=True) for 't' flows runs the windowed kernel (no F.pad, no slice)."""
    o = orc()
    n, c, h, w = 2, 3, 40, 52
    pad = [5, 3, 7, 2]
    flow = smooth_flow(n, h, w, 3.0, 980)
    fmask = holed_mask(n, h, w, 981)
    big = torch.rand(n, c, h + pad[0] + pad[1], w + pad[2] + pad[3], generator=gen(982))
    bigm = torch.rand(n, h + pad[0] + pad[1], w + pad[2] + pad[3], generator=gen(983)) > 0.1
    oo, vo = o.OFlow(flow, 't', fmask).apply(big, bigm, return_valid_area=True, padding=pad, cut=True)
    og, vg = of().Flow(flow.cuda(), 't', fmask.cuda()).apply(big.cuda(), bigm.cuda(), return_valid_area=True,
                                                             padding=pad, cut=True)
    assert torch.equal(og.cpu(), oo) and torch.equal(vg.cpu(), vo)
    z = of().Flow(torch.zeros(n, 2, h, w, device='cuda'), 't', fmask.cuda())
    og, vg = z.apply(big.cuda(), bigm.cuda(), return_valid_area=True, padding=pad, cut=True)
    assert torch.equal(og.cpu(), big[..., pad[0]:pad[0] + h, pad[2]:pad[2] + w])
    assert torch.equal(vg.cpu(), bigm[:, pad[0]:pad[0] + h, pad[2]:pad[2] + w] & fmask)


def test_fused_padding_without_cut_is_bit_identical_to_the_padded_path():
    """Flow.apply(padding=..., cut=False) for 't' flows (reference flow_class.py:905-934) runs the inset-flow kernel: the
    whole target frame warped by a flow that is zero outside its window, valid area false out there; no F.pad.  Checked
    against the oracle (which pads like the reference), gradients included, and for a Flow target and a zero flow."""
    o = orc()
    n, c, h, w = 2, 3, 40, 52
    pad = [5, 3, 7, 2]
    flow = smooth_flow(n, h, w, 3.0, 990)
    fmask = holed_mask(n, h, w, 991)
    big = torch.rand(n, c, h + pad[0] + pad[1], w + pad[2] + pad[3], generator=gen(992))
    bigm = torch.rand(n, h + pad[0] + pad[1], w + pad[2] + pad[3], generator=gen(993)) > 0.1
    gout = torch.randn(big.shape, generator=gen(994))
    fo, bo = flow.clone().requires_grad_(), big.clone().requires_grad_()
    oo, vo = o.OFlow(fo, 't', fmask).apply(bo, bigm, return_valid_area=True, padding=pad, cut=False)
    oo.backward(gout)
    fg, bg = flow.cuda().requires_grad_(), big.cuda().requires_grad_()
    from oflibpytorch_b200 import _native
    launches = _native.launch_count()
    og, vg = of().Flow(fg, 't', fmask.cuda()).apply(bg, bigm.cuda(), return_valid_area=True, padding=pad, cut=False)
    assert og.shape == big.shape and vg.shape == bigm.shape
    assert torch.equal(og.detach().cpu(), oo.detach()) and torch.equal(vg.cpu(), vo)
    og.backward(gout.cuda())
    assert_grad_close(fg.grad, fo.grad, 'gflow (window-sized)')
    assert_grad_close(bg.grad, bo.grad, 'gradient of the whole target frame')
    assert not bool(vg[:, :pad[0]].any()) and not bool(vg[:, :, :pad[2]].any())
    # single batch item flow against a batch of targets, no valid area, 3-d target
    o1 = o.OFlow(flow[:1], 't', fmask[:1]).apply(big, padding=pad, cut=False)
    g1 = of().Flow(flow[:1].cuda(), 't', fmask[:1].cuda()).apply(big.cuda(), padding=pad, cut=False)
    assert torch.equal(g1.cpu(), o1)
    # Flow target
    fb = smooth_flow(n, h + pad[0] + pad[1], w + pad[2] + pad[3], 2.0, 995)
    ro = o.OFlow(flow, 't', fmask).apply(o.OFlow(fb, 't', bigm), padding=pad, cut=False)
    rg = of().Flow(flow.cuda(), 't', fmask.cuda()).apply(of().Flow(fb.cuda(), 't', bigm.cuda()), padding=pad, cut=False)
    assert torch.equal(rg.vecs.cpu(), ro.vecs) and torch.equal(rg.mask.cpu(), ro.mask)
    # zero flow: the target comes back, valid only inside the window
    z = of().Flow(torch.zeros(n, 2, h, w, device='cuda'), 't', fmask.cuda())
    og, vg = z.apply(big.cuda(), bigm.cuda(), return_valid_area=True, padding=pad, cut=False)
    zo, zv = o.OFlow(torch.zeros(n, 2, h, w), 't', fmask).apply(big, bigm, return_valid_area=True, padding=pad, cut=False)
    assert torch.equal(og.cpu(), zo) and torch.equal(vg.cpu(), zv)


def test_padded_s_apply_with_cut_equals_the_reference_padded_path():
    """Flow.apply(padding=..., cut=True) for 's' flows: the reference pads flow (replicate) and mask (False), splats the
    whole padded frame and cuts.  With consider_mask the padded sources are ignored, so this package splats only the
    window -- but in the padded FRAME (ofb_splat_s_win_fwd): computing the end points as x + f instead of
    x + left + f rounds differently, and where the density is below the 1e-3 clamp that difference is amplified past
    the tolerance (measured: 1.6e-3 at one pixel of 17 202 here).  Checked against the oracle, which follows the
    reference, with flows large enough to push pixels out of the window, gradients, a Flow target, and
    consider_mask=False (which keeps the padded path)."""
    o = orc()
    n, c, h, w = 2, 3, 47, 61
    pad = [6, 4, 9, 3]
    flow = smooth_flow(n, h, w, 7.0, 990)
    flow[:, :, 10:20, 15:30] = 0.0                                   # occlusion rule inside the window
    fmask = holed_mask(n, h, w, 991)
    big = torch.rand(n, c, h + pad[0] + pad[1], w + pad[2] + pad[3], generator=gen(992))
    bigm = torch.rand(n, h + pad[0] + pad[1], w + pad[2] + pad[3], generator=gen(993)) > 0.1
    for consider in (True, False):
        oo, vo = o.OFlow(flow, 's', fmask).apply(big, bigm, return_valid_area=True, padding=pad, cut=True,
                                                 consider_mask=consider)
        og, vg = of().Flow(flow.cuda(), 's', fmask.cuda()).apply(big.cuda(), bigm.cuda(), return_valid_area=True,
                                                                 padding=pad, cut=True, consider_mask=consider)
        assert_close(og, oo, 'padded s apply, consider_mask=%s' % consider)
        assert_mask_equal(vg, vo, 'padded s valid, consider_mask=%s' % consider)
    # gradients through the windowed splat (flow and the WHOLE padded image: zero outside the window)
    gout = torch.rand(n, c, h, w, generator=gen(995))
    fo, bo = flow.clone().requires_grad_(), big.clone().requires_grad_()
    o.OFlow(fo, 's', fmask).apply(bo, bigm, padding=pad, cut=True).backward(gout)
    fg, bg = flow.cuda().requires_grad_(), big.cuda().requires_grad_()
    of().Flow(fg, 's', fmask.cuda()).apply(bg, bigm.cuda(), padding=pad, cut=True).backward(gout.cuda())
    assert_grad_close(fg.grad, fo.grad, 'padded s apply, gflow')
    assert_grad_close(bg.grad, bo.grad, 'padded s apply, gimage')
    fb = smooth_flow(n, h + pad[0] + pad[1], w + pad[2] + pad[3], 2.0, 994)
    ro = o.OFlow(flow, 's', fmask).apply(o.OFlow(fb, 's', bigm), padding=pad, cut=True)
    rg = of().Flow(flow.cuda(), 's', fmask.cuda()).apply(of().Flow(fb.cuda(), 's', bigm.cuda()), padding=pad, cut=True)
    assert_close(rg.vecs, ro.vecs, 'padded s apply to a flow')
    assert_mask_equal(rg.mask, ro.mask, 'padded s apply to a flow, mask')


def test_row_band_splat_equals_rows_of_the_full_splat():
    """ofb_splat_s_band_accumulate / _normalise: splatting row bands into halo-extended accumulators, adding the halos
    and normalising per band reproduces the full-field 's' warp (single GPU emulation of the multi-GPU scheme)."""
    from oflibpytorch_b200 import _ops
    n, c, h, w = 2, 3, 90, 72
    flow = smooth_flow(n, h, w, 4.0, 990)
    flow[:, :, 30:40, 10:30] = 0.0
    img = torch.rand(n, c, h, w, generator=gen(991))
    fmask, tmask = holed_mask(n, h, w, 992), holed_mask(n, h, w, 993)
    full, vfull = of().Flow(flow.cuda(), 's', fmask.cuda()).apply(img.cuda(), tmask.cuda(), return_valid_area=True)
    k = int(flow[:, 1].abs().max().ceil()) + 2
    bands = [(0, 35), (35, 61), (61, 90)]
    dm = (tmask & fmask).cuda()
    accs = []
    for a, b in bands:
        lo, hi = max(a - k, 0), min(b + k, h)
        accs.append((lo, hi, _ops.splat_band_accumulate(flow[:, :, a:b].contiguous().cuda(), img[:, :, a:b].contiguous().cuda(),
                                                         dm[:, a:b].contiguous(), fmask[:, a:b].contiguous().cuda(), 1.0,
                                                         True, (a, h, lo, hi - lo))))
    total = torch.zeros((n, h, w, accs[0][2].shape[-1]), device='cuda')
    for lo, hi, acc in accs:                                   # fixed order: band 0, 1, 2
        total[:, lo:hi] += acc
    for a, b in bands:
        out, valid = _ops.splat_band_normalise(total[:, a:b].contiguous(), flow[:, :, a:b].contiguous().cuda(),
                                               img[:, :, a:b].contiguous().cuda(), dm[:, a:b].contiguous(),
                                               fmask[:, a:b].contiguous().cuda(), True, _ops.VALID_GT, (a, h, a, b - a))
        assert_close(out, full[:, :, a:b], 'band out')
        assert_mask_equal(valid, vfull[:, a:b], 'band valid')


def test_row_band_splat_adjoint_equals_rows_of_the_full_gradient():
    """Row-band splat adjoint (ofb_splat_s_band_bwd_prepass / _gather): per-band prepass, halo rows of G copied in from
    the neighbouring bands, gather for the band's own sources -- reproduces the rows of the full-field gradients
    (single GPU emulation of the multi-GPU scheme).  Also the public wrapper with world_size = 1 under autograd."""
    from oflibpytorch_b200 import _ops, sharding
    n, c, h, w = 2, 3, 90, 72
    flow = smooth_flow(n, h, w, 4.0, 996)
    flow[:, :, 30:40, 10:30] = 0.0
    img = torch.rand(n, c, h, w, generator=gen(997))
    gout = torch.rand(n, c, h, w, generator=gen(998))
    fmask, tmask = holed_mask(n, h, w, 999), holed_mask(n, h, w, 1000)
    fg, ig = flow.cuda().requires_grad_(), img.cuda().requires_grad_()
    full, vfull = of().Flow(fg, 's', fmask.cuda()).apply(ig, tmask.cuda(), return_valid_area=True)
    full.backward(gout.cuda())
    # forward pieces of the full field that the bands would hold (out, density, unocc) via the band API itself
    k = int(flow[:, 1].abs().max().ceil()) + 2
    bands = [(0, 35), (35, 61), (61, 90)]
    dm = (tmask & fmask).cuda()
    fl, im, fm, go = flow.cuda(), img.cuda(), fmask.cuda(), gout.cuda()
    accs = []
    for a, b in bands:
        lo, hi = max(a - k, 0), min(b + k, h)
        accs.append((lo, hi, _ops.splat_band_accumulate(fl[:, :, a:b].contiguous(), im[:, :, a:b].contiguous(),
                                                         dm[:, a:b].contiguous(), fm[:, a:b].contiguous(), 1.0, True,
                                                         (a, h, lo, hi - lo))))
    total = torch.zeros((n, h, w, accs[0][2].shape[-1]), device='cuda')
    for lo, hi, acc in accs:
        total[:, lo:hi] += acc
    gs, aux = [], []
    for a, b in bands:
        out, valid, den, unocc = _ops.splat_band_normalise(total[:, a:b].contiguous(), fl[:, :, a:b].contiguous(),
                                                           im[:, :, a:b].contiguous(), dm[:, a:b].contiguous(),
                                                           fm[:, a:b].contiguous(), True, _ops.VALID_GT, (a, h, a, b - a),
                                                           want_aux=True)
        aux.append(unocc)
        gs.append(_ops.splat_band_bwd_prepass(out, den, unocc, go[:, :, a:b].contiguous()))
    g_full = torch.cat(gs, dim=1)                                           # what the halo exchange assembles
    for (a, b), unocc in zip(bands, aux):
        lo, hi = max(a - k, 0), min(b + k, h)
        gflow, gdata = _ops.splat_band_bwd_gather(fl[:, :, a:b].contiguous(), im[:, :, a:b].contiguous(),
                                                  fm[:, a:b].contiguous(), True, g_full[:, lo:hi].contiguous(), unocc,
                                                  go[:, :, a:b].contiguous(), (a, h, lo, hi - lo), True, True)
        assert_grad_close(gflow, fg.grad[:, :, a:b], 'band gflow')
        assert_grad_close(gdata, ig.grad[:, :, a:b], 'band gdata')
    # the public wrapper as a single band: autograd through accumulate / normalise / prepass / gather
    f1, i1 = flow.cuda().requires_grad_(), img.cuda().requires_grad_()
    o1, v1 = sharding.apply_s_rowbands(f1, i1, h_total=h, rank=0, world_size=1, flow_mask_band=fm,
                                       payload_mask_band=tmask.cuda())
    o1.backward(go)
    assert_close(o1, full, 'single-band out')
    assert_mask_equal(v1, vfull, 'single-band valid')
    assert_grad_close(f1.grad, fg.grad, 'single-band gflow')
    assert_grad_close(i1.grad, ig.grad, 'single-band gdata')


def test_chained_combine_apply_backward_matches_oracle():
    """BASELINE.json config 5 at test scale: acc = acc.combine_with(f_i, 3); img_i = acc.apply(img, valid area);
    loss = sum_i <g_i, img_i>; ONE backward through the whole chain -> gradients of every frame's flow and of the image."""
    o, m = orc(), of()
    frames, c, h, w = 5, 3, 48, 64
    flows = [smooth_flow(1, h, w, 2.0, 1100 + i) for i in range(frames)]
    img = torch.rand(1, c, h, w, generator=gen(1200))
    gs = [torch.rand(1, c, h, w, generator=gen(1300 + i)) for i in range(frames)]

    def run(flow_cls, dev, apply_kw):
        fl = [f.detach().clone().to(dev).requires_grad_() for f in flows]
        im = img.detach().clone().to(dev).requires_grad_()
        acc, loss, valids = None, 0.0, []
        for i in range(frames):
            cur = flow_cls(fl[i], 't')
            acc = cur if acc is None else acc.combine_with(cur, 3)
            out, valid = acc.apply(im, return_valid_area=True)
            loss = loss + (out * gs[i].to(dev)).sum()
            valids.append(valid)
        loss.backward()
        return acc, valids, [f.grad for f in fl], im.grad

    acc_o, val_o, gf_o, gi_o = run(o.OFlow, 'cpu', {})
    acc_g, val_g, gf_g, gi_g = run(m.Flow, 'cuda', {})
    assert_close(acc_g.vecs, acc_o.vecs, 'chained flow')
    assert_mask_equal(acc_g.mask, acc_o.mask, 'chained mask')
    for i in range(frames):
        assert_mask_equal(val_g[i], val_o[i], 'valid %d' % i)
        assert_grad_close(gf_g[i], gf_o[i], 'gflow %d' % i)
    assert_grad_close(gi_g, gi_o, 'gimg')


@pytest.mark.parametrize('ref', ['t', 's'])
def test_run_to_run_determinism(ref):
    """Masks and every W output are bit-identical from run to run.  The splat and both adjoints accumulate with
    floating-point atomics whose order is not fixed: their floats may differ in the last bits (bounded here at 1e-5
    of the tensor's scale), their masks never do -- validity is decided by order-independent tests."""
    n, h, w = 2, 211, 333
    flow = smooth_flow(n, h, w, 9.0, 71).cuda()
    img = torch.rand(n, 3, h, w, generator=gen(72)).cuda()
    fmask, tmask = holed_mask(n, h, w, 73).cuda(), holed_mask(n, h, w, 74).cuda()
    gout = torch.rand(n, 3, h, w, generator=gen(75)).cuda()

    def run():
        f, i = flow.clone().requires_grad_(), img.clone().requires_grad_()
        fl = of().Flow(f, ref, fmask)
        out, valid = fl.apply(i, tmask, return_valid_area=True)
        out.backward(gout)
        sw = fl.switch_ref()
        return out.detach(), valid, f.grad, i.grad, sw.vecs.detach(), sw.mask, fl.valid_source(), fl.valid_target()

    first = run()
    for _ in range(3):
        again = run()
        for k, (a, b) in enumerate(zip(first, again)):
            if a.dtype == torch.bool:
                assert torch.equal(a, b), 'mask output %d changed between runs' % k
            elif ref == 't' and k == 0:
                assert torch.equal(a, b), 'the W forward is a pure gather and must be bit-identical'
            else:
                scale = float(a.abs().max()) + 1e-12
                assert float((a - b).abs().max()) <= 1e-5 * scale, 'output %d drifted between runs' % k


def test_deterministic_splat_mode_is_bit_reproducible_and_within_tolerance():
    """of.set_deterministic(True): the 's' forward accumulates in 64-bit fixed point -- float outputs bit-identical from
    run to run (north_star: 'a deterministic global merge'), still within tolerance of the oracle, masks unchanged."""
    from oracle import oracle as o
    m = of()
    n, h, w = 2, 157, 243
    flow = smooth_flow(n, h, w, 9.0, 81)
    img = torch.rand(n, 3, h, w, generator=gen(82)) * 7.0 - 2.0
    fmask, tmask = holed_mask(n, h, w, 83), holed_mask(n, h, w, 84)
    ref_out, ref_valid = o.OFlow(flow, 's', fmask).apply(img, tmask, return_valid_area=True)
    ref_sw = o.OFlow(flow, 's', fmask).switch_ref()
    fast = m.Flow(flow.cuda(), 's', fmask.cuda()).apply(img.cuda(), tmask.cuda(), return_valid_area=True)
    m.set_deterministic(True)
    try:
        runs = []
        for _ in range(4):
            fl = m.Flow(flow.cuda(), 's', fmask.cuda())
            out, valid = fl.apply(img.cuda(), tmask.cuda(), return_valid_area=True)
            sw = fl.switch_ref()
            runs.append((out, valid, sw.vecs, sw.mask))
    finally:
        m.set_deterministic(False)
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b), 'deterministic mode changed between runs'
    assert_close(runs[0][0], ref_out, 'deterministic apply s')
    assert_mask_equal(runs[0][1], ref_valid, 'deterministic apply s valid')
    assert_close(runs[0][2], ref_sw.vecs, 'deterministic switch_ref')
    assert_mask_equal(runs[0][3], ref_sw.mask, 'deterministic switch_ref mask')
    assert_mask_equal(runs[0][1], fast[1], 'deterministic vs default valid')
