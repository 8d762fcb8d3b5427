"""Generates tests/golden/kats_v1.npz: the INPUTS of the reference's own known-answer tests, built by the
reference's constructors (Flow.from_transforms), plus the reference's outputs.  The expected tables themselves are
restated as literals in tests/kat_tables.py, citing the reference test lines.  Build container only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/reference/src')
import oflibpytorch as of  # noqa: E402
from oflibpytorch.utils import get_flow_endpoints, grid_from_unstructured_data, normalise_coords  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
of.set_pure_pytorch()
out = {}

# test/test_flow_class.py:1389-1400, 1505-1516: 45 degree rotation about the origin on a 7x7 field
rot = [['rotation', 0, 0, 45]]
out['rot45_s'] = of.Flow.from_transforms(rot, (7, 7), 's').vecs.numpy()
out['rot45_t'] = of.Flow.from_transforms(rot, (7, 7), 't').vecs.numpy()
# test/test_utils.py:1104-1120: rotation about (50, 75) on 100x150, count of density > 0.95 must be 24664 for a batch of 2
fl = of.Flow.from_transforms([['rotation', 50, 75, -20]], (100, 150), 's')
out['rot20_s'] = fl.vecs.numpy()
out['rot20_rev_s'] = of.Flow.from_transforms([['rotation', 50, 75, 20]], (100, 150), 's').vecs.numpy()
v2 = torch.cat((fl.vecs, fl.vecs))
x, y = get_flow_endpoints(v2, 's')
data, dens = grid_from_unstructured_data(x, y, v2)
out['rot20_count'] = np.array(int((dens > 0.95).sum()))
assert int(out['rot20_count']) == 24664
# test/test_utils.py:456-475 normalise_coords KAT
coords = torch.tensor([[0, 0], [10, 20], [100, 200], [639, 479]], dtype=torch.float)   # (x, y)
out['norm_in'] = coords.numpy()
out['norm_out'] = normalise_coords(coords.clone(), (480, 640)).numpy()
# test/test_utils.py:1082-1101 get_flow_endpoints KAT
f = torch.zeros(1, 2, 3, 4)
f[0, 0] = torch.arange(12).view(3, 4).float() * 0.5
f[0, 1] = -torch.arange(12).view(3, 4).float() * 0.25
for ref in 'st':
    xs, ys = get_flow_endpoints(f, ref)
    out['ep_x_' + ref], out['ep_y_' + ref] = xs.numpy(), ys.numpy()
out['ep_flow'] = f.numpy()
# test/test_flow_operations.py:41-116: combine_flows against analytic flows within the combined mask
shape = (60, 80)
t1 = [['rotation', 20, 30, 10]]
t2 = [['scaling', 40, 30, 1.1]]
for ref in 'st':
    f1 = of.Flow.from_transforms(t1, shape, ref)
    f2 = of.Flow.from_transforms(t2, shape, ref)
    f3 = of.Flow.from_transforms(t1 + t2, shape, ref)
    out['cmb_f1_' + ref], out['cmb_f2_' + ref], out['cmb_f3_' + ref] = f1.vecs.numpy(), f2.vecs.numpy(), f3.vecs.numpy()
    for mode, (a, b) in {1: (f2, f3), 2: (f1, f3), 3: (f1, f2)}.items():
        r = a.combine_with(b, mode)
        out['cmb_%s_%d_vecs' % (ref, mode)], out['cmb_%s_%d_mask' % (ref, mode)] = r.vecs.numpy(), r.mask.numpy()
# flow_class.py:1812-1939 Flow.combine: every mode x ref(self) x ref(other) x ref(result), masked, random smooth flows
import torch.nn.functional as F  # noqa: E402
g = torch.Generator().manual_seed(77)
def smooth(seed):
    gg = torch.Generator().manual_seed(seed)
    lo = torch.randn(2, 2, 4, 5, generator=gg) * 2.5
    return F.interpolate(lo, size=(30, 44), mode='bicubic', align_corners=True).contiguous()
ca, cb = smooth(1), smooth(2)
ma = torch.rand(2, 30, 44, generator=g) > 0.06
mb = torch.rand(2, 30, 44, generator=g) > 0.06
out['cmbx_a'], out['cmbx_b'], out['cmbx_ma'], out['cmbx_mb'] = ca.numpy(), cb.numpy(), ma.numpy(), mb.numpy()
for mode in (1, 2, 3):
    for ra in 'st':
        for rb in 'st':
            for ro in 'st':
                r = of.Flow(ca, ra, ma).combine(of.Flow(cb, rb, mb), mode, ro)
                k = 'cmbx_%d%s%s%s' % (mode, ra, rb, ro)
                out[k + '_v'], out[k + '_m'] = r.vecs.numpy(), r.mask.numpy()
                assert r.ref == ro
# utils.py:941-1042 track_pts / flow_class.py:961-1020 Flow.track: float and integer points, both refs, status
tf = smooth(3)
tm = torch.rand(2, 30, 44, generator=g) > 0.1
pf = torch.stack((torch.rand(2, 17, generator=g) * 29, torch.rand(2, 17, generator=g) * 43), -1)   # (N, M, (y, x))
pi = torch.stack((torch.randint(0, 30, (2, 17), generator=g), torch.randint(0, 44, (2, 17), generator=g)), -1)
gp = torch.rand(2, 17, 2, generator=g)
out['trk_flow'], out['trk_mask'], out['trk_pf'], out['trk_pi'], out['trk_gp'] = tf.numpy(), tm.numpy(), pf.numpy(), pi.numpy(), gp.numpy()
for ref in 'st':
    f2 = tf.clone().requires_grad_()
    p2 = pf.clone().requires_grad_()
    w = of.track_pts(f2, ref, p2)
    w.backward(gp)
    out['trk_%s_f' % ref], out['trk_%s_f_gflow' % ref], out['trk_%s_f_gpts' % ref] = w.detach().numpy(), f2.grad.numpy(), p2.grad.numpy()
    out['trk_%s_i' % ref] = of.track_pts(tf, ref, pi).numpy()
    out['trk_%s_f_int' % ref] = of.track_pts(tf, ref, pf, int_out=True).numpy()
    out['trk_%s_f_2d' % ref] = of.track_pts(tf, ref, pf[0]).numpy()
    wp, st = of.Flow(tf, ref, tm).track(pf, get_valid_status=True)
    out['trk_%s_flowtrack' % ref], out['trk_%s_status' % ref] = wp.numpy(), st.numpy()
np.savez_compressed(os.path.join(HERE, 'kats_v1.npz'), **out)
print('wrote', len(out), 'arrays', os.path.getsize(os.path.join(HERE, 'kats_v1.npz')) // 1024, 'KiB')
