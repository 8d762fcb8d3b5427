"""Generates tests/golden/ingest_v1.npz and the small data files under tests/golden/files/: inputs and the
reference's outputs for get_padding, resize_flow / Flow.resize and the Sintel / KITTI loaders.
The data files are written by THIS script (synthetic content), then read back through the reference's loaders.
Build container only (imports the reference)."""
import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, '/root/reference/src')
import oflibpytorch as of  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = os.path.join(HERE, 'files')
os.makedirs(FILES, exist_ok=True)
of.set_pure_pytorch()
out = {}
g = torch.Generator().manual_seed(1234)


def smooth(n, h, w, sigma):
    lo = torch.randn(n, 2, h // 8 + 2, w // 8 + 2, generator=g) * sigma
    return torch.nn.functional.interpolate(lo, size=(h, w), mode='bicubic', align_corners=True)


# ---- get_padding: the reference's own 7x7 KAT (test/test_flow_class.py:1619-1647) plus smooth random fields
rot = [['rotation', 0, 0, 45]]
m_s = np.ones((7, 7), 'bool'); m_s[:, 4:] = False
m_t = np.ones((7, 7), 'bool'); m_t[4:] = False
for name, fl in (('s', of.Flow.from_transforms(rot, (7, 7), 's')), ('t', of.Flow.from_transforms(rot, (7, 7), 't')),
                 ('s_masked', of.Flow.from_transforms(rot, (7, 7), 's', m_s)),
                 ('t_masked', of.Flow.from_transforms(rot, (7, 7), 't', m_t))):
    out['pad_rot45_%s_v' % name], out['pad_rot45_%s_m' % name] = fl.vecs.numpy(), fl.mask.numpy()
    out['pad_rot45_%s' % name] = np.array(fl.get_padding())
v = smooth(3, 40, 56, 6.0)
v[1] *= 1e-4                                                     # below the threshold: no padding at all
m = torch.rand(3, 40, 56, generator=g) > 0.3
out['pad_rnd_v'], out['pad_rnd_m'] = v.numpy(), m.numpy()
for ref in 'st':
    out['pad_rnd_' + ref] = np.array(of.Flow(v, ref, m).get_padding())
    out['pad_rnd_nomask_' + ref] = np.array(of.Flow(v, ref).get_padding())

# ---- resize: the scales of test/test_utils.py:855-889 and :891-899, values + gradient + rounded mask
fv = of.Flow.from_transforms([['rotation', 30, 50, 30]], [20, 10], 's').vecs
fv = torch.cat((fv, smooth(1, 20, 10, 3.0)))
fm = torch.rand(2, 20, 10, generator=g) > 0.35
out['rsz_v'], out['rsz_m'] = fv.numpy(), fm.numpy()
scales = [.2, .5, 1, 1.5, 2, 10, [.5, 2], (2, .5), 0.7, [1.37, 0.61]]
out['rsz_scales'] = np.array([[s, s] if isinstance(s, (int, float)) else list(s) for s in scales], dtype=np.float64)
for i, s in enumerate(scales):
    x = fv.clone().requires_grad_()
    r = of.resize_flow(x, s)
    go = torch.rand(r.shape, generator=g)
    r.backward(go)
    out['rsz_%d_out' % i], out['rsz_%d_gout' % i], out['rsz_%d_gin' % i] = r.detach().numpy(), go.numpy(), x.grad.numpy()
    # the reference's Flow.resize squeezes the mask with .squeeze(0).squeeze(0) and so only accepts N = 1
    # (flow_class.py:707-712): one item at a time
    out['rsz_%d_mask' % i] = np.concatenate([of.Flow(fv[k], 't', fm[k]).resize(s).mask.numpy() for k in range(2)])
big = of.Flow.from_transforms([['rotation', 0, 0, 30]], (150, 240), 't')
out['rsz_big_v'] = big.vecs.numpy()
out['rsz_big_out'] = of.resize_flow(big.vecs, 1 / 3).numpy()
out['rsz_small_v'] = of.Flow.from_transforms([['rotation', 0, 0, 30]], (50, 80), 't').vecs.numpy()
r3 = of.resize_flow(fv[0], 1.5)                                   # 3-D input keeps 3 dimensions
out['rsz_3d_out'] = r3.numpy()

# ---- loaders: files with synthetic content written here, read back by the reference
for k in range(2):
    payload = (torch.randn(13, 17, 2, generator=g) * 5).numpy().astype('<f4')
    with open(os.path.join(FILES, 'sample%d.flo' % k), 'wb') as fh:
        fh.write(b'PIEH' + (17).to_bytes(4, 'little') + (13).to_bytes(4, 'little') + payload.tobytes())
    out['flo%d' % k] = of.load_sintel(os.path.join(FILES, 'sample%d.flo' % k)).numpy()
with open(os.path.join(FILES, 'wrong_tag.flo'), 'wb') as fh:
    fh.write(b'PIEX' + (2).to_bytes(4, 'little') + (2).to_bytes(4, 'little') + bytes(32))
inv = (torch.rand(13, 17, generator=g) > 0.7).numpy().astype(np.uint8) * 255
cv2.imwrite(os.path.join(FILES, 'sample_invalid.png'), inv)
out['flo_mask'] = of.load_sintel_mask(os.path.join(FILES, 'sample_invalid.png')).numpy()
fl = of.Flow.from_sintel(os.path.join(FILES, 'sample0.flo'), os.path.join(FILES, 'sample_invalid.png'))
out['from_sintel_v'], out['from_sintel_m'] = fl.vecs.numpy(), fl.mask.numpy()
for k in range(2):
    png = torch.randint(0, 65536, (11, 19, 3), generator=g).numpy().astype(np.uint16)
    png[..., 0] = (torch.randint(0, 3, (11, 19), generator=g).numpy() * 7).astype(np.uint16)   # valid: 0, 7 or 14
    cv2.imwrite(os.path.join(FILES, 'sample%d_kitti.png' % k), png)
    out['kitti%d' % k] = of.load_kitti(os.path.join(FILES, 'sample%d_kitti.png' % k)).numpy()
cv2.imwrite(os.path.join(FILES, 'wrong_kitti.png'), np.zeros((5, 6), np.uint16))
fl = of.Flow.from_kitti(os.path.join(FILES, 'sample0_kitti.png'))
out['from_kitti_v'], out['from_kitti_m'] = fl.vecs.numpy(), fl.mask.numpy()

np.savez_compressed(os.path.join(HERE, 'ingest_v1.npz'), **out)
print('wrote', len(out), 'arrays;', sorted(os.listdir(FILES)))
