"""Host-side behaviour that needs no GPU: the validation/error contract of the boundary (reference utils.py:72-232,
flow_class.py:821-875,1714-1726), the Flow container, and the refusal to run the hot path on CPU tensors."""
import numpy as np
import pytest
import torch

import oflibpytorch_b200 as of
from oflibpytorch_b200 import utils as U


def test_get_valid_vecs_contract():
    with pytest.raises(TypeError):
        U.get_valid_vecs('test')
    with pytest.raises(ValueError):
        U.get_valid_vecs(np.zeros((2, 10)))
    with pytest.raises(ValueError):
        U.get_valid_vecs(np.zeros((3, 10, 10)))
    v = np.zeros((2, 10, 12))
    v[0, 3, 4] = np.nan
    with pytest.raises(ValueError):
        U.get_valid_vecs(v)
    v[0, 3, 4] = np.inf
    with pytest.raises(ValueError):
        U.get_valid_vecs(torch.tensor(v))
    with pytest.raises(ValueError):
        U.get_valid_vecs(np.zeros((2, 10, 12)), desired_shape=(10, 13))
    out = U.get_valid_vecs(np.ones((10, 12, 2)))
    assert out.shape == (1, 2, 10, 12) and out.dtype == torch.float32
    out = U.get_valid_vecs(torch.ones(3, 10, 12, 2, dtype=torch.float64))
    assert out.shape == (3, 2, 10, 12) and out.dtype == torch.float32


def test_get_valid_ref_mask_device_padding_shape():
    assert U.get_valid_ref(None) == 't' and U.get_valid_ref('s') == 's'
    with pytest.raises(TypeError):
        U.get_valid_ref(0)
    with pytest.raises(ValueError):
        U.get_valid_ref('x')
    with pytest.raises(TypeError):
        U.get_valid_mask('m')
    with pytest.raises(ValueError):
        U.get_valid_mask(np.ones((2, 3, 4, 5)))
    with pytest.raises(ValueError):
        U.get_valid_mask(np.full((4, 5), 2.0))
    with pytest.raises(ValueError):
        U.get_valid_mask(np.ones((4, 5)), desired_shape=(4, 6))
    m = U.get_valid_mask(np.ones((4, 5)))
    assert m.shape == (1, 4, 5) and m.dtype == torch.bool
    assert U.get_valid_device(None) == torch.device('cpu')
    with pytest.raises(ValueError):
        U.get_valid_device('not-a-device')
    if not torch.cuda.is_available():
        with pytest.raises(ValueError):
            U.get_valid_device('cuda')
    with pytest.raises(TypeError):
        U.get_valid_padding(5)
    with pytest.raises(ValueError):
        U.get_valid_padding([1, 2, 3])
    with pytest.raises(ValueError):
        U.get_valid_padding([1, 2, 3, 4.0])
    with pytest.raises(ValueError):
        U.get_valid_padding([1, 2, 3, -4])
    with pytest.raises(TypeError):
        U.get_valid_shape('s')
    with pytest.raises(ValueError):
        U.get_valid_shape([1, 2, 3, 4])
    with pytest.raises(ValueError):
        U.get_valid_shape([0, 2])
    assert U.get_valid_shape([3, 4]) == (1, 3, 4)


def test_flow_container_and_arithmetic_cpu():
    v = torch.rand(2, 2, 6, 7)
    f = of.Flow(v, 's')
    assert f.shape == (2, 6, 7) and f.ref == 's' and f.device == torch.device('cpu')
    assert f.mask.dtype == torch.bool and bool(f.mask.all())
    g = of.Flow(np.ones((6, 7, 2)), mask=np.ones((6, 7)))
    assert g.shape == (1, 6, 7) and g.ref == 't'
    with pytest.raises(ValueError):
        of.Flow(v, mask=torch.ones(2, 6, 8))
    with pytest.raises(ValueError):
        of.Flow(v, 'x')
    with pytest.raises(TypeError):
        of.Flow(v, 3)
    s = f + of.Flow(v, 's', torch.zeros(2, 6, 7, dtype=torch.bool))
    assert torch.equal(s.vecs, v + v) and not bool(s.mask.any())
    assert torch.equal((f - f).vecs, torch.zeros_like(v))
    assert torch.equal((f * 2).vecs, v * 2) and torch.equal((f / 2).vecs, v / 2) and torch.equal((-f).vecs, -v)
    assert torch.equal((f * [1, 2]).vecs[:, 1], v[:, 1] * 2)
    with pytest.raises(ValueError):
        f * [1, 2, 3]
    with pytest.raises(TypeError):
        f + 'x'
    with pytest.raises(ValueError):
        f + torch.rand(2, 2, 6, 8)
    assert f.select(1).shape == (1, 6, 7)
    with pytest.raises(TypeError):
        f.select('a')
    with pytest.raises(IndexError):
        f.select(5)
    assert f[1:4, 2:5].shape == (2, 3, 3)
    p = f.pad([1, 2, 3, 4])
    assert p.shape == (2, 9, 14) and not bool(p.mask[:, 0].any())
    assert p.unpad([1, 2, 3, 4]).shape == (2, 6, 7)
    with pytest.raises(ValueError):
        f.pad([1, 2, 3, 4], mode='x')
    z = of.Flow.zero((4, 5))
    assert z.shape == (1, 4, 5) and float(z.vecs.abs().sum()) == 0.0
    assert f.copy().shape == f.shape and f.switch_ref('invalid').ref == 't'
    with pytest.raises(ValueError):
        f.switch_ref('x')


def test_apply_argument_errors_cpu():
    """reference test/test_flow_class.py:1099-1141, test/test_utils.py:587-602: exception TYPES at the boundary."""
    f = of.Flow(torch.rand(1, 2, 6, 7))
    img = torch.rand(3, 6, 7)
    with pytest.raises(TypeError):
        f.apply(img, return_valid_area='x')
    with pytest.raises(TypeError):
        f.apply(img, consider_mask='x')
    with pytest.raises(TypeError):
        f.apply(img, cut='x')
    with pytest.raises(TypeError):
        f.apply('x')
    with pytest.raises(ValueError):
        f.apply(torch.rand(5))
    with pytest.raises(ValueError):
        f.apply(torch.rand(3, 6, 8))
    with pytest.raises(TypeError):
        f.apply(img, target_mask='x', return_valid_area=True)
    with pytest.raises(ValueError):
        f.apply(img, target_mask=torch.ones(6, 8, dtype=torch.bool), return_valid_area=True)
    with pytest.raises(TypeError):
        f.apply(img, target_mask=torch.ones(6, 7), return_valid_area=True)
    with pytest.raises(TypeError):
        f.apply(torch.rand(3, 8, 9), padding=5)
    with pytest.raises(ValueError):
        f.apply(torch.rand(3, 8, 9), padding=[1, 2, 3, 4])
    with pytest.raises(TypeError):
        f.combine_with(img, 3)
    with pytest.raises(ValueError):
        f.combine_with(of.Flow(torch.rand(1, 2, 6, 8)), 3)
    with pytest.raises(ValueError):
        f.combine_with(of.Flow(torch.rand(1, 2, 6, 7), 's'), 3)
    with pytest.raises(ValueError):
        f.combine_with(of.Flow(torch.rand(1, 2, 6, 7)), 4)
    with pytest.raises(TypeError):
        f.combine_with(of.Flow(torch.rand(1, 2, 6, 7)), 3, thresholded='x')
    with pytest.raises(TypeError):
        f.valid_target(consider_mask='x')
    with pytest.raises(TypeError):
        f.valid_source(consider_mask='x')
    with pytest.raises(TypeError):
        of.apply_flow(torch.rand(2, 6, 7), 'x', 't') if torch.cuda.is_available() else of.apply_flow('x', img, 't')
    with pytest.raises(ValueError):
        of.apply_flow(torch.rand(2, 6, 7), img, 'q')
    with pytest.raises(TypeError):
        of.apply_flow(torch.rand(2, 6, 7), img, 3)


def test_hot_path_refuses_cpu_tensors_loudly():
    """No CPU fallback: validated CPU inputs end in a RuntimeError that says so."""
    f = of.Flow(torch.rand(1, 2, 6, 7))
    img = torch.rand(3, 6, 7)
    for call in (lambda: f.apply(img), lambda: f.switch_ref(), lambda: f.valid_target(), lambda: f.is_zero(),
                 lambda: of.apply_flow(torch.rand(2, 6, 7), img, 't'),
                 lambda: of.grid_from_unstructured_data(torch.rand(1, 6, 7), torch.rand(1, 6, 7), torch.rand(1, 2, 6, 7)),
                 lambda: of.combine_flows(torch.rand(2, 6, 7), torch.rand(2, 6, 7), 3, 't')):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            call()


def test_scipy_mode_is_out_of_scope():
    assert of.get_pure_pytorch() is True
    of.set_pure_pytorch()
    with pytest.raises(NotImplementedError):
        of.unset_pure_pytorch()


def test_threshold_vectors_matches_reference_rule():
    """reference utils.py:623-643 / test/test_utils.py:605-625"""
    v = torch.tensor([[[[0.0, 5e-4, 1e-3, -9.99e-4, -1e-3, 2.0]], [[0.5, 0.0, 0.0, 0.0, 0.0, 0.0]]]])
    out = of.threshold_vectors(v)
    assert out[0, 0, 0].tolist() == [0.0, 0.0, pytest.approx(1e-3), 0.0, pytest.approx(-1e-3), 2.0]
    assert out[0, 1, 0, 0] == 0.5
