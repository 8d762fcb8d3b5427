"""oflibpytorch_b200 -- B200-native (sm_100a) implementation of oflibpytorch's dense flow-field hot path.

Drop-in for: ``apply_flow``, ``combine_flows`` (modes 1/2/3), ``switch_flow_ref``, ``invert_flow``,
``valid_target``, ``valid_source``, ``grid_from_unstructured_data``, ``apply_s_flow`` and the ``Flow`` methods
``apply / combine_with / combine / switch_ref / invert / valid_target / valid_source / track / get_padding /
resize / from_sintel / from_kitti`` -- same (N,2,H,W) tensors, same 't'/'s' reference semantics, padding and mask
behaviour as oflibpytorch 2.1.1.  CUDA tensors only.
"""
from .flow_class import Flow
from .flow_operations import (combine_flows, get_flow_padding, invert_flow, switch_flow_ref, valid_source,
                              valid_target)
from .utils import (DEFAULT_THRESHOLD, apply_flow, apply_s_flow, get_flow_endpoints, get_pure_pytorch,
                    grid_from_unstructured_data, is_zero_flow, load_kitti, load_kitti_batch, load_sintel,
                    load_sintel_batch, load_sintel_mask, resize_flow, set_pure_pytorch, threshold_vectors,
                    track_pts, unset_pure_pytorch, check_deferred, get_validation, set_validation,
                    validation, set_deterministic)

__version__ = '0.1.0'
__all__ = ['Flow', 'combine_flows', 'invert_flow', 'switch_flow_ref', 'valid_source', 'valid_target',
           'apply_flow', 'apply_s_flow', 'get_flow_endpoints', 'grid_from_unstructured_data', 'is_zero_flow',
           'threshold_vectors', 'track_pts', 'get_pure_pytorch', 'set_pure_pytorch', 'unset_pure_pytorch',
           'DEFAULT_THRESHOLD', 'get_flow_padding', 'resize_flow', 'load_sintel', 'load_sintel_batch',
           'load_sintel_mask', 'load_kitti', 'load_kitti_batch', 'set_validation', 'get_validation', 'validation',
           'check_deferred', 'set_deterministic']
