"""Tensor-level wrappers of the hot path, API-compatible with the reference's ``flow_operations.py``
(combine_flows :84-188, switch_flow_ref :191-207, invert_flow :210-228, valid_target :231-254,
valid_source :257-277).  They wrap tensors in :class:`Flow` (one validation pass each), call the fused
methods and unwrap -- exactly the reference's structure, minus the re-validation of intermediates."""
from typing import Union

import numpy as np
import torch

from .flow_class import Flow

FlowAlias = 'Flow'


def combine_flows(input_1, input_2, mode: int, ref: str = None, thresholded: bool = None):
    """flow_1 (+) flow_2 = flow_3 solved for the flow named by `mode` (reference flow_operations.py:84-188).
    Differentiable w.r.t. both inputs.  Returns a tensor shaped like `input_1` ((2,H,W) or (N,2,H,W))."""
    if isinstance(input_1, Flow) and isinstance(input_2, Flow):
        print("AVOID - future deprecation warning: using combine_flows(flow_obj1, flow_obj2) is deprecated and may "
              "not work anymore in future versions - use flow_obj1.combine_with(flow_obj2) instead. combine_flows() "
              "will be reserved for use with Torch tensors and NumPy arrays only.")
        return input_1.combine_with(input_2, mode=mode, thresholded=thresholded)
    result = Flow(input_1, ref).combine_with(Flow(input_2, ref), mode=mode, thresholded=thresholded)
    return result.vecs if len(input_1.shape) > 3 else result.vecs.squeeze(0)


def switch_flow_ref(flow: Union[np.ndarray, torch.Tensor], input_ref: str) -> torch.Tensor:
    """reference flow_operations.py:191-207"""
    r = Flow(flow, input_ref).switch_ref()
    return r.vecs if len(flow.shape) > 3 else r.vecs.squeeze(0)


def invert_flow(flow: Union[np.ndarray, torch.Tensor], input_ref: str, output_ref: str = None) -> torch.Tensor:
    """reference flow_operations.py:210-228"""
    output_ref = input_ref if output_ref is None else output_ref
    r = Flow(flow, input_ref).invert(output_ref)
    return r.vecs if len(flow.shape) > 3 else r.vecs.squeeze(0)


def valid_target(flow: Union[np.ndarray, torch.Tensor], ref: str) -> torch.Tensor:
    """reference flow_operations.py:231-254"""
    t = Flow(flow, ref).valid_target()
    return t if len(flow.shape) > 3 else t.squeeze(0)


def valid_source(flow: Union[np.ndarray, torch.Tensor], ref: str) -> torch.Tensor:
    """reference flow_operations.py:257-277"""
    s = Flow(flow, ref).valid_source()
    return s if len(flow.shape) > 3 else s.squeeze(0)


def get_flow_padding(flow: Union[np.ndarray, torch.Tensor], ref: str) -> list:
    """reference flow_operations.py:280-303"""
    p = Flow(flow, ref).get_padding()
    return p if len(flow.shape) > 3 else p[0]
