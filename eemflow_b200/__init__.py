"""eemflow_b200 -- B200 (sm_100a) kernels behind the hot-path interfaces of boomluo02/EEMFlow.

Public surface (same names and call signatures as the reference):
    EventSequence, EventSequenceToVoxelGrid_Pytorch         event_utils.py
    CorrBlock, bilinear_sampler, coords_grid, upflow8        corr.py
    SpatialCorrelationSampler, Correlation                   correlation.py
    warp, tensor_tools.torch_warp(_mask), WarpingLayer_no_div,
    upsample2d_flow_as, upsample2d_flows_as (all predictions in one launch), upsample_flow, cdc_blend, InputPadder   warp.py
    event_mask, event_valid_from_volume, flow_error, motion_propagate, center_crop   eval_utils.py
    EEMFlow_cdc                                              models.py (lazy: from eemflow_b200.models import ...)
Everything computes in hand-written CUDA kernels reached through the C ABI of
libeemflow_b200.so (include/eemflow_b200.h); importing the package does not touch CUDA.
"""
from .corr import CorrBlock, bilinear_sampler, coords_grid, upflow8
from .correlation import Correlation, SpatialCorrelationSampler, correlation_select
from .eval_utils import center_crop, event_mask, event_valid_from_volume, flow_error, motion_propagate
from .event_utils import EventSequence, EventSequenceToVoxelGrid_Pytorch
from .ops import local_corr_precision, set_local_corr_precision
from .warp import (InputPadder, WarpingLayer_no_div, blend_warp, cdc_blend, tensor_tools, torch_warp, torch_warp_mask,
                   upsample2d_flow_as, upsample2d_flows_as, upsample_flow, upsample_warp_no_div, warp)

__all__ = [
    "EventSequence", "EventSequenceToVoxelGrid_Pytorch", "CorrBlock", "bilinear_sampler", "coords_grid", "upflow8",
    "SpatialCorrelationSampler", "Correlation", "correlation_select", "warp", "tensor_tools", "torch_warp",
    "torch_warp_mask", "WarpingLayer_no_div", "upsample2d_flow_as", "upsample2d_flows_as", "upsample_flow",
    "upsample_warp_no_div", "blend_warp", "cdc_blend", "InputPadder",
    "event_mask", "event_valid_from_volume", "flow_error", "motion_propagate", "center_crop",
    "set_local_corr_precision", "local_corr_precision",
]
__version__ = "0.1.0"
