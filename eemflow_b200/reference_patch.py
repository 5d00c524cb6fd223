"""Install the B200 hot-path implementations under the module names the reference's scripts import.

    import eemflow_b200.reference_patch as rp
    rp.patch_reference("/path/to/EEMFlow")      # before `import test_EEMFlow_HREM` / `train_EEMFlow_HREM`

What gets replaced (reference file:line of the symbol that is shadowed):
  utils.transformers.EventSequenceToVoxelGrid_Pytorch        utils/transformers.py:18
  loader.loader_utils.EventSequenceToVoxelGrid_Pytorch       loader/loader_utils.py:429   (if importable)
  utils_luo.event_utils.EventSequenceToVoxelGrid_Pytorch     utils_luo/event_utils.py:145 (if importable)
  model.corr.CorrBlock                                        model/corr.py:12
  model.model_utils.{bilinear_sampler, upflow8}               model/model_utils.py:7,30
  spatial_correlation_sampler.SpatialCorrelationSampler      pip package used by model/EEMFlow/EEMFlow*.py:9
  utils_luo.tools.tensor_tools.{torch_warp, torch_warp_mask} utils_luo/tools.py:2217,2262 -- that module cannot
      be imported at all (NameError at :1811, SURVEY.md trap 4), so a stand-in module exposing `tensor_tools`
      and an empty `tools` namespace is registered for `from utils_luo.tools import tools, tensor_tools`
  model.EEMFlow.cdc_utils.{WarpingLayer_no_div, upsample2d_flow_as}   model/EEMFlow/cdc_utils.py:50,80
  utils.image_utils.InputPadder                               utils/image_utils.py:126     (if importable)

Modules that fail to import in the current environment (h5py, matplotlib ... missing) are skipped and
reported in the returned dict; nothing in the reference tree is modified on disk.
"""
from __future__ import annotations

import importlib
import sys
import types

from . import corr as _corr
from . import correlation as _correlation
from . import event_utils as _event_utils
import importlib as _importlib

# the package re-exports the *function* `warp`, which shadows the submodule attribute of the same name
_warp = _importlib.import_module(__package__ + ".warp")


def _try_import(name):
    try:
        return importlib.import_module(name)
    except Exception as exc:  # the reference has several unimportable modules; report, do not fail
        return exc


def patch_reference(reference_root: str | None = None) -> dict:
    """Returns {module name: "patched" | "registered" | "skipped: <reason>"}."""
    report = {}
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    # un-vendored pip dependency -> our module
    scs = types.ModuleType("spatial_correlation_sampler")
    scs.SpatialCorrelationSampler = _correlation.SpatialCorrelationSampler
    sys.modules["spatial_correlation_sampler"] = scs
    report["spatial_correlation_sampler"] = "registered"

    # unimportable grab-bag module -> stand-in with the two hot-path classmethods
    if not isinstance(sys.modules.get("utils_luo.tools"), types.ModuleType) or \
            not hasattr(sys.modules["utils_luo.tools"], "tensor_tools"):
        tools_mod = types.ModuleType("utils_luo.tools")
        tools_mod.tensor_tools = _warp.tensor_tools
        tools_mod.tools = type("tools", (), {})
        sys.modules["utils_luo.tools"] = tools_mod
        pkg = _try_import("utils_luo")
        if isinstance(pkg, types.ModuleType):
            pkg.tools = tools_mod
        report["utils_luo.tools"] = "registered"

    def patch(mod_name, **attrs):
        mod = _try_import(mod_name)
        if isinstance(mod, Exception):
            report[mod_name] = f"skipped: {type(mod).__name__}: {mod}"
            return
        for k, v in attrs.items():
            setattr(mod, k, v)
        report[mod_name] = "patched"

    voxel = _event_utils.EventSequenceToVoxelGrid_Pytorch
    patch("utils.transformers", EventSequenceToVoxelGrid_Pytorch=voxel)
    patch("loader.loader_utils", EventSequenceToVoxelGrid_Pytorch=voxel, EventSequence=_event_utils.EventSequence)
    patch("utils_luo.event_utils", EventSequenceToVoxelGrid_Pytorch=voxel, EventSequence=_event_utils.EventSequence)
    patch("model.model_utils", bilinear_sampler=_corr.bilinear_sampler, upflow8=_corr.upflow8)
    patch("model.corr", CorrBlock=_corr.CorrBlock, bilinear_sampler=_corr.bilinear_sampler)
    patch("utils.image_utils", InputPadder=_warp.InputPadder)
    patch("model.EEMFlow.cdc_utils", WarpingLayer_no_div=_warp.WarpingLayer_no_div,
          upsample2d_flow_as=_warp.upsample2d_flow_as)
    return report


def accelerate_eemflow_model(model):
    """Rebind the hot-path members of an already constructed reference model instance in place.

    EEMFlow / EEMFlow_cdc (model/EEMFlow/EEMFlow.py:71, EEMFlow+.py:74): `corr` -> Correlation on the
    local-correlation kernel, `warp` -> exact-convention backward warp, `cdc_model.warping_layer` ->
    WarpingLayer_no_div, `upsample_flow` -> meshflow resize.  Parameters are untouched (these ops
    have none), so checkpoints load as before.
    """
    if hasattr(model, "corr"):
        model.corr = _correlation.Correlation(getattr(model.corr, "max_displacement", 4))
    if hasattr(model, "warp"):
        model.warp = _warp.warp
    if hasattr(model, "upsample_flow"):
        model.upsample_flow = _warp.upsample_flow
    cdc = getattr(model, "cdc_model", None)
    if cdc is not None and hasattr(cdc, "warping_layer"):
        cdc.warping_layer = _warp.WarpingLayer_no_div()
    return model
