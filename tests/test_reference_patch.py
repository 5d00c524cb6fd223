"""CPU test: patch_reference() installs the drop-ins under the reference's own module names.
Needs the reference tree (build container only); skipped on the GPU box where it does not exist."""
import sys
import types
from pathlib import Path

import pytest

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")


def test_patch_reference_rebinds_hot_path_symbols():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "imageio", "png"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.colors"].hsv_to_rgb = None
    import eemflow_b200 as E
    from eemflow_b200 import reference_patch as rp
    saved = {k: sys.modules.get(k) for k in ("utils_luo.tools", "spatial_correlation_sampler")}
    try:
        report = rp.patch_reference(str(REF))
        assert report["spatial_correlation_sampler"] == "registered"
        import model.corr
        import model.model_utils
        import utils.transformers
        assert model.corr.CorrBlock is E.CorrBlock
        assert model.model_utils.bilinear_sampler is E.bilinear_sampler
        assert utils.transformers.EventSequenceToVoxelGrid_Pytorch is E.EventSequenceToVoxelGrid_Pytorch
        from utils_luo.tools import tensor_tools, tools  # noqa: F401  (unimportable in the stock reference)
        assert tensor_tools.torch_warp.__func__ is E.tensor_tools.torch_warp.__func__
        import model.EEMFlow.cdc_utils as cdc
        assert cdc.WarpingLayer_no_div is E.WarpingLayer_no_div and cdc.upsample2d_flow_as is E.upsample2d_flow_as
        # the reference's own model file now imports and constructs against the drop-ins
        import importlib.util
        spec = importlib.util.spec_from_file_location("eemflow_plus_patched", REF / "model" / "EEMFlow" / "EEMFlow+.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        net = mod.EEMFlow_cdc(None, groups=3, n_first_channels=5)
        assert isinstance(net.corr.corr, E.SpatialCorrelationSampler)
        rp.accelerate_eemflow_model(net)
        assert isinstance(net.corr, E.Correlation) and net.warp is E.warp
        assert isinstance(net.cdc_model.warping_layer, E.WarpingLayer_no_div)
        net.change_imagesize((260, 346))
        assert isinstance(net.image_padder, E.InputPadder) and net.image_padder._pad == [19, 19, 0, 60]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
