"""EEMFlow_cdc with every hot-path op on the sm_100a kernels (convolutions stay on cuDNN).

State-dict compatible with the reference's `EEMFlow_cdc` (model/EEMFlow/EEMFlow+.py:74-234 together with
`cdc_model` model/EEMFlow/cdc_utils.py:105-177): same parameter names and shapes, so the released
checkpoints load with `load_state_dict` (after stripping the DataParallel `module.` prefix, as
test_EEMFlow_HREM.py:62-66 does).  Protocol kept: `change_imagesize((H, W))`, then
`model(events1=..., events2=...) -> ((events1, events2), [5 flows at input resolution])`.

Hot-path ops used: InputPadder (replicate pad), correlation_select (local 9x9 correlation with the
53-channel selection fused), warp (exact convention), WarpingLayer_no_div, upsample2d_flow_as (including
its in-place side effect on the coarse flow), cdc_blend.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .correlation import EEMFLOW_CDC_INDEX, correlation_select
from .warp import InputPadder, WarpingLayer_no_div, cdc_blend, upsample2d_flow_as, warp


def _conv_lrelu(cin, cout, k=3, stride=1, groups=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, 1, groups, bias=True), nn.LeakyReLU(0.1, inplace=True))


def _conv_only(cin, cout, k=3, stride=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=True))


class _Decoder(nn.Module):
    """7-conv flow decoder with grouped middle layers and channel shuffle (EEMFlow+.py:39-71)."""

    def __init__(self, cin, groups):
        super().__init__()
        self.groups = groups
        widths = [(cin, 96, 1), (96, 96, groups), (96, 96, groups), (96, 96, groups), (96, 64, 1), (64, 32, 1)]
        for i, (a, b, g) in enumerate(widths, start=1):
            setattr(self, f"conv{i}", _conv_lrelu(a, b, groups=g))
        self.conv7 = nn.Conv2d(32, 2, 3, 1, 1)

    def _shuffle(self, x):
        b, c, h, w = x.shape
        return x.view(b, self.groups, c // self.groups, h, w).transpose(1, 2).reshape(b, c, h, w)

    def forward(self, x):
        x = self.conv1(x)
        for name in ("conv2", "conv3", "conv4"):
            x = getattr(self, name)(x)
            if self.groups != 1:
                x = self._shuffle(x)
        return self.conv7(self.conv6(self.conv5(x)))


class _DenseEstimator(nn.Module):
    """Densely connected 5-conv estimator producing (inter flow, mask logit) (cdc_utils.py:109-146)."""

    def __init__(self, cin, widths, cout):
        super().__init__()
        n = cin
        for i, wdt in enumerate(widths, start=1):
            setattr(self, f"conv{i}", _conv_lrelu(n, wdt))
            n += wdt
        self.conv_last = _conv_only(n, cout)

    def forward(self, x):
        for i in range(1, 6):
            x = torch.cat([getattr(self, f"conv{i}")(x), x], dim=1)
        return self.conv_last(x)


class _CdcUpsampler(nn.Module):
    """Self-guided flow upsampling (cdc_utils.py:105-177); attribute names follow the reference's state dict."""

    def __init__(self):
        super().__init__()
        self.warping_layer = WarpingLayer_no_div()
        self.dense_estimator_mask = _DenseEstimator(64, (32, 32, 32, 16, 8), 3)
        # present in the reference's state dict, unused by forward()
        self.upsample_output_conv = nn.Sequential(_conv_lrelu(3, 16), _conv_lrelu(16, 16, stride=2), _conv_lrelu(16, 32),
                                                  _conv_lrelu(32, 32, stride=2))

    def forward(self, flow_init, feature_1, feature_2):
        if flow_init.shape[-2:] != feature_1.shape[-2:]:
            flow_init = upsample2d_flow_as(flow_init, feature_1, mode="bilinear", if_rate=True)   # scales its input in place
        feature_2_warp = self.warping_layer(feature_2, flow_init)
        x_out = self.dense_estimator_mask(torch.cat((feature_1, feature_2_warp), dim=1))
        inter_flow = x_out[:, :2].contiguous()
        inter_mask = torch.sigmoid(x_out[:, 2:3]).contiguous()
        return cdc_blend(flow_init, inter_flow, inter_mask)


class EEMFlow_cdc(nn.Module):
    def __init__(self, config=None, groups=3, n_first_channels=15, args=None):
        super().__init__()
        self.args = args
        self.groups = groups
        spec = {"pconv1_1": (n_first_channels, 16, 2), "pconv1_2": (16, 16, 1), "pconv2_1": (16, 32, 2), "pconv2_2": (32, 32, 1),
                "pconv2_3": (32, 32, 1), "pconv3_1": (32, 64, 2), "pconv3_2": (64, 64, 1), "pconv3_3": (64, 64, 1)}
        for name, (a, b, s) in spec.items():
            setattr(self, name, _conv_lrelu(a, b, stride=s))
        self.register_buffer("index", torch.tensor(EEMFLOW_CDC_INDEX), persistent=False)
        for lvl, cin in ((2, 32), (3, 64), (4, 64), (5, 64), (6, 64)):
            setattr(self, f"rconv{lvl}", _conv_lrelu(cin, 32))
        for lvl in (3, 4, 5, 6):      # in the reference's state dict, unused by forward()
            setattr(self, f"up{lvl}", nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True))
        for lvl in (2, 3, 4, 5, 6):
            setattr(self, f"decoder{lvl}", _Decoder(len(EEMFLOW_CDC_INDEX) + 32 + 2, groups))
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self.cdc_model = _CdcUpsampler()
        self.conv_1x1 = nn.ModuleList([_conv_lrelu(c, 32, k=1) for c in (15, 16, 32, 64, 64, 64)])
        self.image_padder = None

    def change_imagesize(self, img_size):
        self.image_size = img_size
        self.image_padder = InputPadder(img_size, mode='chairs', eval_pad_rate=64)

    def _pyramid(self, x):
        f1 = self.pconv1_2(self.pconv1_1(x))
        f2 = self.pconv2_3(self.pconv2_2(self.pconv2_1(f1)))
        f3 = self.pconv3_3(self.pconv3_2(self.pconv3_1(f2)))
        feats = {2: f2, 3: f3}
        for lvl in (4, 5, 6):
            feats[lvl] = F.avg_pool2d(feats[lvl - 1], kernel_size=(2, 2), stride=(2, 2))
        return feats

    def forward(self, events1, events2):
        image1, image2 = self.image_padder.pad(events1, events2)
        p1, p2 = self._pyramid(image1), self._pyramid(image2)
        flows = {}
        f16, f26 = p1[6], p2[6]
        flow7_up = torch.zeros(f16.size(0), 2, f16.size(2), f16.size(3), device=f16.device, dtype=f16.dtype)
        cv = correlation_select(f16, f26, EEMFLOW_CDC_INDEX)
        flows[6] = self.decoder6(torch.cat([cv, self.rconv6(f16), flow7_up], 1))
        for lvl in (5, 4, 3, 2):
            a, b = p1[lvl], p2[lvl]
            proj = self.conv_1x1[lvl]
            flow_up = self.cdc_model(flows[lvl + 1], proj(a), proj(b))
            cv = correlation_select(a, warp(b, flow_up), EEMFLOW_CDC_INDEX)
            feat = getattr(self, f"rconv{lvl}")(a)
            flows[lvl] = getattr(self, f"decoder{lvl}")(torch.cat([cv, feat, flow_up], 1)) + flow_up
        predictions = [upsample2d_flow_as(flows[lvl], events1, mode="bilinear", if_rate=True) for lvl in (6, 5, 4, 3, 2)]
        return (events1, events2), predictions
