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
from .warp import (InputPadder, WarpingLayer_no_div, blend_warp, cdc_blend, upsample2d_flow_as, upsample2d_flows_as,
                   upsample_warp_no_div, warp)


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

    def forward(self, flow_init, feature_1, feature_2, warp_target=None):
        """cdc_utils.py:156-174.  Returns (flow_up, warped, pre_scale):
        `warped` = EEMFlow_cdc.warp(warp_target, flow_up) when the caller passes the map it is going to warp with the
        result (EEMFlow+.py:181...), fused with the blend; `pre_scale` = the in-place scaling of `flow_init` that the
        reference's upsample2d_flow_as would have applied here and that is deferred to the final upsampling (None when
        no resize was needed)."""
        pre_scale = None
        if flow_init.shape[-2:] != feature_1.shape[-2:]:
            flow_init, feature_2_warp, pre_scale = upsample_warp_no_div(flow_init, feature_1, feature_2)
        else:
            feature_2_warp = self.warping_layer(feature_2, flow_init)
        x_out = self.dense_estimator_mask(torch.cat((feature_1, feature_2_warp), dim=1))
        inter_flow = x_out[:, :2].contiguous()
        inter_mask = torch.sigmoid(x_out[:, 2:3]).contiguous()
        if warp_target is None:
            return cdc_blend(flow_init, inter_flow, inter_mask), None, pre_scale
        flow_up, warped = blend_warp(flow_init, inter_flow, inter_mask, warp_target)
        return flow_up, warped, pre_scale


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
        pre_scales = {}
        for lvl in (5, 4, 3, 2):
            a, b = p1[lvl], p2[lvl]
            proj = self.conv_1x1[lvl]
            # upsample + WarpingLayer_no_div in one launch, blend + warp in one launch; the reference's in-place scaling of
            # flows[lvl + 1] (upsample2d_flow_as side effect) is deferred to the final upsampling below
            flow_up, b_warp, pre_scales[lvl + 1] = self.cdc_model(flows[lvl + 1], proj(a), proj(b), warp_target=b)
            cv = correlation_select(a, b_warp, EEMFLOW_CDC_INDEX)
            feat = getattr(self, f"rconv{lvl}")(a)
            flows[lvl] = getattr(self, f"decoder{lvl}")(torch.cat([cv, feat, flow_up], 1)) + flow_up
        # EEMFlow+.py:231-232: five upsample2d_flow_as calls -> one resize launch + one in-place scaling launch
        order = (6, 5, 4, 3, 2)
        predictions = upsample2d_flows_as([flows[lvl] for lvl in order], events1, mode="bilinear", if_rate=True,
                                          pre_scales=[pre_scales.get(lvl) for lvl in order])
        return (events1, events2), predictions


# ====================================================================================================
# ERAFT: the caller of CorrBlock (model/eraft.py:39-175) with its encoders (model/extractor.py) and
# update block (model/update.py).  State-dict compatible with the reference (same parameter and buffer
# names), so ERAFT checkpoints load unchanged.  The all-pairs volume, its pyramid, the 12 window
# lookups and the padding run in this library's kernels; convolutions, norms and the GRU stay on cuDNN.
# ====================================================================================================
def _norm(kind, planes):
    if kind == "batch":
        return nn.BatchNorm2d(planes)
    if kind == "instance":
        return nn.InstanceNorm2d(planes)
    if kind == "group":
        return nn.GroupNorm(num_groups=planes // 8, num_channels=planes)
    return nn.Sequential()


class _ResBlock(nn.Module):
    def __init__(self, cin, planes, kind, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = _norm(kind, planes), _norm(kind, planes)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(kind, planes)      # registered twice (norm3 and downsample.1), as in the reference
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes, 1, stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class BasicEncoder(nn.Module):
    """1/8-resolution feature encoder (model/extractor.py:114-190); a list input is batched through once."""

    def __init__(self, output_dim=128, norm_fn='batch', dropout=0.0, n_first_channels=1):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = nn.GroupNorm(8, 64) if norm_fn == "group" else _norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(n_first_channels, 64, 7, 2, 3)
        self.relu1 = nn.ReLU(inplace=True)
        widths, cin = ((64, 1), (96, 2), (128, 2)), 64
        for k, (planes, stride) in enumerate(widths, start=1):
            setattr(self, f"layer{k}", nn.Sequential(_ResBlock(cin, planes, norm_fn, stride), _ResBlock(planes, planes, norm_fn, 1)))
            cin = planes
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        parts = None
        if isinstance(x, (tuple, list)):
            parts = [t.shape[0] for t in x]
            x = torch.cat(x, dim=0)
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, parts, dim=0) if parts is not None else x


class _FlowHead(nn.Module):
    def __init__(self, cin=128, hidden=256):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(cin, hidden, 3, padding=1), nn.Conv2d(hidden, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


class _SepConvGRU(nn.Module):
    """Horizontal (1x5) then vertical (5x1) GRU update (model/update.py:35-62)."""

    def __init__(self, hidden=128, cin=192 + 128):
        super().__init__()
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(hidden + cin, hidden, k, padding=p))

    def forward(self, h, x):
        for tag in "12":
            hx = torch.cat([h, x], dim=1)
            z = torch.sigmoid(getattr(self, "convz" + tag)(hx))
            r = torch.sigmoid(getattr(self, "convr" + tag)(hx))
            q = torch.tanh(getattr(self, "convq" + tag)(torch.cat([r * h, x], dim=1)))
            h = (1 - z) * h + z * q
        return h


class _MotionEncoder(nn.Module):
    def __init__(self, cor_planes):
        super().__init__()
        self.convc1, self.convc2 = nn.Conv2d(cor_planes, 256, 1), nn.Conv2d(256, 192, 3, padding=1)
        self.convf1, self.convf2 = nn.Conv2d(2, 128, 7, padding=3), nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc2(F.relu(self.convc1(corr))))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        return torch.cat([F.relu(self.conv(torch.cat([cor, flo], dim=1))), flow], dim=1)


class BasicUpdateBlock(nn.Module):
    def __init__(self, args, hidden_dim=128, input_dim=128):
        super().__init__()
        self.args = args
        self.encoder = _MotionEncoder(args.corr_levels * (2 * args.corr_radius + 1) ** 2)
        self.gru = _SepConvGRU(hidden_dim, 128 + hidden_dim)
        self.flow_head = _FlowHead(hidden_dim, 256)
        self.mask = nn.Sequential(nn.Conv2d(hidden_dim, hidden_dim * 2, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(hidden_dim * 2, 64 * 9, 1))

    def forward(self, net, inp, corr, flow, upsample=True):
        net = self.gru(net, torch.cat([inp, self.encoder(flow, corr)], dim=1))
        return net, .25 * self.mask(net), self.flow_head(net)


class ERAFT(nn.Module):
    """model/eraft.py:39-175.  `corr_precision` ("tf32" | "fp32" | None = library default) is the one extra knob."""

    def __init__(self, config=None, n_first_channels=5, corr_precision=None):
        super().__init__()
        from argparse import Namespace
        self.args = Namespace(small=False, dropout=False, mixed_precision=False, clip=1.0, corr_levels=4, corr_radius=4)
        self.hidden_dim = self.context_dim = 128
        self.corr_precision = corr_precision
        self.fnet = BasicEncoder(256, 'instance', 0, n_first_channels)
        self.cnet = BasicEncoder(self.hidden_dim + self.context_dim, 'batch', 0, n_first_channels)
        self.update_block = BasicUpdateBlock(self.args, hidden_dim=self.hidden_dim)

    def change_imagesize(self, img_size):
        self.image_size = img_size
        self.image_padder = InputPadder(img_size, mode='chairs')

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def initialize_flow(self, img):
        from .corr import coords_grid
        n, _, h, w = img.shape
        return coords_grid(n, h // 8, w // 8, device=img.device), coords_grid(n, h // 8, w // 8, device=img.device)

    def upsample_flow(self, flow, mask):
        """[N,2,H/8,W/8] -> [N,2,H,W] by the learned convex combination of the 3x3 coarse neighbours."""
        n, _, h, w = flow.shape
        mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
        nb = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
        return torch.sum(mask * nb, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w)

    def forward(self, events1, events2, iters=12, flow_init=None, upsample=True, normal=False):
        from .corr import CorrBlock, upflow8
        image1, image2 = self.image_padder.pad(events1, events2)
        fmap1, fmap2 = self.fnet([image1.contiguous(), image2.contiguous()])
        corr_fn = CorrBlock(fmap1.float(), fmap2.float(), radius=self.args.corr_radius, precision=self.corr_precision)
        net, inp = torch.split(self.cnet(image1), [self.hidden_dim, self.context_dim], dim=1)
        net, inp = torch.tanh(net), torch.relu(inp)
        coords0, coords1 = self.initialize_flow(image1)
        if flow_init is not None:
            coords1 = coords1 + flow_init
        flow_predictions = []
        for _ in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1)
            net, up_mask, delta_flow = self.update_block(net, inp, corr, coords1 - coords0)
            coords1 = coords1 + delta_flow
            flow_up = upflow8(coords1 - coords0) if up_mask is None else self.upsample_flow(coords1 - coords0, up_mask)
            flow_predictions.append(self.image_padder.unpad(flow_up))
        return (events1, events2), flow_predictions


class GraphedInference:
    """Replay a model's inference forward as ONE CUDA graph per input shape.

    At small batch the reference-shaped forward is launch-bound (EEMFlow_cdc: ~150 kernels, 5 ms per MVSEC pair
    eager on a B200 for ~1 ms of device work).  The first call with a new (shape, dtype, kwargs) runs the model a
    few times on a side stream (cuDNN autotuning, scratch allocation), captures one forward, and later calls copy
    the inputs into the captured buffers and replay.  Inference only (`torch.no_grad`); the returned flow tensors
    are the graph's output buffers and stay valid until the next call with the same shape.
    """

    def __init__(self, model: nn.Module, warmup: int = 3):
        self.model = model
        self.warmup = warmup
        self._captured = {}

    def change_imagesize(self, img_size):
        self.model.change_imagesize(img_size)

    def __call__(self, events1, events2, **kwargs):
        for name, value in kwargs.items():
            if isinstance(value, torch.Tensor):
                # a tensor would be baked into the graph by address: replays would silently reuse stale data
                raise TypeError(f"GraphedInference: tensor keyword argument `{name}` is not supported "
                                "(e.g. ERAFT's flow_init warm start); call the model directly for that")
        if self.model.training:
            raise RuntimeError("GraphedInference replays an inference forward: call model.eval() first "
                               "(a train-mode forward would update BatchNorm statistics on every warm-up and replay)")
        key = (tuple(events1.shape), events1.dtype, events1.device, tuple(sorted(kwargs.items())), getattr(self.model, "image_size", None))
        entry = self._captured.get(key)
        if entry is None:
            static1, static2 = events1.clone(), events2.clone()
            side = torch.cuda.Stream(events1.device)
            side.wait_stream(torch.cuda.current_stream(events1.device))
            with torch.no_grad(), torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self.model(events1=static1, events2=static2, **kwargs)
            torch.cuda.current_stream(events1.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                _, flows = self.model(events1=static1, events2=static2, **kwargs)
            entry = self._captured[key] = (graph, static1, static2, flows)
        graph, static1, static2, flows = entry
        static1.copy_(events1, non_blocking=True)
        static2.copy_(events2, non_blocking=True)
        graph.replay()
        return (events1, events2), flows
