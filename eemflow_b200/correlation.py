"""Drop-in for the local cost volume EEMFlow computes through `spatial_correlation_sampler`.

Mirrors:
  * spatial_correlation_sampler.SpatialCorrelationSampler (pip package, pinned 0.4.0 in the
    reference's requirements.txt:131; call sites model/EEMFlow/EEMFlow.py:19,23 and EEMFlow+.py:21,25)
  * Correlation                     model/EEMFlow/EEMFlow.py:14-23 == EEMFlow+.py:16-25
plus `correlation_select`, the fused form of `index_select(corr(f1, f2), 1, index)`
(EEMFlow.py:160, EEMFlow+.py:178,190,202,214,226).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import autograd as ag


class SpatialCorrelationSampler(nn.Module):
    """`SpatialCorrelationSampler(kernel_size, patch_size, stride, padding, dilation)` -> [B, ph, pw, H, W].

    Only the configuration the reference uses is implemented in CUDA: kernel_size=1, patch_size=9,
    stride=1, padding=0, dilation=1, dilation_patch=1.  Anything else raises NotImplementedError.
    """

    def __init__(self, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.patch_size = patch_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.dilation_patch = dilation_patch
        if (kernel_size, stride, padding, dilation, dilation_patch) != (1, 1, 0, 1, 1) or patch_size != 9:
            raise NotImplementedError(
                "eemflow_b200 implements SpatialCorrelationSampler(1, 9, 1, 0, 1) (the EEMFlow configuration) only")

    def forward(self, input1, input2):
        b, c, h, w = input1.shape
        md = (self.patch_size - 1) // 2
        out = ag.local_corr(input1, input2, max_disp=md)
        return out.view(b, self.patch_size, self.patch_size, h, w)


class Correlation(nn.Module):
    def __init__(self, max_displacement):
        super(Correlation, self).__init__()
        self.max_displacement = max_displacement
        self.kernel_size = 2 * max_displacement + 1
        self.corr = SpatialCorrelationSampler(1, self.kernel_size, 1, 0, 1)

    def forward(self, x, y):
        b, c, h, w = x.shape
        # "/ c" of the reference folded into the kernel's store
        return ag.local_corr(x, y, max_disp=self.max_displacement, scale=1.0 / c)

    def forward_select(self, x, y, index):
        return correlation_select(x, y, index, self.max_displacement)


def correlation_select(x, y, index, max_displacement=4):
    """index_select(Correlation(md)(x, y), dim=1, index) in one kernel: only the kept channels are written."""
    idx = index.tolist() if torch.is_tensor(index) else list(index)
    b, c, h, w = x.shape
    return ag.local_corr(x, y, max_disp=max_displacement, index=[int(v) for v in idx], scale=1.0 / c)


# The two fixed channel lists of the reference.
EEMFLOW_INDEX = [1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 22, 23, 25, 27, 29, 30, 31, 32, 33, 35, 37, 38, 39, 40, 41,
                 42, 43, 45, 47, 48, 49, 50, 51, 53, 55, 57, 58, 59, 61, 63, 65, 67, 69, 71, 73, 75, 77, 79]  # EEMFlow.py:85-94
EEMFLOW_CDC_INDEX = [0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 21, 22, 23, 24, 26, 28, 29, 30, 31, 32, 33, 34, 36, 38,
                     39, 40, 41, 42, 44, 46, 47, 48, 49, 50, 51, 52, 54, 56, 57, 58, 59, 60, 62, 64, 66, 68, 70, 72,
                     74, 76, 78, 80]  # EEMFlow+.py:89-97
