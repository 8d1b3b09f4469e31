"""Keypoint layer and ConvNet block of keymorph/layers.py."""
from __future__ import annotations

import torch.nn as nn

from . import ops


class CenterOfMass3d(nn.Module):
    """keymorph/layers.py:78-134: ReLU + centre of mass in [-1,1]^3.  One pass over the heat map
    (km_com3d) instead of three marginal reductions."""

    def __init__(self, indexing="xy") -> None:
        super().__init__()
        assert indexing in ["xy", "ij"]
        self.indexing = indexing

    def forward(self, vol):
        return ops.com3d(vol, ij=(self.indexing == "ij"))


class ConvBlock(nn.Module):
    """Parameter container with the layout of keymorph/layers.py:137-187 (conv -> norm -> ReLU ->
    optional MaxPool).  The arithmetic is executed by keymorph_b200.engine, not by these modules."""

    def __init__(self, in_channels, out_channels, stride, norm_type, down_sample=True, dim=3):
        super().__init__()
        if dim != 3:
            raise NotImplementedError("keymorph_b200 implements the 3-D path only")
        if stride != 1:
            raise NotImplementedError("stride != 1 is not used by the reference ConvNet")
        if norm_type not in ("instance", "none"):
            raise NotImplementedError(f"norm_type {norm_type!r}: only 'instance' (the default of "
                                      "scripts/register.py) and 'none' are implemented")
        self.norm_type = norm_type
        self.down_sample = down_sample
        self.norm = nn.InstanceNorm3d(out_channels) if norm_type == "instance" else None
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1)
        self.down = nn.MaxPool3d(2)
        self.activation = nn.ReLU(out_channels)
