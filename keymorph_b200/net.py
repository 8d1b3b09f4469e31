"""ConvNet backbone of keymorph/net.py:1-36: nine ConvBlocks, channels 1-32-64-64-128-128-256-256-
512-K, pooling after blocks 2/4/6/8.  Same parameter names (block{1..9}.conv.{weight,bias}) so the
reference's checkpoints load with load_state_dict."""
from __future__ import annotations

import torch.nn as nn

from . import layers

h_dims = [32, 64, 64, 128, 128, 256, 256, 512]


class ConvNet(nn.Module):
    def __init__(self, dim, input_ch, out_dim, norm_type):
        super().__init__()
        if dim != 3 or input_ch != 1:
            raise NotImplementedError("keymorph_b200.ConvNet: 3-D, single input channel only")
        self.dim = dim
        chans = [input_ch] + h_dims + [out_dim]
        for b in range(9):
            pool = b in (1, 3, 5, 7)
            setattr(self, f"block{b + 1}",
                    layers.ConvBlock(chans[b], chans[b + 1], 1, norm_type, pool, dim))
        self._engine = None

    def blocks(self):
        return [getattr(self, f"block{b}") for b in range(1, 10)]

    def forward(self, x):
        """Heat map (N,K,D/16,H/16,W/16) fp32, like the reference module."""
        from .engine import backbone_engine
        return backbone_engine(self).heatmap(x)


# name used by the older reference revision quoted in the README / BASELINE north star
ConvNet3D = ConvNet
