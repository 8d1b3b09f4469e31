"""UNet3D / TruncatedUNet3D parameter containers with the state-dict layout of
keymorph/unet3d/model.py + buildingblocks.py (pytorch-3dunet, layer order "gcr"):

    encoders.{i}.basic_module.SingleConv{1,2}.{groupnorm.weight, groupnorm.bias, conv.weight}
    decoders.{i}.basic_module.SingleConv{1,2}.{...}
    final_conv.{weight, bias}

The modules are created in the same order (and, for the truncated variant, with the same throw-away
decoder / final-conv constructions, keymorph/unet3d/model.py:354-391) as the reference so that a
seeded default initialisation yields bit-identical weights.  The arithmetic is executed by
keymorph_b200.engine on the CUDA kernels; these classes only hold parameters.
"""
from __future__ import annotations

import torch.nn as nn


def number_of_features_per_level(init_channel_number, num_levels):
    """keymorph/unet3d/utils.py:109-110."""
    return [init_channel_number * 2 ** k for k in range(num_levels)]


class SingleConv(nn.Module):
    """GroupNorm -> Conv3d(k3, p1, no bias) -> ReLU (keymorph/unet3d/buildingblocks.py:39-132)."""

    def __init__(self, in_channels, out_channels, num_groups=8):
        super().__init__()
        groups = num_groups if in_channels >= num_groups else 1   # buildingblocks.py:66-68
        assert in_channels % groups == 0
        self.groupnorm = nn.GroupNorm(num_groups=groups, num_channels=in_channels)
        self.conv = nn.Conv3d(in_channels, out_channels, 3, padding=1, bias=False)


class DoubleConv(nn.Module):
    """Channel rule of keymorph/unet3d/buildingblocks.py:171-181."""

    def __init__(self, in_channels, out_channels, encoder, num_groups=8):
        super().__init__()
        if encoder:
            mid = max(out_channels // 2, in_channels)
            c1, c2 = (in_channels, mid), (mid, out_channels)
        else:
            c1, c2 = (in_channels, out_channels), (out_channels, out_channels)
        self.SingleConv1 = SingleConv(*c1, num_groups)
        self.SingleConv2 = SingleConv(*c2, num_groups)


class _Stage(nn.Module):
    def __init__(self, in_channels, out_channels, encoder, num_groups, pool):
        super().__init__()
        self.pool = pool
        self.basic_module = DoubleConv(in_channels, out_channels, encoder, num_groups)


def _encoders(in_channels, f_maps, num_groups):
    return nn.ModuleList(
        _Stage(in_channels if i == 0 else f_maps[i - 1], f, True, num_groups, pool=i > 0)
        for i, f in enumerate(f_maps))


def _decoders(f_maps, num_groups):
    rev = list(reversed(f_maps))
    return nn.ModuleList(
        _Stage(rev[i] + rev[i + 1], rev[i + 1], False, num_groups, pool=False)
        for i in range(len(rev) - 1))


class _UNetBase(nn.Module):
    def __init__(self, in_channels, out_channels, num_truncated_layers, final_sigmoid, f_maps,
                 layer_order, num_groups, num_levels, is_segmentation, conv_padding, truncated,
                 **kwargs):
        super().__init__()
        if layer_order != "gcr" or conv_padding != 1 or in_channels != 1:
            raise NotImplementedError("keymorph_b200 implements layer_order='gcr', conv_padding=1, "
                                      "one input channel (the configuration of scripts/register.py)")
        if is_segmentation:
            raise NotImplementedError("is_segmentation=True (final sigmoid/softmax) is not on the "
                                      "registration path")
        if isinstance(f_maps, int):
            f_maps = number_of_features_per_level(f_maps, num_levels)
        assert len(f_maps) > 1, "Required at least 2 levels in the U-Net"
        self.f_maps = list(f_maps)
        self.num_groups = num_groups
        self.num_truncated_layers = num_truncated_layers
        self.use_checkpoint = kwargs.get("use_checkpoint", False)   # accepted, inference only
        self.final_activation = None
        self.encoders = _encoders(in_channels, self.f_maps, num_groups)
        self.decoders = _decoders(self.f_maps, num_groups)
        self.final_conv = nn.Conv3d(self.f_maps[0], out_channels, 1)
        if truncated:
            # the reference rebuilds the decoder path and the final conv (consuming RNG)
            self.decoders = _decoders(self.f_maps, num_groups)
            if num_truncated_layers > 0:
                self.decoders = self.decoders[:-num_truncated_layers]
            self.final_conv = nn.Conv3d(self.f_maps[num_truncated_layers], out_channels, 1)
        self._engine = None

    def forward(self, x):
        """Heat map (N, K, ...) fp32 as returned by keymorph/unet3d/model.py:115-151."""
        from .engine import backbone_engine
        return backbone_engine(self).heatmap(x)


class UNet3D(_UNetBase):
    """keymorph/unet3d/model.py:154-189."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr",
                 num_groups=8, num_levels=4, is_segmentation=True, conv_padding=1, **kwargs):
        super().__init__(in_channels, out_channels, 0, final_sigmoid, f_maps, layer_order,
                         num_groups, num_levels, is_segmentation, conv_padding, False, **kwargs)


class TruncatedUNet3D(_UNetBase):
    """keymorph/unet3d/model.py:307-430 (drops the last `num_truncated_layers` decoders)."""

    def __init__(self, in_channels, out_channels, num_truncated_layers, final_sigmoid=True,
                 f_maps=64, layer_order="gcr", num_groups=8, num_levels=4, is_segmentation=True,
                 conv_padding=1, **kwargs):
        super().__init__(in_channels, out_channels, num_truncated_layers, final_sigmoid, f_maps,
                         layer_order, num_groups, num_levels, is_segmentation, conv_padding, True,
                         **kwargs)
