"""keymorph_b200 -- B200-native (sm_100a) engine behind the Python surface of alanqrwang/keymorph.

The names below mirror the reference modules on the registration hot path
(keymorph.model / keypoint_aligners / transformations / layers / net / unet3d.model / utils /
loss_ops / augmentation); everything they compute runs in hand-written CUDA kernels reached through the C ABI of
libkm_b200.so (include/km_b200.h).  There is no CPU path.
"""
from . import evaluation, loss_ops, ops  # noqa: F401
from .augmentation import (AffineDeformation3d, affine_augment, random_affine_augment,  # noqa: F401
                           random_affine_augment_pair)
from .keypoint_aligners import (TPS, AffineKeypointAligner, RigidKeypointAligner,  # noqa: F401
                                grid_from_points)
from .layers import CenterOfMass3d, ConvBlock  # noqa: F401
from .loss_ops import DiceLoss, MSELoss  # noqa: F401
from .model import KeyMorph  # noqa: F401
from .net import ConvNet, ConvNet3D  # noqa: F401
from .ops import act_dtype, set_operand_dtype  # noqa: F401
from .transformations import AffineTransform  # noqa: F401
from .unet3d import TruncatedUNet3D, UNet3D  # noqa: F401
from .utils import align_img, align_moving_img, one_hot, uniform_norm_grid  # noqa: F401

__version__ = "0.1.0"
