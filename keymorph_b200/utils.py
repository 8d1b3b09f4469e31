"""Host-side mirror of the hot-path helpers of keymorph/utils.py (same names, same arguments)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def str_or_float(x):
    """keymorph/utils.py:7-11."""
    try:
        return float(x)
    except ValueError:
        return x


def align_img(grid, x, mode="bilinear"):
    """keymorph/utils.py:14-21: F.grid_sample(x, grid, mode, padding_mode='border',
    align_corners=False) -- here the hand-written gather kernel (km_grid_sample3d)."""
    if x.dim() != 5:
        raise NotImplementedError("keymorph_b200.align_img supports 3-D volumes (N,C,D,H,W) only")
    return ops.grid_sample3d(x, grid, mode)


# old name used by BASELINE.json's north star / the reference README
align_moving_img = align_img


def uniform_norm_grid(grid_shape, dim=3, device=None):
    """keymorph/utils.py:387-398 (only needed when a caller wants the identity grid itself; the
    flow-field kernels generate these coordinates in registers)."""
    axes = [torch.linspace(-1, 1, int(s), device=device) for s in grid_shape[2:2 + dim]]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).float()


def one_hot(seg):
    """keymorph/utils.py:200-205: (N,1,D,H,W) integer labels -> (N,C,D,H,W)."""
    return F.one_hot(seg)[:, 0].permute(0, 4, 1, 2, 3)


# ---- real-world coordinates (keymorph/utils.py:243-354): tiny per-keypoint affine maps --------
def convert_points_norm2voxel(points, grid_sizes):
    grid_sizes = torch.as_tensor(grid_sizes, device=points.device)
    assert grid_sizes.shape[-1] == points.shape[-1], "Dimensions don't match"
    return ((points + 1) * grid_sizes) / 2 - 0.5


def convert_points_voxel2norm(points, grid_sizes):
    grid_sizes = torch.as_tensor(grid_sizes, device=points.device)
    assert grid_sizes.shape[-1] == points.shape[-1], "Dimensions don't match"
    return (2 * (points + 0.5) / grid_sizes) - 1


def _apply_affine(points, affine):
    ones = torch.ones(points.shape[0], points.shape[1], 1, device=points.device, dtype=points.dtype)
    hom = torch.cat([points, ones], dim=2)
    return torch.bmm(affine.to(hom), hom.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :-1]


def convert_points_voxel2real(points, affine):
    return _apply_affine(points, affine)


def convert_points_real2voxel(points, affine):
    return _apply_affine(points, torch.inverse(affine))


def convert_points_norm2real(points, affine_matrices, voxel_sizes):
    return convert_points_voxel2real(convert_points_norm2voxel(points, voxel_sizes), affine_matrices)


def convert_points_real2norm(real_world_points, affine_matrices, voxel_sizes):
    return convert_points_voxel2norm(convert_points_real2voxel(real_world_points, affine_matrices),
                                     voxel_sizes)
