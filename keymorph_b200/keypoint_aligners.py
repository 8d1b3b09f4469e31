"""Closed-form keypoint aligners with the constructor / attribute / method surface of
keymorph/keypoint_aligners.py:14-465, computed by km_fit_affine, km_fit_rigid, km_tps_fit and the
flow-field kernels.  All keypoints are (batch, num_points, 3) in 'ij' (z, y, x) order."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .transformations import (AffineTransform, norm2voxel_matrix, raise_if_singular,
                              voxel2norm_matrix)
from .utils import convert_points_norm2real, convert_points_real2norm


def _check_real_world(points_m, points_f, aff_m, aff_f, shape_m, shape_f):
    assert aff_f is not None, "Need to provide aff_f for real-world coords"
    assert aff_m is not None, "Need to provide aff_m for real-world coords"
    assert shape_f is not None, "Need to provide shape_f for real-world coords"
    assert shape_m is not None, "Need to provide shape_m for real-world coords"
    assert points_f.shape[0] == 1, "Batch size must be 1 for real-world coords"
    assert points_m.shape[0] == 1, "Batch size must be 1 for real-world coords"


class AffineKeypointAligner(AffineTransform):
    """keymorph/keypoint_aligners.py:14-147.  The fit maps FIXED -> MOVING points (that matrix is
    `inverse_transform_matrix`, the one grid_sample needs); `transform_matrix` is its inverse."""

    _fit_op = staticmethod(ops.fit_affine)
    _name = "AffineKeypointAligner"

    def __init__(self, points_m, points_f, w=None, dim=3, align_in_real_world_coords=False,
                 aff_m=None, aff_f=None, shape_m=None, shape_f=None):
        nn.Module.__init__(self)
        if dim != 3:
            raise NotImplementedError("keymorph_b200 implements the 3-D path only")
        self.dim = dim
        self.align_in_real_world_coords = align_in_real_world_coords
        self.points_f, self.points_m = points_f, points_m
        self.shape_f, self.shape_m = shape_f, shape_m
        if align_in_real_world_coords:
            _check_real_world(points_m, points_f, aff_m, aff_f, shape_m, shape_f)
            self.aff_f, self.aff_m = aff_f, aff_m
            self.points_m = convert_points_norm2real(self.points_m, aff_m, shape_m)
            self.points_f = convert_points_norm2real(self.points_f, aff_f, shape_f)
        inv, fwd, status = self._fit_op(self.points_f, self.points_m, w)
        raise_if_singular(status, self._name)
        self.inverse_transform_matrix = inv.to(points_m.dtype)
        self.transform_matrix = fwd.to(points_m.dtype)

    def fit(self, x, y, w=None):
        """A (N,3,4) minimising sum_i w_i |y_i - A [x_i;1]|^2 (keymorph/keypoint_aligners.py:76-114)."""
        A, _, status = self._fit_op(x, y, w)
        raise_if_singular(status, self._name)
        return A[:, :3, :]

    def _grid_matrix(self):
        m = self.inverse_transform_matrix
        if self.align_in_real_world_coords:
            # p_m = voxel2norm_m . aff_m^-1 . A^-1 . aff_f . norm2voxel_f  (:132-147), all affine
            dev = m.device
            pre = torch.bmm(self.aff_f.to(m), norm2voxel_matrix(self.shape_f, dev))
            post = torch.bmm(voxel2norm_matrix(self.shape_m, dev), torch.inverse(self.aff_m.to(m)))
            m = torch.bmm(post, torch.bmm(m, pre))
        return m[:, :3, :]

    def get_forward_transformed_points(self, points):
        """keymorph/keypoint_aligners.py:116-130."""
        if self.align_in_real_world_coords:
            points = convert_points_norm2real(points, self.aff_m, self.shape_m)
        points = super().get_forward_transformed_points(points)
        if self.align_in_real_world_coords:
            points = convert_points_real2norm(points, self.aff_f, self.shape_f)
        return points

    def get_inverse_transformed_points(self, points):
        """keymorph/keypoint_aligners.py:132-147."""
        if self.align_in_real_world_coords:
            points = convert_points_norm2real(points, self.aff_f, self.shape_f)
        points = super().get_inverse_transformed_points(points)
        if self.align_in_real_world_coords:
            points = convert_points_real2norm(points, self.aff_m, self.shape_m)
        return points


class RigidKeypointAligner(AffineKeypointAligner):
    """keymorph/keypoint_aligners.py:150-213 (Arun et al.: SVD of the cross-covariance)."""

    _fit_op = staticmethod(ops.fit_rigid)
    _name = "RigidKeypointAligner"


class TPS(nn.Module):
    """keymorph/keypoint_aligners.py:216-465.  The reference re-solves the (K+4)^2 system on the
    host 2 + num_subgrids times per registration; here each direction is fitted once on the device
    and cached, and the dense evaluation streams voxels through one kernel, so `num_subgrids`,
    `compute_on_subgrids` and `use_checkpoint` are accepted and ignored."""

    def __init__(self, points_m, points_f, lmbda, w=None, dim=3, num_subgrids=4,
                 use_checkpoint=False, align_in_real_world_coords=False, aff_m=None, aff_f=None,
                 shape_m=None, shape_f=None, fit_forward=False, fit_inverse=True):
        super().__init__()
        if dim != 3:
            raise NotImplementedError("keymorph_b200 implements the 3-D path only")
        self.dim = dim
        self.num_subgrids = num_subgrids
        self.use_checkpoint = use_checkpoint
        self.lmbda = lmbda
        self.weights = w
        self.align_in_real_world_coords = align_in_real_world_coords
        self.points_f, self.points_m = points_f, points_m
        self.shape_f, self.shape_m = shape_f, shape_m
        if align_in_real_world_coords:
            _check_real_world(points_m, points_f, aff_m, aff_f, shape_m, shape_f)
            self.aff_f, self.aff_m = aff_f, aff_m
            self.points_m = convert_points_norm2real(self.points_m, aff_m, shape_m)
            self.points_f = convert_points_norm2real(self.points_f, aff_f, shape_f)
        # note the flipped order: theta maps FIXED -> MOVING (keymorph/keypoint_aligners.py:268-274)
        self.inverse_theta = self.theta = None
        if not fit_inverse:
            # groupwise iterations only move points forward (keymorph/model.py:331-394): the inverse fit the
            # reference's constructor always performs is deferred until a flow field is asked for
            if fit_forward:
                self.theta = self.fit(self.points_m, self.points_f, lmbda, weights=w)
        elif fit_forward:
            # both directions (the reference fits the forward one lazily in
            # get_forward_transformed_points, :451-465) as ONE batched device solve
            nb = self.points_f.shape[0]
            lam = torch.as_tensor(lmbda).reshape(-1)
            both = self.fit(torch.cat([self.points_f, self.points_m]), torch.cat([self.points_m, self.points_f]),
                            torch.cat([lam, lam]) if lam.numel() == nb else lam,
                            weights=None if w is None else torch.cat([w, w]))
            self.inverse_theta, self.theta = both[:nb], both[nb:]
        else:
            self.inverse_theta = self.fit(self.points_f, self.points_m, lmbda, weights=w)

    def _inverse(self):
        if self.inverse_theta is None:
            self.inverse_theta = self.fit(self.points_f, self.points_m, self.lmbda, weights=self.weights)
        return self.inverse_theta

    def fit(self, c_src, c_dst, lmbda, weights=None):
        """keymorph/keypoint_aligners.py:341-363: theta (bs, T+4, 3)."""
        theta, status = ops.tps_fit(c_src, c_dst, torch.as_tensor(lmbda), weights)
        raise_if_singular(status, "TPS")
        return theta

    @staticmethod
    def d(a, b):
        """keymorph/keypoint_aligners.py:322-334 (small inputs only; the kernels never build it)."""
        return torch.sqrt(torch.square(a[:, :, None, :] - b[:, None, :, :]).sum(-1) + 1e-6)

    @staticmethod
    def u(r):
        """keymorph/keypoint_aligners.py:336-339."""
        return r ** 2 * torch.log(r + 1e-6)

    def transform_points(self, theta, ctrl, points):
        """keymorph/keypoint_aligners.py:399-433."""
        return ops.points_transform_tps(ctrl, theta, points)

    def get_flow_field(self, grid_shape, compute_on_subgrids=False):
        """keymorph/keypoint_aligners.py:365-397: (N,D,H,W,3) in (x,y,z) order."""
        shape = tuple(int(s) for s in tuple(grid_shape)[2:])
        if not self.align_in_real_world_coords:
            return ops.flow_field_tps(self.points_f, self._inverse(), shape)
        # real-world variant (:441-448): fixed voxel -> real -> TPS -> moving voxel -> normalised
        dev = self.points_f.device
        pre = torch.bmm(self.aff_f.float().to(dev), norm2voxel_matrix(self.shape_f, dev))
        real = ops.flow_field_affine(pre[:, :3, :], shape).flip(-1).reshape(1, -1, 3)
        moved = ops.points_transform_tps(self.points_f, self._inverse(), real)
        post = torch.bmm(voxel2norm_matrix(self.shape_m, dev),
                         torch.inverse(self.aff_m.float().to(dev)))
        out = ops.points_transform_affine(post[:, :3, :], moved)
        return out.reshape(1, *shape, 3).flip(-1)

    def get_inverse_transformed_points(self, points):
        """keymorph/keypoint_aligners.py:435-449."""
        if self.align_in_real_world_coords:
            points = convert_points_norm2real(points, self.aff_f, self.shape_f)
        points = self.transform_points(self._inverse(), self.points_f, points)
        if self.align_in_real_world_coords:
            points = convert_points_real2norm(points, self.aff_m, self.shape_m)
        return points

    def get_forward_transformed_points(self, points):
        """keymorph/keypoint_aligners.py:451-465 (second, forward fit: MOVING -> FIXED)."""
        if self.theta is None:
            self.theta = self.fit(self.points_m, self.points_f, self.lmbda, weights=self.weights)
        if self.align_in_real_world_coords:
            points = convert_points_norm2real(points, self.aff_m, self.shape_m)
        points = self.transform_points(self.theta, self.points_m, points)
        if self.align_in_real_world_coords:
            points = convert_points_real2norm(points, self.aff_f, self.shape_f)
        return points


def grid_from_points(points_m, points_f, grid_shape, lmbda=None, weights=None,
                     compute_on_subgrids=True, transform="affine", **kwargs):
    """Name from the older reference revision quoted by the README / BASELINE north star
    (keymorph/model.py:516 still calls it): flow field straight from two keypoint sets."""
    if lmbda is not None or transform == "tps":
        lam = torch.as_tensor(0.0 if lmbda is None else lmbda)
        return TPS(points_m, points_f, lam, w=weights).get_flow_field(grid_shape)
    cls = RigidKeypointAligner if transform == "rigid" else AffineKeypointAligner
    return cls(points_m, points_f, w=weights).get_flow_field(grid_shape)
