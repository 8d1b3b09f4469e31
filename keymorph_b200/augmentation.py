"""Affine augmentation on the fused warp kernel (keymorph/augmentation.py:81-290, 3-D path).

The reference materialises `phi_inv = AffineTransform(matrix=Ma).get_flow_field(...)` (12 B / voxel
written + read) and then runs F.grid_sample.  Here the 3x4 inverse matrix goes straight into
km_warp_loss in KM_COORD_AFFINE mode: every voxel's sampling coordinate is computed in registers, so
an augmentation pass moves 4 B in + 4 B out per voxel and channel and nothing else.
"""
from __future__ import annotations

import torch

from . import ops
from .transformations import AffineTransform


class AffineDeformation3d:
    """keymorph/augmentation.py:81-178."""

    def __init__(self, device="cuda:0"):
        self.device = device

    def build_affine_matrix(self, batch_size, params):
        """params = (scale (b,3), offset (b,3), angles (b,3), shear (b,6)), b == 1 or batch_size;
        returns M = Mz Ms Mt (R3 R2 R1), (batch_size,4,4) fp32 (keymorph/augmentation.py:85-158).
        The sixteen products are composed on the host in fp32 -- the parameters are host tensors
        (torch.FloatTensor(...) in the convenience functions) -- and cross PCIe as 64 bytes per sample."""
        scale, offset, theta, shear = (torch.as_tensor(p, dtype=torch.float32).cpu().reshape(-1, n)
                                       for p, n in zip(params, (3, 3, 3, 6)))
        b = int(batch_size)

        def full(t):
            if t.shape[0] not in (1, b):
                raise ValueError(f"augmentation parameters have batch {t.shape[0]}, expected 1 or {b}")
            return t.expand(b, -1)

        scale, offset, theta, shear = full(scale), full(offset), full(theta), full(shear)
        eye = torch.eye(4).repeat(b, 1, 1)
        Ms = eye.clone()
        Ms[:, 0, 0], Ms[:, 1, 1], Ms[:, 2, 2] = scale[:, 0], scale[:, 1], scale[:, 2]
        Mt = eye.clone()
        Mt[:, :3, 3] = offset
        c, s = torch.cos(theta), torch.sin(theta)
        R1, R2, R3 = eye.clone(), eye.clone(), eye.clone()
        R1[:, 1, 1], R1[:, 1, 2], R1[:, 2, 1], R1[:, 2, 2] = c[:, 0], -s[:, 0], s[:, 0], c[:, 0]
        R2[:, 0, 0], R2[:, 0, 2], R2[:, 2, 0], R2[:, 2, 2] = c[:, 1], s[:, 1], -s[:, 1], c[:, 1]
        R3[:, 0, 0], R3[:, 0, 1], R3[:, 1, 0], R3[:, 1, 1] = c[:, 2], -s[:, 2], s[:, 2], c[:, 2]
        Mz = eye.clone()
        Mz[:, 0, 1], Mz[:, 0, 2] = shear[:, 0], shear[:, 1]
        Mz[:, 1, 0], Mz[:, 1, 2] = shear[:, 2], shear[:, 3]
        Mz[:, 2, 0], Mz[:, 2, 1] = shear[:, 4], shear[:, 5]
        Mr = torch.bmm(R3, torch.bmm(R2, R1))
        M = torch.bmm(Mz, torch.bmm(Ms, torch.bmm(Mt, Mr)))
        return M.to(self.device)

    def deform_img(self, img, params, interp_mode="bilinear"):
        """keymorph/augmentation.py:160-163, with the flow field kept in registers."""
        Ma = self.build_affine_matrix(len(img), params)
        t = AffineTransform(matrix=Ma)
        out, _ = ops.warp_loss(img, None, mat34=t._grid_matrix(), mode=interp_mode)
        return out

    def deform_points(self, points, params):
        """keymorph/augmentation.py:165-167."""
        Ma = self.build_affine_matrix(len(points), params)
        return AffineTransform(matrix=Ma).get_forward_transformed_points(points)

    def __call__(self, img, **kwargs):
        return self.deform_img(img, kwargs["params"], kwargs["interp_mode"])


def _need_3d(img):
    if img.dim() != 5:
        raise NotImplementedError("keymorph_b200 implements the 3-D path only: expected (bs, nch, l, w, h)")


def _apply(img, params, seg, points, return_affine_matrix=False):
    augmenter = AffineDeformation3d(device=img.device)
    Ma = augmenter.build_affine_matrix(len(img), params)
    t = AffineTransform(matrix=Ma)
    mat34 = t._grid_matrix()
    res = (ops.warp_loss(img, None, mat34=mat34, mode="bilinear")[0],)
    if seg is not None:
        res += (ops.warp_loss(seg, None, mat34=mat34, mode="nearest")[0],)
    if points is not None:
        res += (t.get_forward_transformed_points(points),)
    if return_affine_matrix:
        res += (Ma,)
    return res[0] if len(res) == 1 else res


def _sample(max_random_params, scale_params, generator=None):
    s, o, a, z = (p * scale_params for p in max_random_params)
    u = lambda n, lo, hi: torch.empty(1, n).uniform_(lo, hi, generator=generator)  # noqa: E731
    return u(3, 1 - s, 1 + s), u(3, -o, o), u(3, -a, a), u(6, -z, z)


def random_affine_augment(img, seg=None, points=None, max_random_params=(0.2, 0.2, 3.1416, 0.1),
                          scale_params=1, return_affine_matrix=False, generator=None):
    """keymorph/augmentation.py:182-221: one random (scale, offset, angles, shear) draw -- in that
    order, from the global CPU generator unless `generator` is given -- applied to the whole batch;
    `seg` is resampled with nearest-neighbour interpolation, `points` are moved by M."""
    _need_3d(img)
    return _apply(img, _sample(max_random_params, scale_params, generator), seg, points, return_affine_matrix)


def affine_augment(img, fixed_params, seg=None, points=None):
    """keymorph/augmentation.py:224-254: isotropic fixed parameters (s, o, a, z)."""
    _need_3d(img)
    s, o, a, z = fixed_params
    params = (torch.full((1, 3), 1.0 + s), torch.full((1, 3), float(o)), torch.full((1, 3), float(a)),
              torch.full((1, 6), float(z)))
    return _apply(img, params, seg, points)


def random_affine_augment_pair(img1, img2, max_random_params=(0.2, 0.2, 3.1416, 0.1), scale_params=1,
                               generator=None):
    """keymorph/augmentation.py:257-290: the same random draw applied to two images."""
    _need_3d(img1)
    params = _sample(max_random_params, scale_params, generator)
    return _apply(img1, params, None, None), _apply(img2, params, None, None)
