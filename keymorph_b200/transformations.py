"""AffineTransform of keymorph/transformations.py:7-114 on the km_* kernels."""
from __future__ import annotations

import threading

import torch
import torch.nn as nn

from . import ops

# The reference raises torch.linalg.LinAlgError when a matrix is singular.  Checking the device
# status word costs one tiny D2H copy per fit; set to False to keep the whole pair pipeline
# asynchronous (singular fits then yield NaN matrices instead of an exception).
CHECK_SINGULAR = True


# per-thread list of (status tensor, name) while a caller batches the checks: two threads driving
# KeyMorph.forward concurrently (multi-GPU driver threads, a server) must not see each other's flags
_TLS = threading.local()


def raise_if_singular(status, what):
    """LinAlgError like the reference's torch.inverse / torch.linalg.solve.  The status flag lives on
    the device: reading it synchronises, so KeyMorph.forward collects the flags of a whole call
    (`deferred_singular_checks`) and reads them once, after everything has been enqueued."""
    if not CHECK_SINGULAR:
        return
    pending = getattr(_TLS, "pending", None)
    if pending is not None:
        pending.append((status, what))
        return
    if bool(status.any().item()):
        raise torch.linalg.LinAlgError(f"{what}: the matrix is singular (status={status.tolist()})")


class deferred_singular_checks:
    """Context manager: singular-matrix flags raised inside are checked on exit with one device
    synchronisation instead of one per fit (the kernels are safe on singular input; their outputs
    are simply not returned)."""

    def __init__(self, sink=None):
        # sink: a list that receives the (status, name) pairs instead of being read here -- used while a
        # CUDA graph is being captured, where a device->host read is not allowed (KeyMorph(cuda_graph=True))
        self._sink = sink

    def __enter__(self):
        self._outer = getattr(_TLS, "pending", None)
        _TLS.pending = []
        return self

    def __exit__(self, exc_type, exc, tb):
        pending, _TLS.pending = _TLS.pending, self._outer
        if self._sink is not None:
            self._sink.extend(pending)
            return False
        if exc_type is None and pending:
            flags = torch.stack([st.reshape(-1).any() for st, _ in pending]).cpu()
            for bad, (st, what) in zip(flags.tolist(), pending):
                if bad:
                    raise torch.linalg.LinAlgError(f"{what}: the matrix is singular (status={st.tolist()})")
        return False


def norm2voxel_matrix(sizes, device):
    """4x4 matrix of keymorph/utils.py:243-258: v = (p + 1) * size / 2 - 0.5 (per axis)."""
    s = torch.as_tensor(sizes, dtype=torch.float32, device=device).reshape(-1)[-3:]
    m = torch.eye(4, device=device)
    m[:3, :3] = torch.diag(s / 2)
    m[:3, 3] = s / 2 - 0.5
    return m[None]


def voxel2norm_matrix(sizes, device):
    """4x4 matrix of keymorph/utils.py:261-276: p = 2 * (v + 0.5) / size - 1."""
    s = torch.as_tensor(sizes, dtype=torch.float32, device=device).reshape(-1)[-3:]
    m = torch.eye(4, device=device)
    m[:3, :3] = torch.diag(2 / s)
    m[:3, 3] = 1 / s - 1
    return m[None]


class AffineTransform(nn.Module):
    """Affine transformations (keymorph/transformations.py:7-30): exactly one of `matrix`
    (moving -> fixed) or `inverse_matrix` (fixed -> moving), shape (N,4,4)."""

    def __init__(self, matrix=None, inverse_matrix=None, dim=3):
        super().__init__()
        if dim != 3:
            raise NotImplementedError("keymorph_b200 implements the 3-D path only")
        self.dim = dim
        if matrix is not None and inverse_matrix is None:
            self.transform_matrix = matrix
            inv, st = ops.inverse44(matrix)
            raise_if_singular(st, "AffineTransform")
            self.inverse_transform_matrix = inv.to(matrix.dtype)
        elif matrix is None and inverse_matrix is not None:
            self.inverse_transform_matrix = inverse_matrix
            inv, st = ops.inverse44(inverse_matrix)
            raise_if_singular(st, "AffineTransform")
            self.transform_matrix = inv.to(inverse_matrix.dtype)
        else:
            raise ValueError("Only one of matrix or inverse_matrix should be provided")

    def _square(self, matrix):
        """keymorph/transformations.py:32-35."""
        sq = torch.eye(self.dim + 1, device=matrix.device, dtype=matrix.dtype).repeat(
            matrix.shape[0], 1, 1)
        sq[:, : self.dim, : self.dim + 1] = matrix
        return sq

    # the 3x4 matrix taking normalised FIXED-grid coordinates to normalised MOVING coordinates
    def _grid_matrix(self):
        return self.inverse_transform_matrix[:, :3, :]

    def affine_grid(self, grid_shape):
        """keymorph/transformations.py:37-58: moving coordinates of every fixed voxel, (z,y,x) order."""
        return self.get_flow_field(grid_shape).flip(-1)

    def get_flow_field(self, grid_shape, **kwargs):
        """keymorph/transformations.py:60-79: grid for F.grid_sample, (N,D,H,W,3) in (x,y,z) order."""
        return ops.flow_field_affine(self._grid_matrix(), tuple(grid_shape)[2:])

    def get_forward_transformed_points(self, points):
        """keymorph/transformations.py:81-96: p_f = A p_m."""
        return ops.points_transform_affine(self.transform_matrix[:, :3, :], points)

    def get_inverse_transformed_points(self, points):
        """keymorph/transformations.py:98-114: p_m = A^-1 p_f."""
        return ops.points_transform_affine(self.inverse_transform_matrix[:, :3, :], points)
