"""MSELoss / DiceLoss of keymorph/loss_ops.py:9-63 on the deterministic two-stage reduction
kernels (km_pair_stats).  The per-(n,c) sums come back as fp64; the few remaining scalar ops are
done with torch on the device."""
from __future__ import annotations

import torch

from . import ops


class MSELoss(torch.nn.Module):
    """keymorph/loss_ops.py:9-13 (F.mse_loss, mean reduction)."""

    def forward(self, pred, target):
        assert pred.shape == target.shape
        n, c = pred.shape[0], pred.shape[1]
        sums = ops.pair_stats(pred.reshape(n, c, -1), target.reshape(n, c, -1))
        return (sums[..., 0].sum() / pred.numel()).float()


def dice_from_sums(sums, ign_first_ch=False, return_regions=False):
    """sums (N,C,4) = [., sum p*t, sum p*p, sum t*t] -> Dice loss with eps = 1
    (keymorph/loss_ops.py:35,54-63)."""
    if ign_first_ch:
        sums = sums[:, 1:]
    num = 2 * sums[..., 1] + 1
    den = sums[..., 2] + sums[..., 3] + 1
    loss = (1 - num / den).float()
    return loss.mean(0) if return_regions else loss.mean()


class DiceLoss(torch.nn.Module):
    """keymorph/loss_ops.py:16-63. hard=True one-hots argmax_c(pred) first (first maximum wins)."""

    def __init__(self, hard=False, return_regions=False):
        super().__init__()
        self.hard = hard
        self.return_regions = return_regions

    def forward(self, pred, target, ign_first_ch=False):
        assert pred.size() == target.size(), "Input and target are different dim"
        n, c = target.shape[0], target.shape[1]
        sums = ops.pair_stats(pred.reshape(n, c, -1), target.reshape(n, c, -1), hard=self.hard)
        return dice_from_sums(sums, ign_first_ch, self.return_regions)


def _as_field(disp):
    if not isinstance(disp, torch.Tensor):
        disp = torch.as_tensor(disp)
    if not disp.is_cuda:
        raise ops._lib.KMError("keymorph_b200.loss_ops.jdstd / jdlessthan0 run on CUDA tensors only")
    return disp


def jdstd(disp):
    """keymorph/loss_ops.py:237-240: std of the Jacobian determinant of a (1,3,D,H,W) field."""
    return ops.jacobian_stats(_as_field(disp))[0, 0]


def jdlessthan0(disp, as_percentage=False):
    """keymorph/loss_ops.py:243-248: number (or fraction) of interior voxels with det(J) <= 0."""
    st = ops.jacobian_stats(_as_field(disp))[0]
    return st[1] / st[3] if as_percentage else st[1]
