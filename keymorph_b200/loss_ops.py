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


def hausdorff_distance(test_seg, gt_seg, sampling=(1.25, 1.25, 10.0)):
    """keymorph/loss_ops.py:144-157: mean over the batch of the symmetric Hausdorff distance between
    the surfaces of channel 0 of two (bs, n_ch, l, w, h) maps, voxel sizes (1.25, 1.25, 10) as in the
    reference.  Surface extraction, the exact Euclidean distance transforms and the maxima run on the
    device (km_hausdorff); one (N,2) fp64 read-back.  Returns a Python float like the reference's numpy
    scalar; raises where scipy's transform is undefined (a map without any foreground voxel)."""
    a, b = _as_field(test_seg), _as_field(gt_seg)
    assert a.dim() == 5 and a.shape == b.shape
    res = ops.hausdorff(a.float()[:, 0], b.float()[:, 0], sampling).cpu()
    if bool((res[:, 1] != 0).any()):
        raise ops._lib.KMError("hausdorff_distance: a segmentation map has no surface voxel in channel 0")
    return float(res[:, 0].mean())


def fast_dice(x, y):
    """keymorph/loss_ops.py:66-106: mean over the labels present in argmax_c(x) or argmax_c(y) of
    2 |x=l and y=l| / (|x=l| + |y=l| + 1e-5); 1 when a single label is present."""
    assert x.shape == y.shape, f"both inputs should have same size, had {x.shape} and {y.shape}"
    n, c = x.shape[0], x.shape[1]
    y_hard = torch.nn.functional.one_hot(ops.argmax_channels(y).long(), c).movedim(-1, 1).float()
    s = ops.pair_stats(x.reshape(n, c, -1), y_hard.reshape(n, c, -1), hard=True).sum(0).cpu()
    inter, ca, cb = s[:, 1], s[:, 2], s[:, 3]
    present = (ca + cb) > 0
    if int(present.sum()) <= 1:
        return 1.0
    return float((2 * inter[present] / (ca[present] + cb[present] + 1e-5)).mean())


def _load_file(path, device="cuda"):
    """keymorph/loss_ops.py:405-411, onto the device (fp32) without nibabel."""
    import numpy as np
    if path.endswith(".nii") or path.endswith(".nii.gz"):
        from .hostio import read_nifti
        arr = read_nifti(path)[0]
    elif path.endswith(".npy"):
        arr = np.load(path)
    else:
        raise ValueError("File format not supported")
    return torch.as_tensor(np.ascontiguousarray(arr), dtype=torch.float32).to(device)


class _Group:
    """Members of a group, given as one (G, ...) tensor or a list of file paths.  Every file is read
    and uploaded ONCE and stays resident for all O(G^2) pairs (the reference re-reads both files of
    every pair from disk, loss_ops.py:424-431): a group of 256^3 x 14-channel maps is 0.9 GB per
    subject, so a few dozen subjects fit the 180 GB of HBM."""

    def __init__(self, batch, device="cuda"):
        self.batch, self.device, self.cache = batch, device, {}
        self.paths = len(batch) > 0 and isinstance(batch[0], str)

    def __len__(self):
        return len(self.batch)

    def __getitem__(self, i):
        if not self.paths:
            return self.batch[i:i + 1]
        if i not in self.cache:
            self.cache[i] = _load_file(self.batch[i], self.device)
        return self.cache[i]


class _AvgPairwiseLoss(torch.nn.Module):
    """keymorph/loss_ops.py:414-435: mean of metric_fn over all unordered pairs of the group."""

    def __init__(self, metric_fn):
        super().__init__()
        self.metric_fn = metric_fn

    def forward(self, batch_of_imgs):
        group = _Group(batch_of_imgs)
        loss, num = 0, 0
        for i in range(len(group)):
            for j in range(i + 1, len(group)):
                loss = loss + self.metric_fn(group[i], group[j])
                num += 1
        return loss / num


class MSEPairwiseLoss(_AvgPairwiseLoss):
    def __init__(self):
        super().__init__(MSELoss().forward)


class SoftDicePairwiseLoss(_AvgPairwiseLoss):
    def __init__(self):
        super().__init__(DiceLoss().forward)


class HardDicePairwiseLoss(_AvgPairwiseLoss):
    def __init__(self):
        super().__init__(DiceLoss(hard=True).forward)


class HausdorffPairwiseLoss(_AvgPairwiseLoss):
    def __init__(self):
        super().__init__(hausdorff_distance)


class _AvgGridMetric(torch.nn.Module):
    """keymorph/loss_ops.py:466-482: mean of a grid metric over (G,D,H,W,3) grids or grid files."""

    def __init__(self, metric_fn):
        super().__init__()
        self.metric_fn = metric_fn

    def forward(self, batch_of_grids):
        group = _Group(batch_of_grids)
        tot = 0
        for i in range(len(group)):
            tot = tot + self.metric_fn(group[i].permute(0, 4, 1, 2, 3))   # strided view, read in place
        return tot / len(group)


class AvgJDStd(_AvgGridMetric):
    def __init__(self):
        super().__init__(jdstd)


class AvgJDLessThan0(_AvgGridMetric):
    def __init__(self):
        super().__init__(jdlessthan0)


class MultipleAvgSegPairwiseMetric(torch.nn.Module):
    """keymorph/loss_ops.py:499-527: several pairwise segmentation metrics in one sweep over the pairs."""

    def __init__(self):
        super().__init__()
        self.name2fn = {
            "dice": fast_dice,
            "harddice": DiceLoss(hard=True).forward,
            "harddiceroi": DiceLoss(hard=True, return_regions=True).forward,
            "softdice": DiceLoss().forward,
            "hausd": hausdorff_distance,
        }

    def forward(self, batch_of_imgs, fn_names):
        group = _Group(batch_of_imgs)
        res = {name: 0 for name in fn_names}
        num = 0
        for i in range(len(group)):
            for j in range(i + 1, len(group)):
                for name in fn_names:
                    res[name] = res[name] + self.name2fn[name](group[i], group[j])
                num += 1
        return {name: res[name] / num for name in fn_names}


class MultipleAvgGridMetric(torch.nn.Module):
    """keymorph/loss_ops.py:530-551."""

    def __init__(self):
        super().__init__()
        self.name2fn = {"jdstd": jdstd, "jdlessthan0": jdlessthan0}

    def forward(self, batch_of_grids, fn_names):
        group = _Group(batch_of_grids)
        res = {name: 0 for name in fn_names}
        for i in range(len(group)):
            g = group[i].permute(0, 4, 1, 2, 3)
            for name in fn_names:
                res[name] = res[name] + self.name2fn[name](g)
        return {name: res[name] / len(group) for name in fn_names}
