"""Metric set and on-disk layout of the reference's evaluation scripts, for a registration result of this
engine (scripts/pairwise_register_eval.py:303-461, scripts/groupwise_register_eval.py:395-431,
scripts/script_utils.py:118-120) -- SURVEY.md 8f-2.  The CLI / dataset plumbing of those scripts is out of
scope; these are the two pieces a driver needs so that downstream tooling finds the same files with the same
names, shapes and dtypes."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import loss_ops, ops
from .utils import align_img


def save_dict_as_json(d, save_path):
    """scripts/script_utils.py:118-120."""
    with open(save_path, "w") as outfile:
        json.dump(d, outfile, sort_keys=True, indent=4)


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _label_map(seg):
    """np.argmax(seg, axis=1) of the scripts (int64); for a CUDA tensor the argmax runs on the device, so
    1/C of the one-hot volume crosses PCIe."""
    if isinstance(seg, torch.Tensor) and seg.is_cuda:
        return ops.argmax_channels(seg).cpu().numpy().astype(np.int64)
    return np.argmax(_np(seg), axis=1)


def pair_metrics(list_of_eval_metrics, img_f, img_a, seg_f=None, seg_a=None, grid=None, res_dict=None):
    """The metrics dictionary of scripts/pairwise_register_eval.py:303-345 for one aligned pair.
    One deliberate difference: "jdlessthan0" is the count of non-positive Jacobian determinants; the
    reference script stores jdstd under that key (line 344), an obvious slip."""
    metrics = {}
    seg_available = seg_f is not None and seg_a is not None
    if seg_available:
        dice_total = 1 - loss_ops.DiceLoss(hard=True)(seg_a, seg_f, ign_first_ch=True).item()
        dice_roi = (1 - loss_ops.DiceLoss(hard=True, return_regions=True)(seg_a, seg_f, ign_first_ch=True)
                    .cpu().numpy()).tolist()
    for m in list_of_eval_metrics:
        if m == "mse":
            metrics["mse"] = loss_ops.MSELoss()(img_f, img_a).item()
        elif m == "softdice":
            assert seg_available
            metrics["softdiceloss"] = loss_ops.DiceLoss()(seg_a, seg_f).item()
            metrics["softdice"] = 1 - metrics["softdiceloss"]
        elif m == "harddice":
            assert seg_available
            metrics["harddice"] = dice_total
        elif m == "harddiceroi":
            assert seg_available
            metrics["harddiceroi"] = dice_roi
        elif m == "hausd":
            assert seg_available
            metrics["hausd"] = loss_ops.hausdorff_distance(seg_a, seg_f)
        elif m in ("jdstd", "jdlessthan0"):
            if grid is None:
                metrics[m] = res_dict[m]
            else:
                fn = loss_ops.jdstd if m == "jdstd" else loss_ops.jdlessthan0
                metrics[m] = float(fn(grid.permute(0, 4, 1, 2, 3)))
        else:
            raise ValueError('Invalid metric "{}"'.format(m))
    return metrics


def save_pair_outputs(save_dir, i, mod1_str, mod2_str, aug, align_type_str, metrics, img_f, img_m, img_a,
                      grid=None, seg_f=None, seg_m=None, seg_a=None, points_f=None, points_m=None, points_a=None,
                      points_weights=None):
    """File layout of scripts/pairwise_register_eval.py:368-458 for pair i (batch size 1): metrics JSON,
    images / grid / label maps / keypoints as .npy; fixed- and moving-side files are written once and
    shared by all alignment types.  Returns the list of paths written by this call."""
    save_dir = str(save_dir)
    os.makedirs(save_dir, exist_ok=True)
    written = []

    def put(name, make, once=False):
        path = os.path.join(save_dir, name)
        if once and os.path.exists(path):
            return
        np.save(path, make())
        written.append(path)

    mpath = os.path.join(save_dir, f"metrics-{aug}-{align_type_str}.json")
    save_dict_as_json(metrics, mpath)
    written.append(mpath)
    pair = f"{i}-{mod1_str}-{mod2_str}-{aug}-{align_type_str}"
    put(f"img_f_{i}-{mod1_str}.npy", lambda: _np(img_f[0]), once=True)
    put(f"img_m_{i}-{mod2_str}-{aug}.npy", lambda: _np(img_m[0]), once=True)
    put(f"img_a_{pair}.npy", lambda: _np(img_a[0]))
    if grid is not None:
        put(f"grid_{pair}.npy", lambda: _np(grid[0]))
    if seg_f is not None:
        put(f"seg_f_{i}-{mod1_str}.npy", lambda: _label_map(seg_f), once=True)
        put(f"seg_m_{i}-{mod2_str}-{aug}.npy", lambda: _label_map(seg_m), once=True)
        put(f"seg_a_{pair}.npy", lambda: _label_map(seg_a))
    if points_f is not None:
        put(f"points_f_{i}-{mod1_str}.npy", lambda: _np(points_f[0]), once=True)
        put(f"points_m_{i}-{mod2_str}-{aug}.npy", lambda: _np(points_m[0]), once=True)
        put(f"points_a_{pair}.npy", lambda: _np(points_a[0]))
        if points_weights is not None:
            put(f"points_weights_{pair}.npy", lambda: _np(points_weights[0]))
    return written


def save_group_aligned(img_dir, align_type_str, grids, imgs_m, seg_dir=None, segs_m=None):
    """scripts/groupwise_register_eval.py:407-431: warp every group member with its grid and write
    img_a_{align}_{i:03}.npy (and seg_a_...).  Returns (img paths, seg paths) for the pairwise group
    metrics of loss_ops (which read each file once and keep it in HBM)."""
    os.makedirs(str(img_dir), exist_ok=True)
    img_paths, seg_paths = [], []
    for i, (grid, img_m) in enumerate(zip(grids, imgs_m)):
        path = os.path.join(str(img_dir), f"img_a_{align_type_str}_{i:03}.npy")
        np.save(path, _np(align_img(grid, img_m)))
        img_paths.append(path)
        if segs_m is not None:
            os.makedirs(str(seg_dir), exist_ok=True)
            spath = os.path.join(str(seg_dir), f"seg_a_{align_type_str}_{i:03}.npy")
            np.save(spath, _np(align_img(grid, segs_m[i])))
            seg_paths.append(spath)
    return img_paths, seg_paths
