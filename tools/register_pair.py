"""Pairwise registration of two NIfTI volumes on the B200 engine -- the core of the reference's
scripts/register.py (:40-118 loading, :212-275 model, :278-330 registration) without torchio / nibabel.

    python tools/register_pair.py --moving m.nii.gz --fixed f.nii.gz [--moving_seg ms.nii.gz --fixed_seg fs.nii.gz]
        [--size 128] [--num_keypoints 128] [--backbone truncatedunet|unet|conv] [--load_path ckpt.pth.tar]
        [--list_of_aligns rigid affine tps_1] [--save_dir out]

Writes <save_dir>/<align>/{grid,img_a,points_f,points_m,points_a}.npy (+ labels_a.npy) and prints the metrics
the reference's eval loop reports (mse, softdice, harddice, jdstd, jdlessthan0).  Without --load_path the
backbone has the scripts' default seed-23 initialisation (no pretrained weights ship with the reference).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import hostio  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--moving", required=True)
    ap.add_argument("--fixed", required=True)
    ap.add_argument("--moving_seg")
    ap.add_argument("--fixed_seg")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--num_keypoints", type=int, default=128)
    ap.add_argument("--backbone", default="truncatedunet", choices=["truncatedunet", "unet", "conv"])
    ap.add_argument("--num_levels_for_unet", type=int, default=4)
    ap.add_argument("--num_truncated_layers_for_truncatedunet", type=int, default=1)
    ap.add_argument("--load_path")
    ap.add_argument("--list_of_aligns", nargs="+", default=["affine"])
    ap.add_argument("--save_dir")
    ap.add_argument("--seed", type=int, default=23)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(a.seed)
    K = a.num_keypoints
    if a.backbone == "conv":
        net = kb.ConvNet(3, 1, K, norm_type="instance")
    elif a.backbone == "unet":
        net = kb.UNet3D(1, K, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                        num_levels=a.num_levels_for_unet, is_segmentation=False, conv_padding=1)
    else:
        net = kb.TruncatedUNet3D(1, K, a.num_truncated_layers_for_truncatedunet, final_sigmoid=False, f_maps=32,
                                 layer_order="gcr", num_groups=8, num_levels=a.num_levels_for_unet,
                                 is_segmentation=False, conv_padding=1)
    model = kb.KeyMorph(torch.nn.DataParallel(net.to(dev)), K, 3, fused_warp=True).eval()
    if a.load_path:
        state = torch.load(a.load_path, map_location=dev, weights_only=False)["state_dict"]      # scripts/script_utils.py:59-81
        model.backbone.load_state_dict(state)
    img_f, _ = hostio.load_volume(a.fixed, size=a.size)
    img_m, _ = hostio.load_volume(a.moving, size=a.size)
    kw = {}
    if a.fixed_seg and a.moving_seg:
        lab_f, _ = hostio.load_volume(a.fixed_seg, size=a.size, labels=True)
        lab_m, _ = hostio.load_volume(a.moving_seg, size=a.size, labels=True)
        kw = dict(labels_f=lab_f.to(dev), labels_m=lab_m.to(dev), num_classes=int(max(lab_f.max(), lab_m.max())) + 1)
    res = model(img_f.to(dev), img_m.to(dev), transform_type=a.list_of_aligns, return_aligned_points=True, **kw)
    for t, r in res.items():
        jd = kb.ops.jacobian_stats(r["grid"].permute(0, 4, 1, 2, 3))[0]
        line = f"{t}: mse {r['mse'].item():.6f}  jdstd {jd[0].item():.6f}  jdlessthan0 {int(jd[1].item())}"
        if "softdice" in r:
            line += f"  softdice {1 - r['softdice'].item():.4f}  harddice {1 - r['harddice'].item():.4f}"
        print(line)
        if a.save_dir:
            d = os.path.join(a.save_dir, t)
            os.makedirs(d, exist_ok=True)
            for k in ("grid", "img_a", "points_f", "points_m", "points_a", "labels_a"):
                if k in r:
                    np.save(os.path.join(d, k + ".npy"), r[k].cpu().numpy())


if __name__ == "__main__":
    main()
