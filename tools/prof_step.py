"""One pairwise registration step of the bench workload (for ncu launch lists / captures).
Usage: python tools/prof_step.py [S] [K] [transform] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
transform = sys.argv[3] if len(sys.argv) > 3 else "affine"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                         num_levels=4, is_segmentation=False, conv_padding=1).eval().cuda()
model = kb.KeyMorph(torch.nn.DataParallel(net), K, 3, fused_warp=True).eval()
f = O.gaussian_phantom(S, 1000).cuda()
m = O.gaussian_phantom(S, 2000).cuda()
for _ in range(steps):
    r = model(f, m, transform_type=transform, return_aligned_points=True)[transform]
    torch.cuda.synchronize()
print("mse", r["mse"].item())
