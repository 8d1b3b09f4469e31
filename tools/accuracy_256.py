"""256^3 keypoint error vs the fp32 CPU oracle for engine variants (GroupNorm fold on / off), next to torch's
own bf16-autocast drift on the same network and volume.  python tools/accuracy_256.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S, K = 256, 64
torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                         is_segmentation=False, conv_padding=1).eval()
sd = {k: v.clone() for k, v in net.state_dict().items()}
model = kb.KeyMorph(net.to("cuda"), K, 3, fused_warp=True).eval()
f_cpu = O.gaussian_phantom(S, 1000)
f = f_cpu.cuda()
torch.set_num_threads(os.cpu_count() or 1)
ref = O.center_of_mass3d(O.unet3d_forward(sd, f_cpu, 4, 1))
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    ac = O.center_of_mass3d(O.unet3d_forward({k: v.cuda() for k, v in sd.items()}, f, 4, 1).float()).cpu()
print(f"torch bf16 autocast drift: max {(ac - ref).abs().max():.3e} mean {(ac - ref).abs().mean():.3e}")
for fold, stem in ((False, False), (True, False), (True, True)):
    ops.USE_GN_FOLD, ops.USE_GN_FOLD_STEM = fold, stem
    pts = model.get_keypoints(f) if hasattr(model, "get_keypoints") else model(f, f, transform_type="affine", return_aligned_points=False)["affine"]["points_f"]
    e = (pts.cpu() - ref).abs()
    print(f"fold={fold} stem_fold={stem}: max {e.max():.3e} mean {e.mean():.3e}")
