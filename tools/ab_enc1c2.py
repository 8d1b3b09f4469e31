"""A/B of enc1.c2 (32 -> 64 at 128^3) on the z-folded pair kernel with the folded GroupNorm vs the plain pair kernel
after a normalisation pass: 256^3 keypoint error against the fp32 CPU oracle and time of the backbone.
python tools/ab_enc1c2.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S, K = 256, 256
torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                         is_segmentation=False, conv_padding=1).eval()
sd = {k: v.clone() for k, v in net.state_dict().items()}
model = kb.KeyMorph(net.to("cuda"), K, 3, fused_warp=True).eval()
f_cpu = O.gaussian_phantom(S, 1000)
g_cpu = O.gaussian_phantom(S, 2000)
f = f_cpu.cuda()
pair = torch.cat([f, g_cpu.cuda()])
torch.set_num_threads(os.cpu_count() or 1)
ref = O.center_of_mass3d(O.unet3d_forward(sd, f_cpu, 4, 1))
for flag in (False, True, False, True, None):
    ops.USE_ZFOLD_PAIR_CIN32 = flag
    pts = model.get_keypoints(f)
    e = (pts.cpu() - ref).abs()
    for _ in range(3):
        model.get_keypoints(pair)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        model.get_keypoints(pair)
    e1.record()
    torch.cuda.synchronize()
    print(f"enc1.c2 on the z-folded pair kernel = {flag}: keypoint error max {e.max():.3e} mean {e.mean():.3e}   "
          f"keypoints of 2 volumes {e0.elapsed_time(e1) / 20:.3f} ms")
