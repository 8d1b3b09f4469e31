"""256^3 keypoint error vs the fp32 CPU oracle for engine variants (operand type x GroupNorm fold on / off), next to
torch's own autocast drift in both 16-bit types on the same network and volume, plus the speed of each variant.
python tools/accuracy_256.py [K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = 256
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
for dt in (torch.float16, torch.bfloat16):
    with torch.no_grad(), torch.autocast("cuda", dtype=dt):
        ac = O.center_of_mass3d(O.unet3d_forward({k: v.cuda() for k, v in sd.items()}, f, 4, 1).float()).cpu()
    print(f"torch {dt} autocast drift: max {(ac - ref).abs().max():.3e} mean {(ac - ref).abs().mean():.3e}")
for name in ("fp16", "bf16"):
    kb.set_operand_dtype(name)
    for fold, stem in ((True, False), (True, True), (False, False)):
        ops.USE_GN_FOLD, ops.USE_GN_FOLD_STEM = fold, stem
        pts = model.get_keypoints(f)
        e = (pts.cpu() - ref).abs()
        for _ in range(2):
            model.get_keypoints(pair)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model.get_keypoints(pair)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} operands, fold={fold} stem_fold={stem}: max {e.max():.3e} mean {e.mean():.3e}   "
              f"keypoints of 2 volumes {e0.elapsed_time(e1) / 10:.2f} ms")
kb.set_operand_dtype("fp16")
ops.USE_GN_FOLD, ops.USE_GN_FOLD_STEM = True, None
