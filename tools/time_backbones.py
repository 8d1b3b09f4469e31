"""Keypoint-extraction time of the three backbones at the bench size (two 256^3 volumes as one batch), CUDA events.
Usage: python tools/time_backbones.py [S] [K] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
x = torch.cat([O.gaussian_phantom(S, 1000), O.gaussian_phantom(S, 2000)]).cuda()
# dense conv FLOPs per image (BASELINE.md section 3, S = 256, K = 256)
FLOPS = {"TruncatedUNet3D": 3.939e12, "UNet3D": 7.856e12, "ConvNet": 4.349e12}
for name in ("TruncatedUNet3D", "UNet3D", "ConvNet"):
    torch.manual_seed(23)
    if name == "TruncatedUNet3D":
        net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                                 is_segmentation=False, conv_padding=1)
    elif name == "UNet3D":
        net = kb.UNet3D(1, K, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                        is_segmentation=False, conv_padding=1)
    else:
        net = kb.ConvNet(3, 1, K, norm_type="instance")
    model = kb.KeyMorph(net.eval().cuda(), K, 3).eval()
    try:
        for _ in range(2):
            model.get_keypoints(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            pts = model.get_keypoints(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 2 * FLOPS[name] / (ms * 1e-3) / 1e12 if (S, K) == (256, 256) else float("nan")
        print(f"{name:16s} S={S} K={K}: keypoints of 2 volumes {ms:8.2f} ms  = {tf:7.1f} TFLOP/s of dense conv work "
              f"({kb.act_dtype()}), peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
    except Exception as e:  # noqa: BLE001
        print(f"{name}: failed: {type(e).__name__}: {str(e)[:200]}")
    del model, net
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
