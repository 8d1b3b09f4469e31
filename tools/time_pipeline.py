"""Stage timings of one pairwise registration for any transform (CUDA events around the stages).
Usage: python tools/time_pipeline.py [S] [K] [transform ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
transforms = sys.argv[3:] or ["tps_0"]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                         num_levels=4, is_segmentation=False, conv_padding=1).eval().cuda()
model = kb.KeyMorph(net, K, 3, fused_warp=True).eval()
f = O.gaussian_phantom(S, 1000).cuda()
m = O.gaussian_phantom(S, 2000).cuda()
ms, pts = timed(lambda: model.get_keypoints(torch.cat([f, m])))
print(f"S={S} K={K}: keypoints (both volumes): {ms:.2f} ms")
pf, pm = pts[:1].contiguous(), pts[1:].contiguous()
for t in transforms:
    ms_all, _ = timed(lambda: model(f, m, transform_type=t, return_aligned_points=True))
    print(f"  forward({t}) incl. fused warp+MSE: {ms_all:.2f} ms -> {1e3 / ms_all:.1f} registrations/s")
    if t.startswith("tps"):
        lam = torch.tensor([float(t[4:])], device="cuda")
        ms_fit, (theta, _) = timed(lambda: ops.tps_fit(pf, pm, lam))
        ms_flow, grid = timed(lambda: ops.flow_field_tps(pf, theta, (S, S, S)))
        ms_warp, _ = timed(lambda: ops.warp_loss(m, f, grid=grid))
        print(f"    tps_fit {ms_fit:.2f} ms, flow_field_tps {ms_flow:.2f} ms "
              f"({K * S ** 3 / ms_flow / 1e6:.1f} G RBF evaluations/s), warp+MSE(grid) {ms_warp:.3f} ms")
    else:
        ms_warp, _ = timed(lambda: ops.warp_loss(m, f, mat34=torch.eye(4, device='cuda')[None, :3]))
        print(f"    fused affine warp+MSE {ms_warp:.3f} ms ({12.0 * S ** 3 / ms_warp / 1e6:.0f} GB/s of 12 B/voxel)")
