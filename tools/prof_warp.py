"""Launch every HBM-/issue-bound kernel of the registration path twice at full size (for ncu captures).
Usage: python tools/prof_warp.py [S] [K] [which ...]   which in {affine, grid1, grid14, gs1, gs14, tps, fit, labels}"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
which = set(sys.argv[3:]) or {"affine", "grid1", "grid14", "gs1", "gs14", "tps", "fit", "labels"}
dev = "cuda"
f = O.gaussian_phantom(S, 1000).to(dev)
m = O.gaussian_phantom(S, 2000).to(dev)
inv = torch.inverse(O.affine_matrix_3d(0.1, 0.05, 0.3, 0.02)).to(dev)[:, :3]
grid = ops.flow_field_affine(inv, (S, S, S))
g = torch.Generator().manual_seed(5)
pf = (torch.rand(1, K, 3, generator=g) * 1.2 - 0.6).to(dev)
pm = (pf + 0.05 * torch.randn(1, K, 3, generator=g).to(dev)).contiguous()
lam = torch.zeros(1, device=dev)
for _ in range(2):
    if "affine" in which:
        ops.warp_loss(m, f, mat34=inv, want_grid=True)
    if "grid1" in which:
        ops.warp_loss(m, f, grid=grid)
    if "gs1" in which:
        ops.grid_sample3d(m, grid)
    if "fit" in which or "tps" in which:
        theta, _ = ops.tps_fit(pf, pm, lam)
    if "tps" in which:
        ops.flow_field_tps(pf, theta, (S, S, S))
    torch.cuda.synchronize()
if which & {"grid14", "gs14", "labels"}:
    seg = torch.rand(1, 14, S, S, S, device=dev)
    lab_m = torch.randint(0, 14, (1, S, S, S), device=dev, dtype=torch.uint8)
    lab_f = torch.randint(0, 14, (1, S, S, S), device=dev, dtype=torch.uint8)
    for _ in range(2):
        if "grid14" in which:
            ops.warp_loss(seg, seg, grid=grid)
        if "gs14" in which:
            ops.grid_sample3d(seg, grid)
        if "labels" in which:
            ops.warp_labels_dice(lab_m, lab_f, 14, grid=grid, want_labels=True)
        torch.cuda.synchronize()
print("done")
