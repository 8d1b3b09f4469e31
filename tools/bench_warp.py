"""Time the warp kernels alone at 256^3 (CUDA events, L2 flushed by the >126 MB working set).
Usage: python tools/bench_warp.py [S] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda"
f = O.gaussian_phantom(S, 1000).to(dev)
m = O.gaussian_phantom(S, 2000).to(dev)
inv = torch.inverse(O.affine_matrix_3d(0.1, 0.05, 0.3, 0.02)).to(dev)[:, :3]
grid = ops.flow_field_affine(inv, (S, S, S))
seg = torch.rand(1, 14, S, S, S, device=dev) if S <= 256 else None


def timed(fn, nbytes, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:58s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic)")


N = S ** 3
timed(lambda: ops.warp_loss(m, f, mat34=inv, want_grid=True), 24 * N, "warp_loss affine + grid store + MSE (24 B/vox)")
timed(lambda: ops.warp_loss(m, f, mat34=inv), 12 * N, "warp_loss affine + MSE (12 B/vox)")
timed(lambda: ops.warp_loss(m, f, mat34=inv, store=False), 8 * N, "warp_loss affine, MSE only (8 B/vox)")
timed(lambda: ops.warp_loss(m, f, grid=grid), 24 * N, "warp_loss grid + MSE (24 B/vox)")
timed(lambda: ops.warp_loss(m, None, grid=grid, mode="nearest"), 20 * N, "warp_loss grid nearest (20 B/vox)")
timed(lambda: ops.grid_sample3d(m, grid), 20 * N, "grid_sample (align_img) bilinear (20 B/vox)")
timed(lambda: ops.flow_field_affine(inv, (S, S, S)), 12 * N, "flow_field_affine (12 B/vox)")
if seg is not None:
    timed(lambda: ops.warp_loss(seg, seg, grid=grid), (12 + 14 * 12) * N, "warp_loss grid C=14 + Dice sums (180 B/vox)")
    timed(lambda: ops.grid_sample3d(seg, grid), (12 + 14 * 8) * N, "grid_sample C=14 (124 B/vox)")
if seg is not None:
    lab_m = torch.randint(0, 14, (1, S, S, S), device=dev, dtype=torch.uint8)
    lab_f = torch.randint(0, 14, (1, S, S, S), device=dev, dtype=torch.uint8)
    timed(lambda: ops.warp_labels_dice(lab_m, lab_f, 14, grid=grid, want_labels=True), 15 * N,
          "label-map warp + soft & hard Dice C=14 (15 B/vox)")
    timed(lambda: ops.pair_stats(seg, seg, hard=True), 14 * 8 * N, "pair_stats hard Dice C=14 (one-hot path, 112 B/vox)")
