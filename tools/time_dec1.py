"""dec1.c1 of the bench network (192 -> 64 at 128^3, two volumes) three ways: in-place concat kernel (+ the
upsample it needs) vs the coarse-lattice split (conv_up2 + skip kernel with addend).  python tools/time_dec1.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import ops  # noqa: E402

N, Cs, Cu, Dc = 2, 64, 128, 64
g = torch.Generator().manual_seed(1)
dev = "cuda"
skip = torch.relu(torch.randn(N, 2 * Dc, 2 * Dc, 2 * Dc, Cs, generator=g)).to(dev, ops.act_dtype())
coarse = torch.relu(torch.randn(N, Dc, Dc, Dc, Cu, generator=g)).to(dev, ops.act_dtype())
w = (torch.randn(64, Cs + Cu, 3, 3, 3, generator=g) / (27 * (Cs + Cu)) ** 0.5).to(dev)
scale = (torch.rand(N, Cs + Cu, generator=g) + 0.5).to(dev)
shift = torch.randn(N, Cs + Cu, generator=g).to(dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


up = ops.upsample2(coarse)
part = ops.conv3d_up2_gn(coarse, w, scale, Cs)
t_up = timed(lambda: ops.upsample2(coarse))
t_cat = timed(lambda: ops.conv3d_zfold_pair_gn(skip, w, scale, shift, relu=True, want_stats=True, x1=up))
t_part = timed(lambda: ops.conv3d_up2_gn(coarse, w, scale, Cs))
t_add = timed(lambda: ops.conv3d_zfold_pair_gn_add(skip, w, scale, shift, part, relu=True, want_stats=True))
fl_up = 2.0 * 8 * Cu * 64 * (2 * Dc) ** 3 * N
fl_skip = 2.0 * 27 * Cs * 64 * (2 * Dc) ** 3 * N
fl_cat = 2.0 * 27 * (Cs + Cu) * 64 * (2 * Dc) ** 3 * N
print(f"upsample2 {t_up:.3f} ms + in-place concat conv {t_cat:.3f} ms ({fl_cat / t_cat / 1e9:.0f} TFLOP/s) = {t_up + t_cat:.3f} ms")
print(f"conv_up2 {t_part:.3f} ms ({fl_up / t_part / 1e9:.0f} TFLOP/s executed) + skip conv with addend {t_add:.3f} ms "
      f"({fl_skip / t_add / 1e9:.0f} TFLOP/s) = {t_part + t_add:.3f} ms")
