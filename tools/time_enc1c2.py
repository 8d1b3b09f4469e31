"""enc1.c2 of the bench network (32 -> 64 at 128^3, two volumes; or [Cin Cout edge] from the command line): normalisation pass + plain CTA-pair kernel (N = 64
per MMA) vs the z-folded pair kernel with the GroupNorm folded in (N = 192).  python tools/time_enc1c2.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import ops  # noqa: E402

N, Cin, Cout, E = 2, 32, 64, 128
if len(sys.argv) > 3:
    Cin, Cout, E = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(3)
x = torch.relu(torch.randn(N, E, E, E, Cin, generator=g)).to("cuda", ops.act_dtype())
w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).cuda()
scale = (torch.rand(N, Cin, generator=g) + 0.5).cuda()
shift = torch.randn(N, Cin, generator=g).cuda()
wp = ops.pack_weights(w)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def unfolded():
    xn = ops.norm_apply(x, scale, shift)
    return ops.conv3d_tc_pair(xn, wp, relu=True, want_stats=True)


fl = 2.0 * 27 * Cin * Cout * E ** 3 * N
a = timed(unfolded)
b = timed(lambda: ops.conv3d_zfold_pair_gn(x, w, scale, shift, relu=True, want_stats=True))
c = timed(lambda: ops.conv3d_tc_pair(x, wp, relu=True, want_stats=True))
print(f"{Cin} -> {Cout} @ {E}^3: norm_apply + conv_tc2 (N = Cout per MMA): {a:7.1f} us ({fl / a / 1e6:6.1f} TFLOP/s); the conv alone {c:7.1f} us")
print(f"{Cin} -> {Cout} @ {E}^3: conv_zf2 with the folded GroupNorm (N = 3 Cout): {b:7.1f} us ({fl / b / 1e6:6.1f} TFLOP/s)")
