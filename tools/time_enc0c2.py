"""enc0.c2 (16 -> 32 @ 256^3, pooled in the epilogue): one-SM z-folded kernel vs the cta_group::2 z-folded kernel,
plain and with the GroupNorm folded in.  Usage: python tools/time_enc0c2.py [S] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.randn(N, S, S, S, 16, device="cuda").abs().to(ops.act_dtype())
w = torch.randn(32, 16, 3, 3, 3, device="cuda") / (27 * 16) ** 0.5
scale = torch.rand(N, 16, device="cuda") + 0.5
shift = torch.randn(N, 16, device="cuda") * 0.1
wz, wz2 = ops.pack_weights_zfold(w), ops.pack_weights_zfold_pair(w)
fl = 2.0 * 27 * 16 * 32 * S ** 3 * N


def timed(fn, label):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"{label:58s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
    return out


a = timed(lambda: ops.conv3d_zfold(x, wz, relu=True, want_stats=True, pool=True, store=False), "conv_zf  (one SM per MMA), pooled")
b = timed(lambda: ops.conv3d_zfold_pair(x, wz2, relu=True, want_stats=True, pool=True, store=False), "conv_zf2 (cta_group::2), pooled")
print("   pooled outputs equal:", torch.equal(a[1], b[1]), " max |diff|", (a[1].float() - b[1].float()).abs().max().item())
c = timed(lambda: ops.conv3d_zfold_gn(x, w, scale, shift, relu=True, want_stats=True, pool=True, store=False),
          "conv_zf  + folded GroupNorm, pooled")
d = timed(lambda: ops.conv3d_zfold_pair_gn(x, w, scale, shift, relu=True, want_stats=True, pool=True, store=False),
          "conv_zf2 + folded GroupNorm, pooled")
print("   pooled outputs equal:", torch.equal(c[1], d[1]), " max |diff|", (c[1].float() - d[1].float()).abs().max().item())
