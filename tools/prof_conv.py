"""Run one convolution layer a few times (for ncu captures of conv_tc_kernel).
Usage: python tools/prof_conv.py Cin Cout S [taps] [iters] [N] [com]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import ops  # noqa: E402

cin, cout, s = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
taps = int(sys.argv[4]) if len(sys.argv) > 4 else 27
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
n = int(sys.argv[6]) if len(sys.argv) > 6 else 1
com = len(sys.argv) > 7 and sys.argv[7] == "com"
x = torch.randn(n, s, s, s, cin, device="cuda").to(ops.act_dtype())
wp = (torch.randn(taps, cout, cin, device="cuda") / (taps * cin) ** 0.5).to(ops.act_dtype())
bias = torch.zeros(cout, device="cuda")
for _ in range(iters):
    if com:
        ops.conv3d_tc(x, wp, bias=bias, want_com=True, store=False)
    else:
        ops.conv3d_tc(x, wp, relu=True, want_stats=True)
torch.cuda.synchronize()
print("done")
