"""A/B timing of km_tps_fit: cooperative single launch (0), one-CTA LU (1), multi-launch blocked elimination (2);
mode 3 = 0 + per-phase cycle counts printed by the kernel.  Usage: python tools/time_tps_fit.py [K] [N] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import _lib, ops  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lib = _lib.load()
g = torch.Generator().manual_seed(5)
src = (torch.rand(N, K, 3, generator=g) * 1.2 - 0.6).cuda()
dst = (src + 0.05 * torch.randn(N, K, 3, generator=g).cuda()).contiguous()
lam = torch.zeros(N).cuda()
for mode in (0, 2, 3):
    lib.km_set_option(_lib.KM_OPT_TPS_SINGLE_CTA, mode)
    for _ in range(3 if mode != 3 else 1):
        ops.tps_fit(src, dst, lam)
    torch.cuda.synchronize()
    if mode == 3:
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.tps_fit(src, dst, lam)
    e1.record()
    torch.cuda.synchronize()
    print(f"km_tps_fit mode {mode} (K={K}, N={N}): {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
lib.km_set_option(_lib.KM_OPT_TPS_SINGLE_CTA, 0)
