"""A/B timing of the dense TPS field kernel (KM_OPT_TPS_PACKED x KM_OPT_TPS_VPT) at full size.
Usage: python tools/time_tps_field.py [S] [K] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import _lib, ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = _lib.load()
g = torch.Generator().manual_seed(5)
pf = (torch.rand(1, K, 3, generator=g) * 1.2 - 0.6).cuda()
pm = (pf + 0.05 * torch.randn(1, K, 3, generator=g).cuda()).contiguous()
theta, _ = ops.tps_fit(pf, pm, torch.zeros(1).cuda())
ref = None
for fast in (1, 0):
    lib.km_set_option(_lib.KM_OPT_TPS_FAST, fast)
    for packed in (1, 0):
        for vpt in (8, 4, 2):
            lib.km_set_option(_lib.KM_OPT_TPS_PACKED, packed)
            lib.km_set_option(_lib.KM_OPT_TPS_VPT, vpt)
            for _ in range(2):
                grid = ops.flow_field_tps(pf, theta, (S, S, S))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                grid = ops.flow_field_tps(pf, theta, (S, S, S))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if ref is None:
                ref = grid.clone()
            err = (grid - ref).abs().max().item()
            print(f"fast={fast} packed={packed} vpt={vpt}: {ms:7.3f} ms  {K * S ** 3 / ms / 1e6:8.1f} G terms/s  "
                  f"max|grid - first variant| {err:.2e}")
lib.km_set_option(_lib.KM_OPT_TPS_FAST, 1)
lib.km_set_option(_lib.KM_OPT_TPS_PACKED, 1)
lib.km_set_option(_lib.KM_OPT_TPS_VPT, 0)
