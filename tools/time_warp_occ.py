"""A/B timing of the fused affine / grid warp kernels at the occupancy targets KM_OPT_WARP_OCC in {4, 5, 6}.
Usage: python tools/time_warp_occ.py [S] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import _lib, ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = _lib.load()
f = O.gaussian_phantom(S, 1000).cuda()
m = O.gaussian_phantom(S, 2000).cuda()
inv = torch.inverse(O.affine_matrix_3d(0.1, 0.05, 0.3, 0.02)).cuda()[:, :3]
grid = ops.flow_field_affine(inv, (S, S, S))


def timed(fn):
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


for occ in (4, 5, 6):
    lib.km_set_option(_lib.KM_OPT_WARP_OCC, occ)
    a = timed(lambda: ops.warp_loss(m, f, mat34=inv, want_grid=True))
    g = timed(lambda: ops.warp_loss(m, f, grid=grid))
    print(f"occupancy target {occ}: affine + grid store + MSE {a:7.1f} us ({24 * S ** 3 / a / 1e3:7.1f} GB/s), "
          f"grid + MSE {g:7.1f} us ({24 * S ** 3 / g / 1e3:7.1f} GB/s)")
lib.km_set_option(_lib.KM_OPT_WARP_OCC, 4)
