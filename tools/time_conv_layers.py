"""Per-layer timing of the tcgen05 convolution on the bench network's layer shapes, with an
optional km_set_option override.  Usage: python tools/time_conv_layers.py [key=value ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keymorph_b200 import _lib, ops  # noqa: E402

for kv in sys.argv[1:]:
    k, v = kv.split("=")
    rc = _lib.load().km_set_option(int(k), int(v))
    assert rc == 0
S, N = 256, 2
LAYERS = [("enc0.c2", 16, 32, S), ("enc1.c1", 32, 32, S // 2), ("enc1.c2", 32, 64, S // 2),
          ("enc2.c1", 64, 64, S // 4), ("enc2.c2", 64, 128, S // 4), ("enc3.c1", 128, 128, S // 8),
          ("enc3.c2", 128, 256, S // 8), ("dec0.c1", 384, 128, S // 4), ("dec0.c2", 128, 128, S // 4),
          ("dec1.c1", 192, 64, S // 2), ("dec1.c2", 64, 64, S // 2)]
tot = 0.0
for name, cin, cout, e in LAYERS:
    x = torch.randn(N, e, e, e, cin, device="cuda").to(ops.act_dtype())
    wp = (torch.randn(27, cout, cin, device="cuda") / (27 * cin) ** 0.5).to(ops.act_dtype())
    for _ in range(2):
        ops.conv3d_tc(x, wp, relu=True, want_stats=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        ops.conv3d_tc(x, wp, relu=True, want_stats=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * 27 * cin * cout * e ** 3 * N
    tot += us
    print(f"{name:8s} {cin:4d}->{cout:4d} @{e:3d}^3 x{N}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
    del x
print(f"total {tot:.1f} us")
# the z-folded kernel on the first tensor-core layer
x = torch.randn(N, S, S, S, 16, device="cuda").to(ops.act_dtype())
wz = ops.pack_weights_zfold(torch.randn(32, 16, 3, 3, 3, device="cuda") / (27 * 16) ** 0.5)
for _ in range(2):
    ops.conv3d_zfold(x, wz, relu=True, want_stats=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv3d_zfold(x, wz, relu=True, want_stats=True)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 5 * 1e3
print(f"enc0.c2 z-folded: {us:8.1f} us  {2.0 * 27 * 16 * 32 * S ** 3 * N / us / 1e6:7.1f} TFLOP/s")
# the 2-CTA (cta_group::2) kernel on the layers it supports
for name, cin, cout, e in LAYERS:
    if not ops.pair_supported(cin, cout, e, e, e):
        continue
    x = torch.randn(N, e, e, e, cin, device="cuda").to(ops.act_dtype())
    wp = (torch.randn(27, cout, cin, device="cuda") / (27 * cin) ** 0.5).to(ops.act_dtype())
    for _ in range(2):
        ops.conv3d_tc_pair(x, wp, relu=True, want_stats=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv3d_tc_pair(x, wp, relu=True, want_stats=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    print(f"{name:8s} {cin:4d}->{cout:4d} @{e:3d}^3 x{N} 2-CTA: {us:8.1f} us  {2.0 * 27 * cin * cout * e ** 3 * N / us / 1e6:7.1f} TFLOP/s", flush=True)
    del x
# the z-folded 2-CTA kernel on the Cout = 64 layers
for name, cin, cout, e in LAYERS:
    if not ops.zfold_pair_supported(cin, cout, e, e, e):
        continue
    x = torch.randn(N, e, e, e, cin, device="cuda").to(ops.act_dtype())
    wz = ops.pack_weights_zfold_pair(torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5)
    for _ in range(2):
        ops.conv3d_zfold_pair(x, wz, relu=True, want_stats=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv3d_zfold_pair(x, wz, relu=True, want_stats=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    print(f"{name:8s} {cin:4d}->{cout:4d} @{e:3d}^3 x{N} z-fold + 2-CTA: {us:8.1f} us  {2.0 * 27 * cin * cout * e ** 3 * N / us / 1e6:7.1f} TFLOP/s", flush=True)
    del x
# first layer through the 2-CTA z-folded kernel (Cin = 16), with and without the fused pool
x = torch.randn(N, S, S, S, 16, device="cuda").to(ops.act_dtype())
wz2 = ops.pack_weights_zfold_pair(torch.randn(32, 16, 3, 3, 3, device="cuda") / (27 * 16) ** 0.5)
wz1 = ops.pack_weights_zfold(torch.randn(32, 16, 3, 3, 3, device="cuda") / (27 * 16) ** 0.5)
for label, fn in (("zfold  1-CTA, full store", lambda: ops.conv3d_zfold(x, wz1, relu=True, want_stats=True)),
                  ("zfold  1-CTA, pooled only", lambda: ops.conv3d_zfold(x, wz1, relu=True, want_stats=True, pool=True, store=False)),
                  ("zfold  2-CTA, full store", lambda: ops.conv3d_zfold_pair(x, wz2, relu=True, want_stats=True)),
                  ("zfold  2-CTA, pooled only", lambda: ops.conv3d_zfold_pair(x, wz2, relu=True, want_stats=True, pool=True, store=False))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"enc0.c2 16->32 @256^3 x{N} {label}: {e0.elapsed_time(e1) / 5 * 1e3:8.1f} us", flush=True)
