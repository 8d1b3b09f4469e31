"""Where does the end-to-end gap come from?  Times the bench step in four ways (see labels)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200.hostio import prefetch_to_device  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S, K, steps = 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                         num_levels=4, is_segmentation=False, conv_padding=1).eval().to(dev)
model = kb.KeyMorph(torch.nn.DataParallel(net), K, 3, fused_warp=True).eval()
fh = O.gaussian_phantom(S, 1000).pin_memory()
mh = O.gaussian_phantom(S, 2000).pin_memory()
f, m = fh.to(dev), mh.to(dev)


def step(a, b):
    return model(a, b, transform_type="affine", return_aligned_points=True)["affine"]["mse"]


def run(label, fn):
    fn(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(steps)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"{label:60s} {ms:7.2f} ms/step  {1e3 / ms:6.1f} reg/s", flush=True)


def resident_async(n):
    for _ in range(n):
        step(f, m)


def resident_item(n):
    for _ in range(n):
        step(f, m).item()


def h2d_item(n):
    for a, b in prefetch_to_device([(fh, mh)] * n, dev):
        step(a, b).item()


def h2d_async(n):
    for a, b in prefetch_to_device([(fh, mh)] * n, dev):
        step(a, b)


def h2d_item_lagged(n):
    prev = None
    for a, b in prefetch_to_device([(fh, mh)] * n, dev):
        cur = step(a, b)
        host = torch.empty((), dtype=cur.dtype, pin_memory=True)
        host.copy_(cur, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if prev is not None:
            prev[1].synchronize()
            prev[0].item()
        prev = (host, ev)
    prev[1].synchronize()
    prev[0].item()


def host_only(n):
    # how long does the host need to enqueue one step? (GPU saturated, measure enqueue rate)
    t0 = time.perf_counter()
    for _ in range(n):
        step(f, m)
    print(f"   host enqueue time per step: {(time.perf_counter() - t0) * 1e3 / n:.2f} ms")


def h2d_only(n):
    for a, b in prefetch_to_device([(fh, mh)] * n, dev):
        pass


run("H pinned H2D copies alone (2 x 64 MB per step)", h2d_only)
run("A device-resident, async (bench 'value')", resident_async)
run("B device-resident, .item() every step", resident_item)
run("C pinned H2D prefetch + .item() every step (bench 'e2e')", h2d_item)
run("D pinned H2D prefetch, async", h2d_async)
run("E pinned H2D prefetch, result read one step late", h2d_item_lagged)
host_only(10)
torch.cuda.synchronize()
