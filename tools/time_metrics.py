"""Device timings of the evaluation-side kernels at 256^3: surface Hausdorff distance,
affine augmentation through the fused warp, label-map Dice, Jacobian statistics.
    python tools/time_metrics.py [size]"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import ops  # noqa: E402


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = "cuda"
    lin = torch.linspace(-1, 1, S, device=dev)
    zz, yy, xx = torch.meshgrid(lin, lin, lin, indexing="ij")
    a = (((zz - 0.1) ** 2 + yy ** 2 + xx ** 2) < 0.5).float()[None]
    b = ((zz ** 2 + (yy + 0.05) ** 2 / 1.2 + xx ** 2) < 0.45).float()[None]
    ms = timed(lambda: ops.hausdorff(a, b))
    hd = float(ops.hausdorff(a, b)[0, 0])
    print(f"hausdorff {S}^3: {ms:.3f} ms (distance {hd:.4f}); {2 * 2 * S ** 4 / ms / 1e6:.1f} G envelope terms/s")
    img = torch.rand(1, 1, S, S, S, device=dev)
    seg = torch.randint(0, 14, (1, 1, S, S, S), device=dev).float()
    gen = torch.Generator().manual_seed(0)
    ms = timed(lambda: kb.random_affine_augment(img, seg=seg, generator=gen))
    print(f"random_affine_augment img+seg {S}^3: {ms:.3f} ms ({4 * 4 * S ** 3 / ms / 1e6:.0f} GB/s algorithmic)")
    grid = kb.AffineTransform(matrix=torch.eye(4, device=dev)[None]).get_flow_field((1, 1, S, S, S))
    ms = timed(lambda: ops.jacobian_stats(grid.permute(0, 4, 1, 2, 3)))
    print(f"jacobian_stats {S}^3: {ms:.3f} ms ({12 * S ** 3 / ms / 1e6:.0f} GB/s algorithmic)")


if __name__ == "__main__":
    main()
