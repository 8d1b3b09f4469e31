"""e2e (host buffers) throughput of the CUDA-graph replay path with the two input-staging variants, next to the
eager path, in one process.  python tools/graph_e2e_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import model as M  # noqa: E402
from keymorph_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
ctx = {"dev": dev, "world": 1, "rank": 0, "sync": torch.cuda.synchronize, "max": float}
img_f_cpu, base_m_cpu, minv = bench.host_pair(0)
img_f_host = img_f_cpu.pin_memory()
img_m = ops.warp_loss(base_m_cpu.to(dev), None, mat34=minv.to(dev)[:, :3])[0]
img_m_host = img_m.cpu().pin_memory()
for name in ("affine", "tps"):
    cfg = bench.CONFIGS[name]
    net = torch.nn.DataParallel(bench.seeded_backbone(cfg["K"]).to(dev))
    eager = kb.KeyMorph(net, cfg["K"], 3, fused_warp=True).eval()
    r = bench.measure_e2e(eager, name, img_f_host, img_m_host, 20, ctx)
    print(f"{name} eager e2e: {r['ms_per_step']:.2f} ms")
    for stage_kernel in (True, False, True):
        M.GRAPH_STAGE_WITH_KERNEL = stage_kernel
        gm = kb.KeyMorph(net, cfg["K"], 3, fused_warp=True, cuda_graph=True).eval()
        for clone in (True, False):
            r = bench.measure_e2e(gm, name, img_f_host, img_m_host, 20, ctx, clone_outputs=clone)
            print(f"{name} graph e2e (stage with kernel={stage_kernel}, clone outputs={clone}): {r['ms_per_step']:.2f} ms  "
                  f"{list(gm.graph_state().values())}")
        del gm
        torch.cuda.empty_cache()
