"""Groupwise registration sharded over the ranks of a torchrun job (NCCL all-reduce of the mean
keypoints), checked against the single-process result on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/groupwise_multi_gpu.py [S] [K] [G]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import parallel  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                         num_levels=4, is_segmentation=False, conv_padding=1).eval().to(dev)
model = kb.KeyMorph(net, K, 3).eval()
base = O.gaussian_phantom(S, 1000)
subjects = torch.cat([O.affine_augment(base, (0.02 * i, 0.01 * i - 0.03, 0.05 * i, 0.0)) for i in range(G)])
mine = subjects[list(parallel.shard_range(G, rank, world))].to(dev)
types = ["rigid", "affine", "tps_1"]
iters = 5
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
res = parallel.groupwise_register_sharded(model, mine, types, iters)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0

ok = True
if rank == 0:
    ref = model.groupwise_register(subjects.to(dev), transform_type=types, device=dev, num_iters=iters,
                                   log_to_console=False, save_results_to_disk=False)
    idx = list(parallel.shard_range(G, 0, world))
    for t in types:
        e_a = (res[t]["grouppoints_a"] - ref[t]["grouppoints_a"][idx]).abs().max().item()
        e_g = (res[t]["groupgrids"] - ref[t]["groupgrids"][idx]).abs().max().item()
        print(f"groupwise {t} world={world}: sharded vs single-process: points {e_a:.2e} grids {e_g:.2e}")
        ok &= e_a < 1e-5 and e_g < 1e-4
    print(f"sharded groupwise ({G} subjects, {S}^3, K={K}, {iters} iters, {len(types)} transforms): {dt * 1e3:.1f} ms "
          f"-> {'PASS' if ok else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
