"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the exchange.

* Pairwise registration shards embarrassingly: pairs are independent, `shard_range` hands every
  rank a contiguous block and there is NO collective on the data path.
* Groupwise registration (keymorph/model.py:295-530) has exactly one cross-subject reduction per
  iteration, the mean of the keypoints (model.py:344).  Subjects are partitioned across ranks; each
  iteration all-reduces the local (K,3) keypoint sums (6 KB at K=512) and the subject count, then
  every rank refits its own subjects against the shared mean.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bind_cpu_to_gpu(device_index: int):
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (its NUMA node).  One process
    per GPU feeds its board from pinned host buffers: allocated after this call they land on the local node
    (first touch), so H2D / D2H copies do not cross the socket interconnect.  Returns the CPU list, or None when
    NVML / sched_setaffinity is unavailable or reports nothing usable (never raises)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) >= len(allowed):
            return None          # nothing to narrow (single node, or the mask is empty)
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 -- an optimisation, never a requirement
        return None


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced block of item indices for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def global_mean_points(local_points: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over ALL subjects of all ranks of (G_local,K,3) keypoints -> (1,K,3).
    One all-reduce of K*3+1 floats (sum of points and subject count packed together)."""
    k3 = local_points.shape[1] * local_points.shape[2]
    buf = torch.empty(k3 + 1, dtype=torch.float32, device=local_points.device)
    buf[:k3] = local_points.float().sum(dim=0).reshape(-1)
    buf[k3] = float(local_points.shape[0])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[:k3] / buf[k3]).reshape(1, local_points.shape[1], local_points.shape[2]).to(
        local_points.dtype)


def gather_all_points(local_points: torch.Tensor, group=None):
    """All subjects' keypoints on every rank: (G_local,K,3) -> ((G,K,3) in rank order, offset of this
    rank's block).  ONE all-gather of G_max*K*3+1 floats per rank (shards padded to the largest one, the
    subject count travels in the last slot)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_points, 0
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    g_local, K, d = local_points.shape
    # every rank derives the padded size from the same rule as shard_range: sizes differ by at most one
    cnt = torch.tensor([float(g_local)], device=local_points.device)
    gmax = torch.tensor([float(g_local)], device=local_points.device)
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    gmax = int(gmax.item())
    buf = torch.zeros(gmax * K * d + 1, dtype=torch.float32, device=local_points.device)
    buf[:g_local * K * d] = local_points.float().reshape(-1)
    buf[-1] = cnt[0]
    allbuf = torch.empty(ws * buf.numel(), dtype=torch.float32, device=local_points.device)
    dist.all_gather_into_tensor(allbuf, buf, group=group)
    allbuf = allbuf.reshape(ws, -1)
    counts = [int(c) for c in allbuf[:, -1].tolist()]
    parts = [allbuf[r, :counts[r] * K * d].reshape(counts[r], K, d) for r in range(ws)]
    return torch.cat(parts, 0).to(local_points.dtype), sum(counts[:rk])


def groupwise_iterate(local_points, register_fn, num_iters, group=None, mode="allreduce"):
    """keymorph/model.py:435-444 across ranks.  register_fn(points (G,K,3), mean (1,K,3)) returns the
    points registered to the mean.  Returns (aligned local points, mean at the start of the last
    iteration) -- the pair the reference uses to build the final flow fields (model.py:456-510).

    mode "allreduce": every iteration all-reduces the local keypoint sums (num_iters collectives of
    K*3+1 floats) and refits only the local subjects.
    mode "allgather": ONE all-gather of every subject's keypoints, then every rank iterates over ALL
    subjects locally (redundant fits, no further exchange): the mean is torch.mean over the same (G,K,3)
    tensor on every rank, i.e. bit-identical to the single-process result (option 2)."""
    if mode == "allgather":
        allp, off = gather_all_points(local_points, group)
        cur = allp.clone()
        mean = None
        for _ in range(num_iters):
            mean = torch.mean(cur, dim=0, keepdim=True)
            cur = register_fn(cur, mean)
        return cur[off:off + local_points.shape[0]].contiguous(), mean
    if mode != "allreduce":
        raise ValueError(f"unknown groupwise exchange mode {mode!r}")
    cur = local_points.clone()
    mean = None
    for _ in range(num_iters):
        mean = global_mean_points(cur, group)
        cur = register_fn(cur, mean)
    return cur, mean


def _parse_align(model, align_str, device):
    from .utils import str_or_float
    if align_str.startswith("tps"):
        return "tps", model._convert_tps_lmbda(1, str_or_float(align_str[4:])).to(device)
    return align_str, None


@torch.no_grad()
def extract_keypoints(model, local_imgs, batch=2):
    """Keypoints of the local subjects (G_local,1,D,H,W) -> (G_local,K,3), `batch` volumes per backbone pass."""
    return torch.cat([model.get_keypoints(local_imgs[i:i + batch]) for i in range(0, local_imgs.shape[0], batch)], 0)


@torch.no_grad()
def groupwise_iterate_points(model, pts, align_str, num_iters, group=None, mode="allreduce"):
    """-> (aligned local points, mean of the last iteration's start, number of collectives issued)."""
    from .transformations import deferred_singular_checks
    kind, lam = _parse_align(model, align_str, pts.device)
    with deferred_singular_checks():
        cur, mean = groupwise_iterate(pts, lambda p, m: model._register_to_mean(p, m, kind, lam), num_iters, group,
                                      mode)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    ncoll = 0 if not multi else (num_iters if mode == "allreduce" else 2)   # all-gather + its size all-reduce
    return cur, mean, ncoll


@torch.no_grad()
def groupwise_grids(model, pts, mean, align_str, local_imgs=None, consume=None):
    """Final flow fields of the local subjects: ORIGINAL keypoints against the last mean (model.py:456-510).
    consume=None returns the (G_local,D,H,W,3) grids; consume="warp" warps each subject with its grid as
    scripts/groupwise_register_eval.py:378-389 does (one grid alive at a time) and returns the sum of the
    warped volumes' means as a checksum."""
    kind, lam = _parse_align(model, align_str, pts.device)
    G = pts.shape[0]
    shape = (1, 1) + tuple(local_imgs.shape[2:]) if local_imgs is not None else None
    if consume is None:
        fixed = mean.expand(G, -1, -1).contiguous()
        aligner = model._make_aligner(kind, pts, fixed, None, None if lam is None else lam.reshape(-1)[:1].repeat(G))
        return aligner.get_flow_field((G, 1) + tuple(shape[2:]))
    from . import ops
    acc = torch.zeros((), dtype=torch.float64, device=pts.device)
    for i in range(G):
        grid = _one_grid(model, kind, lam, pts[i:i + 1], mean, shape)
        img_a, _ = ops.warp_loss(local_imgs[i:i + 1], None, grid=grid)
        acc += img_a.double().mean()
    return float(acc.item())


def _one_grid(model, kind, lam, points_m, mean, shape):
    aligner = model._make_aligner(kind, points_m, mean, None, lam)
    return aligner.get_flow_field(shape)


@torch.no_grad()
def groupwise_register_sharded(model, local_imgs, transform_type, num_iters, group=None,
                               want_grids=True, mode="allreduce"):
    """Groupwise registration of the subjects held by THIS rank (local_imgs: (G_local,1,D,H,W) on
    this rank's GPU) against the mean keypoints of all ranks.  Returns per align string
    {grouppoints_m, grouppoints_a, mean_points, [groupgrids]} for the local subjects."""
    from .utils import str_or_float
    if isinstance(transform_type, str):
        transform_type = [transform_type]
    pts = torch.cat([model.get_keypoints(local_imgs[i:i + 1]) for i in range(local_imgs.shape[0])], 0)
    out = {}
    for align_str in transform_type:
        if align_str.startswith("tps"):
            kind = "tps"
            lam = model._convert_tps_lmbda(1, str_or_float(align_str[4:])).to(pts.device)
        else:
            kind, lam = align_str, None
        cur, mean = groupwise_iterate(
            pts, lambda p, m: model._register_to_mean(p, m, kind, lam), num_iters, group, mode)
        res = {"grouppoints_m": pts, "grouppoints_a": cur, "mean_points": mean}
        if want_grids:
            G = pts.shape[0]
            fixed = mean.expand(G, -1, -1).contiguous()
            aligner = model._make_aligner(kind, pts, fixed, None,
                                          None if lam is None else lam.reshape(-1)[:1].repeat(G))
            res["groupgrids"] = aligner.get_flow_field(local_imgs.shape)
        out[align_str] = res
    return out
