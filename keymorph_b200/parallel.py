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


def groupwise_iterate(local_points, register_fn, num_iters, group=None):
    """keymorph/model.py:435-444 across ranks.  register_fn(points (G,K,3), mean (1,K,3)) returns the
    points registered to the mean.  Returns (aligned local points, mean at the start of the last
    iteration) -- the pair the reference uses to build the final flow fields (model.py:456-510)."""
    cur = local_points.clone()
    mean = None
    for _ in range(num_iters):
        mean = global_mean_points(cur, group)
        cur = register_fn(cur, mean)
    return cur, mean


@torch.no_grad()
def groupwise_register_sharded(model, local_imgs, transform_type, num_iters, group=None,
                               want_grids=True):
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
            pts, lambda p, m: model._register_to_mean(p, m, kind, lam), num_iters, group)
        res = {"grouppoints_m": pts, "grouppoints_a": cur, "mean_points": mean}
        if want_grids:
            G = pts.shape[0]
            fixed = mean.expand(G, -1, -1).contiguous()
            aligner = model._make_aligner(kind, pts, fixed, None,
                                          None if lam is None else lam.reshape(-1)[:1].repeat(G))
            res["groupgrids"] = aligner.get_flow_field(local_imgs.shape)
        out[align_str] = res
    return out
