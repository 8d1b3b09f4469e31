"""CPU oracle for the KeyMorph registration hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain restatement (torch-CPU / numpy, fp32 by default, fp64 on request) of the
reference algorithm at alanqrwang/keymorph @ dcb7996.  It exists to CHECK the CUDA engine and to
serve as the timed CPU baseline of bench.py; it is never imported by the product package
(keymorph_b200/) -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.

Pinning: tests/test_oracle_golden.py checks every function below against
  (a) the reference's own known-answer tests (test/test.py:117-253 CoM, :256-413 rigid,
      :416-480 affine), restated as fixtures, and
  (b) golden vectors produced by running the reference itself in the build container
      (oracle/gen_golden.py -> tests/golden/*.npz).
The arithmetic of the path lives in PyTorch (conv3d, group_norm, grid_sample, linalg.solve, svd),
pinned by the reference only as torch>=1.7 (setup.py:27); here it is torch 2.11 CPU kernels.
"""
from __future__ import annotations

import math
import re

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# grids and warps


def uniform_norm_grid(shape, dtype=torch.float32):
    """keymorph/utils.py:387-398 -- linspace(-1,1,S) per axis, 'ij' meshgrid, (D,H,W,3) in (z,y,x)."""
    axes = [torch.linspace(-1, 1, int(s)).to(dtype) for s in shape]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)


def align_img(grid, x, mode="bilinear"):
    """keymorph/utils.py:14-21."""
    return F.grid_sample(x, grid=grid, mode=mode, padding_mode="border", align_corners=False)


def grid_sample3d_numpy(x, grid, mode="bilinear"):
    """Independent restatement of ATen grid_sampler_3d (border padding, align_corners=False) in
    numpy fp32 -- pure loops over corners, used to pin the conventions of align_img:
    source index = ((g+1)*size-1)/2, clamp to [0,size-1], floor, 8-corner blend with out-of-range
    corners skipped; nearest = round-half-even of the clamped index."""
    x = np.asarray(x, dtype=np.float32)
    g = np.asarray(grid, dtype=np.float32)
    N, C, D, H, W = x.shape
    out = np.zeros((N, C) + g.shape[1:4], dtype=np.float32)

    def src(c, size):
        v = ((c + np.float32(1)) * np.float32(size) - np.float32(1)) / np.float32(2)
        return np.minimum(np.float32(size - 1), np.maximum(v, np.float32(0))).astype(np.float32)

    for n in range(N):
        ix, iy, iz = src(g[n, ..., 0], W), src(g[n, ..., 1], H), src(g[n, ..., 2], D)
        if mode == "nearest":
            xi = np.rint(ix).astype(np.int64)
            yi = np.rint(iy).astype(np.int64)
            zi = np.rint(iz).astype(np.int64)
            out[n] = x[n][:, zi, yi, xi]
            continue
        x0, y0, z0 = np.floor(ix), np.floor(iy), np.floor(iz)
        acc = np.zeros((C,) + ix.shape, dtype=np.float32)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    wx = (ix - x0) if dx else (x0 + 1 - ix)
                    wy = (iy - y0) if dy else (y0 + 1 - iy)
                    wz = (iz - z0) if dz else (z0 + 1 - iz)
                    xi, yi, zi = (x0 + dx).astype(np.int64), (y0 + dy).astype(np.int64), (z0 + dz).astype(np.int64)
                    ok = (xi < W) & (yi < H) & (zi < D)
                    xi, yi, zi = np.minimum(xi, W - 1), np.minimum(yi, H - 1), np.minimum(zi, D - 1)
                    wgt = ((wx * wy).astype(np.float32) * wz).astype(np.float32)
                    acc = acc + np.where(ok, x[n][:, zi, yi, xi] * wgt, np.float32(0)).astype(np.float32)
        out[n] = acc
    return out


# --------------------------------------------------------------------------------------------
# keypoint layer


def center_of_mass3d(vol, ij=True):
    """keymorph/layers.py:92-134 -- ReLU, three marginals, sum(lin*m)/(sum(m)+1e-8), *2-1."""
    v = F.relu(vol)
    n, c, dz, dy, dx = v.shape
    eps = 1e-8
    outs = []
    for size, keep in ((dx, (2, 3)), (dy, (2, 4)), (dz, (3, 4))):
        lin = torch.linspace(0, 1, size).to(v).view(1, 1, -1)
        m = v.sum(dim=keep)
        total = m.sum(dim=-1, keepdim=True) + eps
        outs.append((lin * m).sum(dim=-1, keepdim=True) / total)
    cx, cy, cz = outs
    order = [cz, cy, cx] if ij else [cx, cy, cz]
    return torch.cat(order, dim=-1) * 2 - 1


def weight_by_power(feat1, feat2):
    """keymorph/model.py:95-109."""
    p1 = F.relu(feat1).flatten(2).sum(-1)
    p2 = F.relu(feat2).flatten(2).sum(-1)
    w = p1 * p2
    return w / w.sum(dim=1)


def jacobian_determinant(disp):
    """keymorph/loss_ops.py:161-234 without scipy: central differences (-0.5, 0, 0.5) of the three
    components of a (1,3,D,H,W) fp32 field along z, y, x (each rounded to fp32 as
    scipy.ndimage.correlate does for a float32 input), + identity in fp64, cofactor determinant,
    2-voxel crop.  Returns the float64 (D-4,H-4,W-4) array."""
    import numpy as np
    d = np.asarray(disp.detach().cpu().numpy() if isinstance(disp, torch.Tensor) else disp)[0].astype(np.float32)

    def grad(a, axis):
        a = a.astype(np.float64)
        g = np.zeros_like(a)
        hi = [slice(None)] * 3
        lo = [slice(None)] * 3
        mid = [slice(None)] * 3
        hi[axis], lo[axis], mid[axis] = slice(2, None), slice(None, -2), slice(1, -1)
        g[tuple(mid)] = 0.5 * a[tuple(hi)] - 0.5 * a[tuple(lo)]
        return g.astype(np.float32).astype(np.float64)

    J = np.zeros((3, 3) + d.shape[1:], dtype=np.float64)
    for a in range(3):
        for b in range(3):
            J[a, b] = grad(d[b], a)
        J[a, a] += 1.0
    J = J[:, :, 2:-2, 2:-2, 2:-2]
    return (J[0, 0] * (J[1, 1] * J[2, 2] - J[1, 2] * J[2, 1]) - J[1, 0] * (J[0, 1] * J[2, 2] - J[0, 2] * J[2, 1])
            + J[2, 0] * (J[0, 1] * J[1, 2] - J[0, 2] * J[1, 1]))


def jdstd(disp):
    """keymorph/loss_ops.py:237-240."""
    return float(jacobian_determinant(disp).std())


def jdlessthan0(disp, as_percentage=False):
    """keymorph/loss_ops.py:243-248."""
    import numpy as np
    jd = jacobian_determinant(disp)
    n = int(np.count_nonzero(jd <= 0))
    return n / jd.size if as_percentage else n


# --------------------------------------------------------------------------------------------
# surface Hausdorff distance and the pairwise group metrics (keymorph/loss_ops.py:66-157, 405-551)


def surface_mask(vol):
    """keymorph/loss_ops.py:121-129: A minus binary_erosion(A) with the 6-connected element and
    border_value 0 (scipy.ndimage default): a set voxel is on the surface when one of its six
    neighbours is clear or lies outside the array."""
    import numpy as np
    a = np.asarray(vol).astype(bool)
    p = np.pad(a, 1, constant_values=False)
    er = (p[1:-1, 1:-1, 1:-1] & p[:-2, 1:-1, 1:-1] & p[2:, 1:-1, 1:-1] & p[1:-1, :-2, 1:-1] & p[1:-1, 2:, 1:-1]
          & p[1:-1, 1:-1, :-2] & p[1:-1, 1:-1, 2:])
    return a & ~er


def surface_distances(vol1, vol2, sampling=(1.25, 1.25, 10.0), chunk=2048):
    """keymorph/loss_ops.py:121-141 (_surfd) with scipy's distance_transform_edt replaced by its
    definition: for every surface voxel of one volume the Euclidean distance (voxel sizes
    `sampling`) to the nearest surface voxel of the other, both directions concatenated."""
    import numpy as np
    s1, s2 = surface_mask(vol1), surface_mask(vol2)
    c1 = np.argwhere(s1) * np.asarray(sampling, dtype=np.float64)
    c2 = np.argwhere(s2) * np.asarray(sampling, dtype=np.float64)

    def nearest(src, dst):
        out = np.empty(len(src))
        for i in range(0, len(src), chunk):
            d = src[i:i + chunk, None, :] - dst[None, :, :]
            out[i:i + chunk] = np.sqrt((d * d).sum(-1).min(1))
        return out

    return np.concatenate([nearest(c2, c1), nearest(c1, c2)])


def hausdorff_distance(test_seg, gt_seg, sampling=(1.25, 1.25, 10.0)):
    """keymorph/loss_ops.py:144-157: mean over the batch of the maximum surface distance of channel 0."""
    a = test_seg.detach().cpu().numpy() if isinstance(test_seg, torch.Tensor) else test_seg
    b = gt_seg.detach().cpu().numpy() if isinstance(gt_seg, torch.Tensor) else gt_seg
    return sum(surface_distances(a[i, 0], b[i, 0], sampling).max() for i in range(len(a))) / len(a)


def fast_dice(x, y):
    """keymorph/loss_ops.py:66-106 without the histogram detour: Dice of argmax label maps per label
    present in either input, eps 1e-5 in the denominator, averaged; 1 for a single label."""
    import numpy as np
    x = np.asarray(x).argmax(1)
    y = np.asarray(y).argmax(1)
    labels = np.unique(np.concatenate([np.unique(x), np.unique(y)]))
    if len(labels) <= 1:
        return 1.0
    return float(np.mean([2 * np.sum((x == l) & (y == l)) / (np.sum(x == l) + np.sum(y == l) + 1e-5)
                          for l in labels]))


def avg_pairwise(batch, fn):
    """keymorph/loss_ops.py:414-435 / 499-527: mean of fn over the unordered pairs of a (G, ...) batch."""
    tot, num = 0, 0
    for i in range(len(batch)):
        for j in range(i + 1, len(batch)):
            tot = tot + fn(batch[i:i + 1], batch[j:j + 1])
            num += 1
    return tot / num


# --------------------------------------------------------------------------------------------
# closed-form aligners


def _homog(p):
    return torch.cat([p, torch.ones_like(p[..., :1])], dim=-1)


def fit_affine(x, y, w=None):
    """keymorph/keypoint_aligners.py:76-114 -- A = Y W X^T (X W X^T)^-1, (N,3,4)."""
    X = _homog(x.float() if x.dtype != torch.float64 else x).transpose(1, 2)   # (N,4,K)
    Y = (y.float() if y.dtype != torch.float64 else y).transpose(1, 2)         # (N,3,K)
    if w is not None:
        Wm = torch.diag_embed(w.to(X.dtype))
        left = torch.bmm(torch.bmm(X, Wm), X.transpose(1, 2))
        right = torch.bmm(Wm, X.transpose(1, 2))
    else:
        left = torch.bmm(X, X.transpose(1, 2))
        right = X.transpose(1, 2)
    inv = torch.inverse(left)
    return torch.bmm(Y, torch.bmm(right, inv))


def fit_rigid(p1, p2, w=None):
    """keymorph/keypoint_aligners.py:151-213 -- Arun et al.; returns (N,3,4) = [R | T]."""
    dt = torch.float64 if p1.dtype == torch.float64 else torch.float32
    a = p1.to(dt).transpose(1, 2)
    b = p2.to(dt).transpose(1, 2)
    if w is not None:
        wt = w.to(dt)
        ca = (a * wt[:, None]).sum(2, keepdim=True)      # weighted SUM (weights assumed normalised)
        cb = (b * wt[:, None]).sum(2, keepdim=True)
    else:
        ca = a.mean(2, keepdim=True)
        cb = b.mean(2, keepdim=True)
    qa, qb = a - ca, b - cb
    if w is not None:
        qa, qb = qa * wt[:, None], qb * wt[:, None]
    Hm = torch.bmm(qa, qb.transpose(1, 2))
    U, _, Vt = torch.linalg.svd(Hm)
    V = Vt.transpose(1, 2)
    R = torch.bmm(V, U.transpose(1, 2))
    # :199-203: dets is stacked along axis 1 and tiled along axis 2 -> the last ROW of V is negated
    # (R <- diag(1, 1, sign det) V U^T), not the last column as in Arun et al.
    sgn = torch.sign(torch.det(R))
    flip = torch.ones_like(V)
    flip[:, 2, :] = sgn[:, None]
    V = V * flip
    R = torch.bmm(V, U.transpose(1, 2))
    T = cb - torch.bmm(R, ca)
    return torch.cat([R, T], dim=-1)


def square(m34):
    """keymorph/transformations.py:32-35."""
    n = m34.shape[0]
    sq = torch.eye(4, dtype=m34.dtype).repeat(n, 1, 1)
    sq[:, :3, :] = m34
    return sq


def aligner_matrices(points_m, points_f, w=None, kind="affine"):
    """keymorph/keypoint_aligners.py:67-74 + transformations.py:22-30: the aligners fit
    FIXED -> MOVING (inverse_transform_matrix) and invert it for transform_matrix."""
    fit = fit_affine if kind == "affine" else fit_rigid
    inv = square(fit(points_f, points_m, w)).to(points_m.dtype)
    return torch.inverse(inv), inv            # (transform_matrix, inverse_transform_matrix)


def transform_points(m44, pts):
    """keymorph/transformations.py:81-114: p' = M[:3,:] [p;1]."""
    return torch.bmm(m44[:, :3, :], _homog(pts).transpose(1, 2)).transpose(1, 2)


def affine_flow_field(inverse_matrix, shape):
    """keymorph/transformations.py:37-79: grid (1,D,H,W,3) in (x,y,z) order."""
    g = uniform_norm_grid(shape, inverse_matrix.dtype).reshape(1, -1, 3)
    moved = transform_points(inverse_matrix, g)
    return moved.reshape(1, *[int(s) for s in shape], 3).flip(-1)


# --------------------------------------------------------------------------------------------
# thin-plate splines


def tps_d(a, b):
    """keymorph/keypoint_aligners.py:322-334."""
    return torch.sqrt(torch.square(a[:, :, None, :] - b[:, None, :, :]).sum(-1) + 1e-6)


def tps_u(r):
    """keymorph/keypoint_aligners.py:336-339."""
    return r ** 2 * torch.log(r + 1e-6)


def tps_fit(c_src, c_dst, lmbda, w=None):
    """keymorph/keypoint_aligners.py:276-363 -- theta (N,K+4,3): one dense solve per output dim."""
    dt = c_src.dtype
    bs, T, dim = c_src.shape
    U = tps_u(tps_d(c_src, c_src))
    lam = lmbda.to(dt).view(bs, 1, 1)
    if w is not None:
        Kmat = U + torch.reciprocal(torch.diag_embed(w.to(dt)) + 1e-6) * lam   # dense reciprocal (sic)
    else:
        Kmat = U + torch.eye(T, dtype=dt).repeat(bs, 1, 1) * lam
    P = torch.ones(bs, T, dim + 1, dtype=dt)
    P[:, :, 1:] = c_src
    A = torch.zeros(bs, T + dim + 1, T + dim + 1, dtype=dt)
    A[:, :T, :T] = Kmat
    A[:, :T, T:] = P
    A[:, T:, :T] = P.transpose(1, 2)
    cols = []
    for k in range(dim):
        v = torch.zeros(bs, T + dim + 1, dtype=dt)
        v[:, :T] = c_dst[..., k]
        cols.append(torch.linalg.solve(A, v))
    return torch.stack(cols, dim=-1)


def tps_transform(theta, ctrl, pts):
    """keymorph/keypoint_aligners.py:399-433."""
    dim = ctrl.shape[-1]
    wts, aff = theta[:, :-(dim + 1), :], theta[:, -(dim + 1):, :]
    U = tps_u(tps_d(ctrl, pts))
    P = torch.cat([torch.ones_like(pts[..., :1]), pts], dim=-1)
    return torch.bmm(P, aff) + torch.bmm(U.transpose(1, 2), wts)


def tps_flow_field(points_m, points_f, lmbda, shape, w=None, chunk=1 << 16):
    """keymorph/keypoint_aligners.py:365-397,435-449 (sub-grid evaluation; the chunk size does not
    change the per-voxel arithmetic)."""
    theta = tps_fit(points_f, points_m, lmbda, w)
    g = uniform_norm_grid(shape, points_f.dtype).reshape(1, -1, 3)
    out = torch.empty_like(g)
    for s in range(0, g.shape[1], chunk):
        out[:, s:s + chunk] = tps_transform(theta, points_f, g[:, s:s + chunk])
    return out.reshape(1, *[int(s) for s in shape], 3).flip(-1)


def tps_forward_points(points_m, points_f, lmbda, pts, w=None):
    """keymorph/keypoint_aligners.py:451-465 (a second, forward fit)."""
    theta = tps_fit(points_m, points_f, lmbda, w)
    return tps_transform(theta, points_m, pts)


# --------------------------------------------------------------------------------------------
# real-world coordinates (keymorph/utils.py:243-354; keypoint_aligners.py:53-74,116-147,435-465)


def convert_points_norm2voxel(points, grid_sizes):
    """keymorph/utils.py:243-258: [-1, 1] -> voxel index space, -1 is the outer face of voxel 0."""
    return ((points + 1) * torch.as_tensor(grid_sizes).to(points)) / 2 - 0.5


def convert_points_voxel2norm(points, grid_sizes):
    """keymorph/utils.py:261-276."""
    return (2 * (points + 0.5) / torch.as_tensor(grid_sizes).to(points)) - 1


def convert_points_voxel2real(points, affine):
    """keymorph/utils.py:279-296: homogeneous product with the (N,d+1,d+1) affine."""
    return torch.bmm(affine.to(points), _homog(points).transpose(1, 2)).transpose(1, 2)[:, :, :-1]


def convert_points_real2voxel(points, affine):
    """keymorph/utils.py:299-322: the same with torch.inverse(affine)."""
    return torch.bmm(torch.inverse(affine.to(points)), _homog(points).transpose(1, 2)).transpose(1, 2)[:, :, :-1]


def convert_points_norm2real(points, affine, sizes):
    """keymorph/utils.py:325-338."""
    return convert_points_voxel2real(convert_points_norm2voxel(points, sizes), affine)


def convert_points_real2norm(points, affine, sizes):
    """keymorph/utils.py:341-354."""
    return convert_points_voxel2norm(convert_points_real2voxel(points, affine), sizes)


def register_points_real_world(points_f, points_m, transform, shape_f, shape_m, aff_f, aff_m, grid_shape):
    """The aligners with align_in_real_world_coords=True (keypoint_aligners.py:53-74,116-147,255-274,
    435-465): keypoints are moved to scanner space before the fit; every transformed point goes
    norm -> real (source image) -> fit -> real -> norm (target image).  Batch 1, like the reference."""
    kind, lam = parse_transform(transform)
    rf = convert_points_norm2real(points_f, aff_f, shape_f)
    rm = convert_points_norm2real(points_m, aff_m, shape_m)
    g = uniform_norm_grid(grid_shape, points_f.dtype).reshape(1, -1, 3)
    g_real = convert_points_norm2real(g, aff_f, shape_f)
    pm_real = convert_points_norm2real(points_m, aff_m, shape_m)
    res = {}
    if kind in ("rigid", "affine"):
        tm, inv = aligner_matrices(rm, rf, None, kind)
        res["matrix"], res["inverse"] = tm, inv
        moved = transform_points(inv, g_real)
        fwd = transform_points(tm, pm_real)
    else:
        lmbda = torch.tensor(lam, dtype=points_f.dtype).repeat(points_f.shape[0])
        theta = tps_fit(rf, rm, lmbda)
        res["inverse_theta"] = theta
        moved = tps_transform(theta, rf, g_real)
        fwd = tps_transform(tps_fit(rm, rf, lmbda), rm, pm_real)
    res["grid"] = convert_points_real2norm(moved, aff_m, shape_m).reshape(1, *[int(v) for v in grid_shape], 3).flip(-1)
    res["points_a"] = convert_points_real2norm(fwd, aff_f, shape_f)
    return res


# --------------------------------------------------------------------------------------------
# losses


def mse_loss(pred, target):
    """keymorph/loss_ops.py:9-13."""
    return F.mse_loss(pred, target)


def dice_loss(pred, target, hard=False, ign_first_ch=False, return_regions=False):
    """keymorph/loss_ops.py:16-63 (eps = 1)."""
    n, c = target.shape[:2]
    t = target.reshape(n, c, -1)
    p = pred.reshape(n, c, -1)
    if hard:
        idx = torch.argmax(p, dim=1, keepdim=True)
        p = torch.zeros_like(p).scatter(1, idx, 1.0)
    if ign_first_ch:
        t, p = t[:, 1:], p[:, 1:]
    num = (2 * t * p).sum(2) + 1
    den = (p * p).sum(2) + (t * t).sum(2) + 1
    loss = 1 - num / den
    return loss.mean(0) if return_regions else loss.mean()


# --------------------------------------------------------------------------------------------
# backbones (functional, driven by a reference-format state dict without the "module." prefix)


def _strip(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def _single_conv(sd, prefix, x, groups=8):
    """keymorph/unet3d/buildingblocks.py:39-132 with order 'gcr': GroupNorm -> Conv3d(no bias) -> ReLU."""
    c = x.shape[1]
    g = groups if c >= groups else 1
    x = F.group_norm(x, g, sd[prefix + ".groupnorm.weight"], sd[prefix + ".groupnorm.bias"], 1e-5)
    return F.relu(F.conv3d(x, sd[prefix + ".conv.weight"], None, padding=1))


def _double_conv(sd, prefix, x):
    x = _single_conv(sd, prefix + ".SingleConv1", x)
    return _single_conv(sd, prefix + ".SingleConv2", x)


def unet3d_forward(sd, x, num_levels=4, num_truncated=0):
    """keymorph/unet3d/model.py:115-151 (AbstractUNet.forward) for UNet3D / TruncatedUNet3D with
    DoubleConv blocks, max-pool encoders, nearest-upsample + concat decoders, final 1x1x1 conv."""
    sd = _strip(sd)
    feats = []
    for i in range(num_levels):
        if i > 0:
            x = F.max_pool3d(x, 2)
        x = _double_conv(sd, f"encoders.{i}.basic_module", x)
        feats.insert(0, x)
    feats = feats[1:]
    for i in range(num_levels - 1 - num_truncated):
        skip = feats[i]
        up = F.interpolate(x, size=skip.shape[2:], mode="nearest")
        x = _double_conv(sd, f"decoders.{i}.basic_module", torch.cat((skip, up), dim=1))
    return F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])


def convnet_forward(sd, x, norm_type="instance"):
    """keymorph/net.py:7-36 + keymorph/layers.py:137-187: 9 x [conv(bias) -> norm -> ReLU -> (pool)]."""
    sd = _strip(sd)
    for b in range(1, 10):
        x = F.conv3d(x, sd[f"block{b}.conv.weight"], sd[f"block{b}.conv.bias"], padding=1)
        if norm_type == "instance":
            x = F.instance_norm(x, eps=1e-5)
        elif norm_type != "none":
            raise NotImplementedError(norm_type)
        x = F.relu(x)
        if b in (2, 4, 6, 8):
            x = F.max_pool3d(x, 2)
    return x


def backbone_forward(kind, sd, x, **kw):
    if kind == "truncatedunet":
        return unet3d_forward(sd, x, kw.get("num_levels", 4), kw.get("num_truncated", 1))
    if kind == "unet":
        return unet3d_forward(sd, x, kw.get("num_levels", 4), 0)
    if kind == "conv":
        return convnet_forward(sd, x, kw.get("norm_type", "instance"))
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------
# pipeline


def parse_transform(s):
    """keymorph/model.py:134-140,200-207."""
    if s in ("rigid", "affine"):
        return s, None
    m = re.match(r"^tps_(.*)$", s)
    if not m:
        raise ValueError(f"Invalid transform_type {s}")
    return "tps", float(m.group(1))


def register_points(points_f, points_m, transform, shape, w=None, return_aligned_points=True):
    """The alignment half of keymorph/model.py:198-288 for one pair (batch 1)."""
    kind, lam = parse_transform(transform)
    res = {"points_f": points_f, "points_m": points_m, "points_weights": w}
    if kind in ("rigid", "affine"):
        tm, inv = aligner_matrices(points_m, points_f, w, kind)
        res["matrix"] = tm
        res["grid"] = affine_flow_field(inv, shape)
        res["tps_lmbda"] = None
        if return_aligned_points:
            res["points_a"] = transform_points(tm, points_m)
    else:
        lmbda = torch.tensor(lam).repeat(points_f.shape[0])
        res["tps_lmbda"] = lmbda
        res["grid"] = tps_flow_field(points_m, points_f, lmbda, shape, w)
        if return_aligned_points:
            res["points_a"] = tps_forward_points(points_m, points_f, lmbda, points_m, w)
    return res


def keymorph_forward(kind, sd, img_f, img_m, transform_type="affine", weight_keypoints=None, **kw):
    """keymorph/model.py:142-289 (eval mode, no AMP)."""
    if not isinstance(transform_type, (list, tuple)):
        transform_type = [transform_type]
    feat_f = backbone_forward(kind, sd, img_f, **kw)
    feat_m = backbone_forward(kind, sd, img_m, **kw)
    pf, pm = center_of_mass3d(feat_f), center_of_mass3d(feat_m)
    w = weight_by_power(feat_f, feat_m) if weight_keypoints == "power" else None
    return {t: register_points(pf, pm, t, img_f.shape[2:], w) for t in transform_type}


def groupwise_points(group_points, transform, num_iters):
    """keymorph/model.py:331-444: iterate {mean keypoints -> register every subject to the mean}.
    Returns (aligned points after num_iters, mean points taken at the START of the last iteration)."""
    kind, lam = parse_transform(transform)
    cur = group_points.clone()
    mean = None
    for _ in range(num_iters):
        mean = cur.mean(dim=0, keepdim=True)
        nxt = torch.zeros_like(cur)
        for i in range(len(cur)):
            pm = cur[i:i + 1]
            if kind == "tps":
                lmbda = torch.tensor(lam).repeat(1)
                nxt[i:i + 1] = tps_forward_points(pm, mean, lmbda, pm)
            else:
                tm, _ = aligner_matrices(pm, mean, None, kind)
                nxt[i:i + 1] = transform_points(tm, pm)
        cur = nxt
    return cur, mean


# --------------------------------------------------------------------------------------------
# synthetic inputs shared by tests and bench


def gaussian_phantom(size, seed, n_blobs=12, dtype=torch.float32):
    """Sum of isotropic Gaussians with centres U(-0.6,0.6)^3, sigma U(0.1,0.35), amplitude U(0,1),
    normalised to max 1.  Returns (1,1,S,S,S)."""
    gen = torch.Generator().manual_seed(seed)
    ctr = torch.rand(n_blobs, 3, generator=gen, dtype=torch.float64) * 1.2 - 0.6
    sig = torch.rand(n_blobs, generator=gen, dtype=torch.float64) * 0.25 + 0.1
    amp = torch.rand(n_blobs, generator=gen, dtype=torch.float64)
    lin = torch.linspace(-1, 1, size, dtype=torch.float64)
    vol = torch.zeros(size, size, size, dtype=torch.float64)
    for b in range(n_blobs):
        gz = torch.exp(-0.5 * ((lin - ctr[b, 0]) / sig[b]) ** 2)
        gy = torch.exp(-0.5 * ((lin - ctr[b, 1]) / sig[b]) ** 2)
        gx = torch.exp(-0.5 * ((lin - ctr[b, 2]) / sig[b]) ** 2)
        vol += amp[b] * gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
    vol = vol / vol.max()
    return vol.to(dtype)[None, None]


def affine_matrix_3d(scale, offset, angle, shear, dtype=torch.float32):
    """keymorph/augmentation.py:81-158 build_affine_matrix for isotropic fixed parameters
    (scale = 1+s on every axis, same offset / rotation angle / shear everywhere)."""
    s, o, a, z = 1 + scale, offset, angle, shear
    Ms = torch.diag(torch.tensor([s, s, s, 1.0], dtype=torch.float64))
    Mt = torch.eye(4, dtype=torch.float64)
    Mt[:3, 3] = o
    c, sn = math.cos(a), math.sin(a)
    R1 = torch.tensor([[1, 0, 0, 0], [0, c, -sn, 0], [0, sn, c, 0], [0, 0, 0, 1]], dtype=torch.float64)
    R2 = torch.tensor([[c, 0, sn, 0], [0, 1, 0, 0], [-sn, 0, c, 0], [0, 0, 0, 1]], dtype=torch.float64)
    R3 = torch.tensor([[c, -sn, 0, 0], [sn, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64)
    Mz = torch.eye(4, dtype=torch.float64)
    Mz[0, 1] = Mz[0, 2] = Mz[1, 0] = Mz[1, 2] = Mz[2, 0] = Mz[2, 1] = z
    M = Mz @ Ms @ Mt @ (R3 @ R2 @ R1)
    return M.to(dtype)[None]


def affine_matrix_3d_params(scale, offset, theta, shear, dtype=torch.float32):
    """keymorph/augmentation.py:85-158 for per-axis parameters: scale (3,), offset (3,), theta (3,)
    (rotations about axis 0, 1, 2, applied in that order), shear (6,) filling the off-diagonal of Mz
    row by row.  Returns (1,4,4)."""
    scale, offset, theta, shear = (torch.as_tensor(v, dtype=torch.float64).reshape(-1) for v in
                                   (scale, offset, theta, shear))
    Ms = torch.diag(torch.cat([scale, torch.ones(1, dtype=torch.float64)]))
    Mt = torch.eye(4, dtype=torch.float64)
    Mt[:3, 3] = offset
    c, sn = torch.cos(theta), torch.sin(theta)
    R1, R2, R3 = (torch.eye(4, dtype=torch.float64) for _ in range(3))
    R1[1, 1], R1[1, 2], R1[2, 1], R1[2, 2] = c[0], -sn[0], sn[0], c[0]
    R2[0, 0], R2[0, 2], R2[2, 0], R2[2, 2] = c[1], sn[1], -sn[1], c[1]
    R3[0, 0], R3[0, 1], R3[1, 0], R3[1, 1] = c[2], -sn[2], sn[2], c[2]
    Mz = torch.eye(4, dtype=torch.float64)
    Mz[0, 1], Mz[0, 2], Mz[1, 0], Mz[1, 2], Mz[2, 0], Mz[2, 1] = shear
    return (Mz @ Ms @ Mt @ (R3 @ R2 @ R1)).to(dtype)[None]


def deform(M, img=None, seg=None, points=None):
    """keymorph/augmentation.py:160-167: bilinear image / nearest segmentation through
    AffineTransform(matrix=M).get_flow_field, points through the forward matrix."""
    res = ()
    if img is not None:
        grid = affine_flow_field(torch.inverse(M), img.shape[2:])
        res += (align_img(grid, img),)
    if seg is not None:
        grid = affine_flow_field(torch.inverse(M), seg.shape[2:])
        res += (align_img(grid, seg, mode="nearest"),)
    if points is not None:
        res += (transform_points(M, points),)
    return res[0] if len(res) == 1 else res


def affine_augment(img, params, seg=None):
    """keymorph/augmentation.py:160-167,223-254: warp with AffineTransform(matrix=M).get_flow_field."""
    M = affine_matrix_3d(*params, dtype=img.dtype)
    grid = affine_flow_field(torch.inverse(M), img.shape[2:])
    out = align_img(grid, img)
    if seg is not None:
        return out, align_img(grid, seg, mode="nearest")
    return out
