"""Tensor-level wrappers over the C ABI: torch owns device memory and streams, the kernels do
the work.  Every function requires CUDA tensors and raises otherwise (no CPU path)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (KM_CONV_COM, KM_CONV_RELU, KM_CONV_STATS, KM_COORD_AFFINE, KM_COORD_GRID,
                   KM_COORD_TPS, KM_INTERP_BILINEAR, KM_INTERP_NEAREST)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.KMError("keymorph_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _mode(mode: str) -> int:
    if mode in ("bilinear", "trilinear"):
        return KM_INTERP_BILINEAR
    if mode == "nearest":
        return KM_INTERP_NEAREST
    raise ValueError(f"unsupported interpolation mode {mode!r}")


def act_dtype():
    """torch dtype of the 16-bit activation / packed-weight tensors of the backbone: torch.float16 (default,
    the AMP dtype of the reference, keymorph/model.py:175-177) or torch.bfloat16 -- see set_operand_dtype."""
    return torch.float16 if _lib.query("km_operand_is_fp16") else torch.bfloat16


def set_operand_dtype(name):
    """Select the tcgen05 operand type of the backbone: "fp16" (11-bit significand; default) or "bf16"
    (exponent range of fp32).  Same tensor-core rate; affects tensors created afterwards (engines re-pack
    their weights when it changes)."""
    name = {torch.float16: "fp16", torch.bfloat16: "bf16"}.get(name, name)
    if name not in ("fp16", "bf16"):
        raise ValueError(f"operand dtype must be 'fp16' or 'bf16', got {name!r}")
    rc = _lib.load().km_set_option(_lib.KM_OPT_OPERAND_FP16, 1 if name == "fp16" else 0)
    if rc != 0:
        raise _lib.KMError("km_set_option(KM_OPT_OPERAND_FP16) failed")


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------- warp
def grid_sample3d(x, grid, mode="bilinear"):
    """F.grid_sample(x, grid, mode, padding_mode='border', align_corners=False) for 5-D inputs."""
    _need_cuda(x, grid)
    x, grid = _f32c(x), _f32c(grid)
    N, Cc, Di, Hi, Wi = x.shape
    assert grid.shape[0] == N and grid.shape[-1] == 3 and grid.dim() == 5
    _, Do, Ho, Wo, _ = grid.shape
    out = torch.empty((N, Cc, Do, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_grid_sample3d", _ptr(x), _ptr(grid), _ptr(out), N, Cc, Di, Hi, Wi, Do, Ho,
                  Wo, _mode(mode), _stream())
    return out


def flow_field_affine(mat34, shape):
    """mat34: (N,3,4) rows in (z,y,x); shape: (D,H,W) -> grid (N,D,H,W,3) in (x,y,z) order."""
    _need_cuda(mat34)
    mat34 = _f32c(mat34)
    N = mat34.shape[0]
    D, H, W = (int(s) for s in shape)
    grid = torch.empty((N, D, H, W, 3), dtype=torch.float32, device=mat34.device)
    with torch.cuda.device(mat34.device):
        _lib.call("km_flow_field_affine", _ptr(mat34), _ptr(grid), N, D, H, W, _stream())
    return grid


def flow_field_tps(ctrl, theta, shape):
    _need_cuda(ctrl, theta)
    ctrl, theta = _f32c(ctrl), _f32c(theta)
    N, K, _ = ctrl.shape
    assert theta.shape == (N, K + 4, 3)
    D, H, W = (int(s) for s in shape)
    grid = torch.empty((N, D, H, W, 3), dtype=torch.float32, device=ctrl.device)
    with torch.cuda.device(ctrl.device):
        _lib.call("km_flow_field_tps", _ptr(ctrl), _ptr(theta), _ptr(grid), N, K, D, H, W,
                  _stream())
    return grid


def points_transform_affine(mat34, pts):
    _need_cuda(mat34, pts)
    mat34, pts = _f32c(mat34), _f32c(pts)
    N, P, _ = pts.shape
    out = torch.empty_like(pts)
    with torch.cuda.device(pts.device):
        _lib.call("km_points_transform_affine", _ptr(mat34), _ptr(pts), _ptr(out), N, P, _stream())
    return out


def points_transform_tps(ctrl, theta, pts):
    _need_cuda(ctrl, theta, pts)
    ctrl, theta, pts = _f32c(ctrl), _f32c(theta), _f32c(pts)
    N, K, _ = ctrl.shape
    P = pts.shape[1]
    out = torch.empty_like(pts)
    with torch.cuda.device(pts.device):
        _lib.call("km_points_transform_tps", _ptr(ctrl), _ptr(theta), _ptr(pts), _ptr(out), N, K,
                  P, _stream())
    return out


def warp_loss(moving, fixed=None, *, mat34=None, ctrl=None, theta=None, grid=None,
              mode="bilinear", store=True, want_grid=False):
    """Fused warp (+ loss sums).  Exactly one of mat34 / (ctrl, theta) / grid selects where the
    sampling coordinates come from.  Returns (warped or None, sums (N,C,4) fp64 or None) with
    sums[..., :] = [sum (a-f)^2, sum a*f, sum a*a, sum f*f]; with want_grid=True (mat34 / TPS modes)
    a third value, the (N,D,H,W,3) flow field written by the same pass."""
    _need_cuda(moving, fixed, mat34, ctrl, theta, grid)
    moving, fixed = _f32c(moving), _f32c(fixed)
    N, Cc, D, H, W = moving.shape
    K = 0
    a = th = g = None
    if mat34 is not None:
        coord, a = KM_COORD_AFFINE, _f32c(mat34)
    elif ctrl is not None:
        coord, a, th = KM_COORD_TPS, _f32c(ctrl), _f32c(theta)
        K = a.shape[1]
    elif grid is not None:
        coord, g = KM_COORD_GRID, _f32c(grid)
        assert g.shape == (N, D, H, W, 3)
    else:
        raise ValueError("warp_loss needs mat34, (ctrl, theta) or grid")
    out = torch.empty_like(moving) if store else None
    gout = None
    if want_grid:
        if g is not None:
            raise ValueError("want_grid needs mat34 or (ctrl, theta)")
        gout = torch.empty((N, D, H, W, 3), dtype=torch.float32, device=moving.device)
    sums = ws = None
    if fixed is not None:
        assert fixed.shape == moving.shape
        sums = torch.empty((N, Cc, 4), dtype=torch.float64, device=moving.device)
        ws = _ws(_lib.query("km_warp_loss_workspace_bytes", N, Cc), moving.device)
    with torch.cuda.device(moving.device):
        _lib.call("km_warp_loss", coord, _ptr(a), _ptr(th), K, _ptr(g), _ptr(moving), _ptr(fixed),
                  _ptr(out), _ptr(gout), _ptr(sums), _ptr(ws), N, Cc, D, H, W, _mode(mode), _stream())
    return (out, sums, gout) if want_grid else (out, sums)


def warp_labels_dice(labels_m, labels_f, num_classes, *, mat34=None, grid=None, want_labels=False):
    """Segmentation warp + soft and hard Dice sums straight from uint8 label maps (N,D,H,W) or
    (N,1,D,H,W): returns (soft_sums, hard_sums[, warped hard labels]) with the (N,C,4) fp64 layout of
    warp_loss / pair_stats, without ever building the C-channel one-hot volumes."""
    _need_cuda(labels_m, labels_f, mat34, grid)
    if int(num_classes) > 255:
        raise ValueError("warp_labels_dice takes uint8 label ids: num_classes must be <= 255 (remap the labels)")
    lm = labels_m.reshape(labels_m.shape[0], *labels_m.shape[-3:]).to(torch.uint8).contiguous()
    lf = labels_f.reshape(labels_f.shape[0], *labels_f.shape[-3:]).to(torch.uint8).contiguous()
    assert lm.shape == lf.shape
    N, D, H, W = lm.shape
    Cc = int(num_classes)
    if mat34 is not None:
        coord, a, g = KM_COORD_AFFINE, _f32c(mat34), None
    elif grid is not None:
        coord, a, g = KM_COORD_GRID, None, _f32c(grid)
        assert g.shape == (N, D, H, W, 3)
    else:
        raise ValueError("warp_labels_dice needs mat34 or grid")
    soft = torch.empty((N, Cc, 4), dtype=torch.float64, device=lm.device)
    hard = torch.empty_like(soft)
    lout = torch.empty_like(lm) if want_labels else None
    ws = _ws(_lib.query("km_warp_labels_workspace_bytes", N, Cc), lm.device)
    with torch.cuda.device(lm.device):
        _lib.call("km_warp_labels_dice", coord, _ptr(a), _ptr(g), _ptr(lm), _ptr(lf), _ptr(lout), _ptr(soft),
                  _ptr(hard), _ptr(ws), N, Cc, D, H, W, _stream())
    return (soft, hard, lout) if want_labels else (soft, hard)


def jacobian_stats(field):
    """field: fp32 (N,3,D,H,W) -- any strides, e.g. grid.permute(0,4,1,2,3) -- -> (N,4) fp64
    [std, count(det <= 0), mean, #interior voxels] of the Jacobian determinant (loss_ops.py:161-247)."""
    _need_cuda(field)
    if field.dtype != torch.float32:
        field = field.float()
    N, Cc, D, H, W = field.shape
    assert Cc == 3
    out = torch.empty((N, 4), dtype=torch.float64, device=field.device)
    ws = _ws(_lib.query("km_jacobian_stats_workspace_bytes", N), field.device)
    sn, sc, sz, sy, sx = field.stride()
    with torch.cuda.device(field.device):
        _lib.call("km_jacobian_stats", _ptr(field), sn, sc, sz, sy, sx, _ptr(out), _ptr(ws), N, D, H, W,
                  _stream())
    return out


def hausdorff(a, b, sampling=(1.25, 1.25, 10.0)):
    """a, b: fp32 (N,D,H,W) binary volumes (voxel != 0), dense in (D,H,W), any batch stride -- e.g.
    seg[:, 0] -> (N,2) fp64 [symmetric surface Hausdorff distance, empty flag] (loss_ops.py:120-157)."""
    _need_cuda(a, b)
    assert a.dim() == 4 and a.shape == b.shape and a.dtype == torch.float32 and b.dtype == torch.float32
    N, D, H, W = a.shape
    dense = (H * W, W, 1)
    if a.stride()[1:] != dense:
        a = a.contiguous()
    if b.stride()[1:] != dense:
        b = b.contiguous()
    out = torch.empty((N, 2), dtype=torch.float64, device=a.device)
    ws = _ws(_lib.query("km_hausdorff_workspace_bytes", D, H, W), a.device)
    sz, sy, sx = (float(v) for v in sampling)
    with torch.cuda.device(a.device):
        _lib.call("km_hausdorff", _ptr(a), _ptr(b), a.stride(0), b.stride(0), N, D, H, W, sz, sy, sx,
                  _ptr(out), _ptr(ws), _stream())
    return out


def pair_stats(pred, target, hard=False):
    """sums (N,C,4) fp64 = [sum (p-t)^2, sum p*t, sum p*p, sum t*t] over the flattened spatial dims;
    hard=True replaces pred by one_hot(argmax_c pred)."""
    _need_cuda(pred, target)
    pred, target = _f32c(pred), _f32c(target)
    assert pred.shape == target.shape
    N, Cc = pred.shape[0], pred.shape[1]
    M = pred[0, 0].numel()
    sums = torch.empty((N, Cc, 4), dtype=torch.float64, device=pred.device)
    ws = _ws(_lib.query("km_pair_stats_workspace_bytes", N, Cc, M, int(hard)), pred.device)
    with torch.cuda.device(pred.device):
        _lib.call("km_pair_stats", _ptr(pred), _ptr(target), _ptr(sums), _ptr(ws), N, Cc, M,
                  int(hard), _stream())
    return sums


def argmax_channels(pred):
    _need_cuda(pred)
    pred = _f32c(pred)
    N, Cc = pred.shape[0], pred.shape[1]
    M = pred[0, 0].numel()
    labels = torch.empty((N,) + tuple(pred.shape[2:]), dtype=torch.int32, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.call("km_argmax_channels", _ptr(pred), _ptr(labels), N, Cc, M, _stream())
    return labels


# --------------------------------------------------------------------------------------- CoM
def com3d(heat, ij=True, return_mass=False):
    _need_cuda(heat)
    heat = _f32c(heat)
    N, K, D, H, W = heat.shape
    pts = torch.empty((N, K, 3), dtype=torch.float32, device=heat.device)
    mass = torch.empty((N, K), dtype=torch.float32, device=heat.device) if return_mass else None
    ws = _ws(_lib.query("km_com3d_workspace_bytes", N, K), heat.device)
    with torch.cuda.device(heat.device):
        _lib.call("km_com3d", _ptr(heat), _ptr(pts), _ptr(mass), _ptr(ws), N, K, D, H, W, int(ij),
                  _stream())
    return (pts, mass) if return_mass else pts


# --------------------------------------------------------------------------------------- fits
def _fit(name, x, y, w):
    _need_cuda(x, y, w)
    x, y, w = _f32c(x), _f32c(y), _f32c(w)
    N, K, d = x.shape
    assert d == 3 and y.shape == x.shape
    A = torch.empty((N, 4, 4), dtype=torch.float32, device=x.device)
    Ainv = torch.empty_like(A)
    status = torch.empty((N,), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call(name, _ptr(x), _ptr(y), _ptr(w), _ptr(A), _ptr(Ainv), _ptr(status), N, K,
                  _stream())
    return A, Ainv, status


def fit_affine(x, y, w=None):
    """A44 = square(argmin_A |y - A [x;1]|), its inverse and a per-sample singularity status."""
    return _fit("km_fit_affine", x, y, w)


def fit_rigid(x, y, w=None):
    return _fit("km_fit_rigid", x, y, w)


def inverse44(m):
    _need_cuda(m)
    m = _f32c(m)
    N = m.shape[0]
    inv = torch.empty_like(m)
    status = torch.empty((N,), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        _lib.call("km_inverse44", _ptr(m), _ptr(inv), _ptr(status), N, _stream())
    return inv, status


def tps_fit(c_src, c_dst, lmbda, w=None):
    """theta (N,K+4,3) solving the TPS system that maps c_src -> c_dst (fp64 LU on the device)."""
    _need_cuda(c_src, c_dst, lmbda, w)
    c_src, c_dst, w = _f32c(c_src), _f32c(c_dst), _f32c(w)
    N, K, _ = c_src.shape
    lmbda = _f32c(lmbda.to(c_src.device)).reshape(-1)
    if lmbda.numel() == 1 and N > 1:
        lmbda = lmbda.repeat(N)
    assert lmbda.numel() == N
    theta = torch.empty((N, K + 4, 3), dtype=torch.float32, device=c_src.device)
    status = torch.empty((N,), dtype=torch.int32, device=c_src.device)
    ws = _ws(_lib.query("km_tps_fit_workspace_bytes", N, K), c_src.device)
    with torch.cuda.device(c_src.device):
        _lib.call("km_tps_fit", _ptr(c_src), _ptr(c_dst), _ptr(lmbda), _ptr(w), _ptr(theta),
                  _ptr(status), _ptr(ws), N, K, _stream())
    return theta, status


# --------------------------------------------------------------------------------------- backbone
def pack_weights(w):
    """fp32 (Cout,Cin,kd,kh,kw) -> bf16 [tap][Cout][Cin]."""
    _need_cuda(w)
    w = _f32c(w)
    Cout, Cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    out = torch.empty((taps, Cout, Cin), dtype=act_dtype(), device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("km_pack_weights", _ptr(w), _ptr(out), Cout, Cin, taps, _stream())
    return out


def conv_nparts():
    return _lib.query("km_conv_nparts")


def red_nparts():
    return _lib.query("km_pool_nparts")


def conv3d_tc(x, wp, bias=None, relu=False, want_stats=False, want_com=False, store=True):
    """x: bf16 (N,D,H,W,Cin); wp: bf16 (taps,Cout,Cin).  Returns (out|None, stats|None, com|None)."""
    _need_cuda(x, wp, bias)
    assert x.dtype == act_dtype() and wp.dtype == act_dtype()
    x, wp = x.contiguous(), wp.contiguous()
    N, D, H, W, Cin = x.shape
    taps, Cout, Cin2 = wp.shape
    assert Cin2 == Cin
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0) | \
        (KM_CONV_COM if want_com else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x.device) if store else None
    nparts = conv_nparts()
    stats = torch.empty((nparts, N, Cout, 2), dtype=torch.float32, device=x.device) \
        if want_stats else None
    com = torch.empty((nparts, N, Cout, 4), dtype=torch.float32, device=x.device) \
        if want_com else None
    bias = _f32c(bias)
    with torch.cuda.device(x.device):
        _lib.call("km_conv3d_tc", _ptr(x), _ptr(wp), _ptr(bias), _ptr(out), _ptr(stats), _ptr(com),
                  N, Cin, Cout, D, H, W, taps, flags, _stream())
    return out, stats, com


USE_PAIR_CONV = True      # engine switch: 2-CTA kernel for the Cout in {64,128} layers (A/B testing)


def pair_supported(Cin, Cout, D, H, W):
    return bool(_lib.query("km_conv3d_tc_pair_supported", Cin, Cout, D, H, W))


def conv3d_tc_pair(x, wp, relu=False, want_stats=False):
    """2-CTA (cta_group::2) tcgen05 conv, 3x3x3, Cout in {64,128}.  Same arguments as conv3d_tc."""
    _need_cuda(x, wp)
    assert x.dtype == act_dtype() and wp.dtype == act_dtype()
    x, wp = x.contiguous(), wp.contiguous()
    N, D, H, W, Cin = x.shape
    taps, Cout, Cin2 = wp.shape
    assert taps == 27 and Cin2 == Cin
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x.device)
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x.device) \
        if want_stats else None
    with torch.cuda.device(x.device):
        _lib.call("km_conv3d_tc_pair", _ptr(x), _ptr(wp), _ptr(out), _ptr(stats), N, Cin, Cout, D, H, W,
                  flags, _stream())
    return out, stats


USE_ZFOLD_PAIR = True     # engine switch: z-folded 2-CTA kernel for the Cout = 64 layers (A/B testing)
# ... also for 32 -> 64 (enc1.c2) with its GroupNorm folded in: N = 192 per MMA instead of 64 (571 -> 412 us for
# two 128^3 volumes, tools/time_enc1c2.py).  The fold stores that layer's input un-normalised: fine with fp16
# operands (256^3 keypoint error 2.07e-3 -> 2.21e-3 max), not with bf16 (round 1: fails the 256^3 criterion).
# None = automatic (fp16 only); True / False force it.
USE_ZFOLD_PAIR_CIN32 = None


def zfold_pair_cin32_enabled():
    if USE_ZFOLD_PAIR_CIN32 is None:
        return act_dtype() == torch.float16
    return bool(USE_ZFOLD_PAIR_CIN32)


def zfold_pair_supported(Cin, Cout, D, H, W):
    return bool(_lib.query("km_conv3d_zfold_pair_supported", Cin, Cout, D, H, W))


def pack_weights_zfold_pair(w):
    """fp32 (64,Cin,3,3,3) -> bf16 (3 rotations, 3 dx, 3 dy, 192, Cin) for conv3d_zfold_pair."""
    _need_cuda(w)
    w = _f32c(w)
    Cout, Cin = w.shape[0], w.shape[1]
    out = torch.empty((3, 3, 3, 3 * Cout, Cin), dtype=act_dtype(), device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("km_pack_weights_zfold_pair", _ptr(w), _ptr(out), Cout, Cin, _stream())
    return out


def conv3d_zfold_pair(x, wz, relu=False, want_stats=False, pool=False, store=True):
    """z-folded 2-CTA tcgen05 conv (3x3x3, pad 1; Cout 64 or 32).  x: bf16 (N,D,H,W,Cin).  Returns
    (out | None, stats | None) or, with pool=True, (out | None, pooled, stats | None) like conv3d_zfold."""
    _need_cuda(x, wz)
    assert x.dtype == act_dtype() and wz.dtype == act_dtype()
    x, wz = x.contiguous(), wz.contiguous()
    N, D, H, W, Cin = x.shape
    Cout = wz.shape[3] // 3
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x.device) if store else None
    pooled = torch.empty((N, D // 2, H // 2, W // 2, Cout), dtype=act_dtype(), device=x.device) \
        if pool else None
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x.device) \
        if want_stats else None
    with torch.cuda.device(x.device):
        _lib.call("km_conv3d_zfold_pair", _ptr(x), _ptr(wz), _ptr(out), _ptr(pooled), _ptr(stats), N, Cin,
                  Cout, D, H, W, flags, _stream())
    return (out, pooled, stats) if pool else (out, stats)


def zfold_supported(Cin, Cout, D, H, W):
    return bool(_lib.query("km_conv3d_zfold_supported", Cin, Cout, D, H, W))


def pack_weights_zfold(w):
    """fp32 (Cout,Cin,3,3,3) -> bf16 (3 rotations, 9 (dx,dy), 3*Cout, Cin) for conv3d_zfold."""
    _need_cuda(w)
    w = _f32c(w)
    Cout, Cin = w.shape[0], w.shape[1]
    out = torch.empty((3, 9, 3 * Cout, Cin), dtype=act_dtype(), device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("km_pack_weights_zfold", _ptr(w), _ptr(out), Cout, Cin, _stream())
    return out


def conv3d_zfold(x, wz, relu=False, want_stats=False, pool=False, store=True):
    """z-folded tcgen05 conv (Cin=16 -> Cout=32, 3x3x3, pad 1).  x: bf16 (N,D,H,W,16).
    Returns (out | None, stats | None) or, with pool=True, (out | None, pooled, stats | None) where
    pooled = MaxPool3d(2)(out) and the stats describe the pooled tensor."""
    _need_cuda(x, wz)
    assert x.dtype == act_dtype() and wz.dtype == act_dtype()
    x, wz = x.contiguous(), wz.contiguous()
    N, D, H, W, Cin = x.shape
    Cout = wz.shape[2] // 3
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x.device) if store else None
    pooled = torch.empty((N, D // 2, H // 2, W // 2, Cout), dtype=act_dtype(), device=x.device) \
        if pool else None
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x.device) \
        if want_stats else None
    with torch.cuda.device(x.device):
        _lib.call("km_conv3d_zfold", _ptr(x), _ptr(wz), _ptr(out), _ptr(pooled), _ptr(stats), N, Cin,
                  Cout, D, H, W, flags, _stream())
    return (out, pooled, stats) if pool else (out, stats)


USE_GN_FOLD = True
# The stem -> conv_zf pair can be folded too (the stem then runs once instead of twice).  There the fold REPLACES
# an fp32 normalisation before the single 16-bit rounding by a rounding of the un-centred activation, and a flat
# background then carries one coherent rounding offset under every centre of mass.  Measured at 256^3, K = 256
# (tools/accuracy_256.py, profiles/r02_accuracy_256_operand_types_and_folds.log), keypoint error vs the fp32 oracle:
#   fp16 operands: 3.50e-3 / 1.45e-4 (max / mean) without, 2.83e-3 / 1.63e-4 with the stem fold -- no loss;
#   bf16 operands: 2.74e-2 / 1.01e-3 without, 3.88e-2 / 1.18e-3 with -- 3 mantissa bits fewer make the offset visible.
# None = automatic: folded with fp16 operands (the default type), not folded with bf16.  True / False force it.
USE_GN_FOLD_STEM = None


def gn_fold_stem_enabled():
    if USE_GN_FOLD_STEM is None:
        return act_dtype() == torch.float16
    return bool(USE_GN_FOLD_STEM)
# The fold is also built for the plain CTA-pair kernel (conv3d_tc_pair_gn, parity-tested) but stays off
# in the engine: its staged epilogue is on the critical path, the bias loads cost more (+0.3 ms over the
# four layers) than the four in-place normalisation passes they replace (0.2 ms), and at 256^3 the extra
# un-centred bf16 roundings push the mean keypoint error over torch's own autocast drift.
USE_GN_FOLD_TC_PAIR = False


def conv3d_zfold_gn(x_raw, w, scale, shift, relu=False, want_stats=False, pool=False, store=True):
    """conv3d_zfold of GroupNorm(x_raw) without the normalisation pass: x_raw bf16 (N,D,H,W,16) is the
    un-normalised activation, w the fp32 (32,16,3,3,3) weights, scale / shift (N,16) from
    norm_finalize.  Same return convention as conv3d_zfold."""
    _need_cuda(x_raw, w, scale, shift)
    assert x_raw.dtype == act_dtype()
    x_raw, w, scale, shift = x_raw.contiguous(), _f32c(w), _f32c(scale), _f32c(shift)
    N, D, H, W, Cin = x_raw.shape
    Cout = w.shape[0]
    assert w.shape[1] == Cin and scale.numel() == N * Cin and shift.numel() == N * Cin
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x_raw.device) if store else None
    pooled = torch.empty((N, D // 2, H // 2, W // 2, Cout), dtype=act_dtype(), device=x_raw.device) \
        if pool else None
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x_raw.device) \
        if want_stats else None
    ws = _ws(_lib.query("km_conv3d_zfold_gn_workspace_bytes", N), x_raw.device)
    with torch.cuda.device(x_raw.device):
        _lib.call("km_conv3d_zfold_gn", _ptr(x_raw), _ptr(w), _ptr(scale), _ptr(shift), _ptr(out), _ptr(pooled),
                  _ptr(stats), _ptr(ws), N, Cin, Cout, D, H, W, flags, _stream())
    return (out, pooled, stats) if pool else (out, stats)


def upsample2(x):
    """bf16 (N,Dc,Hc,Wc,C) -> (N,2Dc,2Hc,2Wc,C), nearest neighbour."""
    _need_cuda(x)
    assert x.dtype == act_dtype()
    x = x.contiguous()
    N, Dc, Hc, Wc, Cc = x.shape
    out = torch.empty((N, 2 * Dc, 2 * Hc, 2 * Wc, Cc), dtype=act_dtype(), device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_upsample2_ndhwc", _ptr(x), _ptr(out), N, Cc, Dc, Hc, Wc, _stream())
    return out


def conv3d_zfold_pair_gn(x_raw, w, scale, shift, relu=False, want_stats=False, pool=False, store=True, x1=None):
    """conv3d_zfold_pair of GroupNorm(x_raw) without the normalisation pass (see conv3d_zfold_gn).  With
    x1 the input is cat(x_raw, x1) along the channels, read in place through two tensor maps."""
    _need_cuda(x_raw, w, scale, shift, x1)
    assert x_raw.dtype == act_dtype()
    x_raw, w, scale, shift = x_raw.contiguous(), _f32c(w), _f32c(scale), _f32c(shift)
    N, D, H, W, Cin0 = x_raw.shape
    Cin1 = 0
    if x1 is not None:
        assert x1.dtype == act_dtype() and x1.shape[:4] == x_raw.shape[:4]
        x1 = x1.contiguous()
        Cin1 = x1.shape[4]
    Cin = Cin0 + Cin1
    Cout = w.shape[0]
    assert w.shape[1] == Cin and scale.numel() == N * Cin and shift.numel() == N * Cin
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x_raw.device) if store else None
    pooled = torch.empty((N, D // 2, H // 2, W // 2, Cout), dtype=act_dtype(), device=x_raw.device) \
        if pool else None
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x_raw.device) \
        if want_stats else None
    ws = _ws(_lib.query("km_conv3d_zfold_pair_gn_workspace_bytes", N, Cin, Cout), x_raw.device)
    with torch.cuda.device(x_raw.device):
        _lib.call("km_conv3d_zfold_pair_gn_cat", _ptr(x_raw), _ptr(x1), Cin0, Cin1, _ptr(w), _ptr(scale),
                  _ptr(shift), _ptr(out), _ptr(pooled), _ptr(stats), _ptr(ws), N, Cout, D, H, W, flags, _stream())
    return (out, pooled, stats) if pool else (out, stats)


USE_COARSE_UPCONV = True   # engine switch: upsampled half of a decoder's first conv on the coarse lattice


def up2_supported(Cu, Cout, Dc, Hc, Wc):
    return bool(_lib.query("km_conv3d_up2_supported", Cu, Cout, Dc, Hc, Wc))


def conv3d_up2_gn(x_coarse, w, scale, Cs):
    """Partial sums of conv3d(GN(cat(skip, upsample2(x_coarse)))) over the upsampled channels [Cs, Cs + Cu),
    computed on the coarse lattice with 8 pre-summed taps per output parity class (the upsampled tensor is never
    written).  -> (N, 2Dc, 2Hc, 2Wc, Cout) 16-bit, to be passed as `addend` to conv3d_zfold_pair_gn_add."""
    _need_cuda(x_coarse, w, scale)
    assert x_coarse.dtype == act_dtype()
    x_coarse, w = x_coarse.contiguous(), _f32c(w)
    scale = None if scale is None else _f32c(scale)
    N, Dc, Hc, Wc, Cu = x_coarse.shape
    Cout = w.shape[0]
    assert w.shape[1] == Cs + Cu and (scale is None or scale.numel() == N * (Cs + Cu))
    out = torch.empty((N, 2 * Dc, 2 * Hc, 2 * Wc, Cout), dtype=act_dtype(), device=x_coarse.device)
    ws = _ws(_lib.query("km_conv3d_up2_gn_workspace_bytes", N, Cu, Cout), x_coarse.device)
    with torch.cuda.device(x_coarse.device):
        _lib.call("km_conv3d_up2_gn", _ptr(x_coarse), _ptr(w), _ptr(scale), Cs, Cu, _ptr(out), _ptr(ws), N, Cout,
                  Dc, Hc, Wc, _stream())
    return out


def conv3d_zfold_pair_gn_add(x_raw, w, scale, shift, addend, relu=False, want_stats=False, kernel="zfold_pair"):
    """The skip half of the same layer: input channels [0, Cs) from x_raw, plus `addend` (conv3d_up2_gn), then
    the folded-norm bias of ALL channels, ReLU and statistics.  kernel: "zfold_pair" (Cout 64, large volumes) or
    "tc_pair" (Cout 64 / 128)."""
    _need_cuda(x_raw, w, scale, shift, addend)
    assert x_raw.dtype == act_dtype() and addend.dtype == act_dtype()
    x_raw, w, scale, shift, addend = x_raw.contiguous(), _f32c(w), _f32c(scale), _f32c(shift), addend.contiguous()
    N, D, H, W, Cs = x_raw.shape
    Cout = w.shape[0]
    Cu = w.shape[1] - Cs
    assert Cu > 0 and scale.numel() == N * (Cs + Cu) and shift.numel() == N * (Cs + Cu)
    assert tuple(addend.shape) == (N, D, H, W, Cout)
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x_raw.device)
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x_raw.device) \
        if want_stats else None
    ws = _ws(_lib.query(f"km_conv3d_{kernel}_gn_workspace_bytes", N, Cs, Cout), x_raw.device)
    with torch.cuda.device(x_raw.device):
        _lib.call(f"km_conv3d_{kernel}_gn_add", _ptr(x_raw), Cs, Cu, _ptr(w), _ptr(scale), _ptr(shift),
                  _ptr(addend), _ptr(out), _ptr(stats), _ptr(ws), N, Cout, D, H, W, flags, _stream())
    return out, stats


def conv3d_tc_pair_gn(x_raw, w, scale, shift, relu=False, want_stats=False):
    """conv3d_tc_pair of GroupNorm(x_raw) without the normalisation pass (see conv3d_zfold_gn)."""
    _need_cuda(x_raw, w, scale, shift)
    assert x_raw.dtype == act_dtype()
    x_raw, w, scale, shift = x_raw.contiguous(), _f32c(w), _f32c(scale), _f32c(shift)
    N, D, H, W, Cin = x_raw.shape
    Cout = w.shape[0]
    assert w.shape[1] == Cin and scale.numel() == N * Cin and shift.numel() == N * Cin
    flags = (KM_CONV_RELU if relu else 0) | (KM_CONV_STATS if want_stats else 0)
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x_raw.device)
    stats = torch.empty((conv_nparts(), N, Cout, 2), dtype=torch.float32, device=x_raw.device) \
        if want_stats else None
    ws = _ws(_lib.query("km_conv3d_tc_pair_gn_workspace_bytes", N, Cin, Cout), x_raw.device)
    with torch.cuda.device(x_raw.device):
        _lib.call("km_conv3d_tc_pair_gn", _ptr(x_raw), _ptr(w), _ptr(scale), _ptr(shift), _ptr(out), _ptr(stats),
                  _ptr(ws), N, Cin, Cout, D, H, W, flags, _stream())
    return out, stats


def conv1x1_com(x, wp, bias=None):
    """Final 1x1x1 conv + ReLU + centre-of-mass partials without the heat map.
    x: bf16 (N,D,H,W,Cin); wp: bf16 (1,Cout,Cin) with Cout % 128 == 0 -> com (nparts,N,Cout,4)."""
    _need_cuda(x, wp, bias)
    assert x.dtype == act_dtype() and wp.dtype == act_dtype()
    x, wp = x.contiguous(), wp.contiguous()
    N, D, H, W, Cin = x.shape
    taps, Cout, Cin2 = wp.shape
    assert taps == 1 and Cin2 == Cin
    com = torch.empty((_lib.query("km_conv1x1_com_nparts"), N, Cout, 4), dtype=torch.float32,
                      device=x.device)
    bias = _f32c(bias)
    with torch.cuda.device(x.device):
        _lib.call("km_conv1x1_com", _ptr(x), _ptr(wp), _ptr(bias), _ptr(com), N, Cin, Cout, D, H, W,
                  _stream())
    return com


def com_finalize(com, return_mass=False):
    nparts, N, K, _ = com.shape
    pts = torch.empty((N, K, 3), dtype=torch.float32, device=com.device)
    mass = torch.empty((N, K), dtype=torch.float32, device=com.device) if return_mass else None
    with torch.cuda.device(com.device):
        _lib.call("km_com_finalize", _ptr(com), nparts, _ptr(pts), _ptr(mass), N, K, _stream())
    return (pts, mass) if return_mass else pts


def volume_stats(x):
    """x: fp32 (N, ...) -> partial stats (nparts, N, 1, 2)."""
    _need_cuda(x)
    x = _f32c(x)
    N = x.shape[0]
    M = x[0].numel()
    stats = torch.empty((red_nparts(), N, 1, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_volume_stats", _ptr(x), _ptr(stats), N, M, _stream())
    return stats


def channel_stats(x):
    """x: bf16 NDHWC -> partial stats (nparts, N, C, 2)."""
    _need_cuda(x)
    x = x.contiguous()
    N, Cc = x.shape[0], x.shape[-1]
    nvox = x[0, ..., 0].numel()
    stats = torch.empty((red_nparts(), N, Cc, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_channel_stats", _ptr(x), _ptr(stats), N, Cc, nvox, _stream())
    return stats


def norm_finalize(stats0, count0, gamma, beta, groups, eps=1e-5, stats1=None, count1=0.0, rep1=1.0):
    """partial stats -> per-(n,c) scale / shift of GroupNorm(groups) (InstanceNorm: groups=C)."""
    nparts0, N, C0, _ = stats0.shape
    C1 = 0 if stats1 is None else stats1.shape[2]
    nparts1 = 0 if stats1 is None else stats1.shape[0]
    Cc = C0 + C1
    scale = torch.empty((N, Cc), dtype=torch.float32, device=stats0.device)
    shift = torch.empty_like(scale)
    gamma, beta = _f32c(gamma), _f32c(beta)
    with torch.cuda.device(stats0.device):
        _lib.call("km_norm_finalize", _ptr(stats0), nparts0, C0, float(count0), _ptr(stats1),
                  nparts1, C1, float(count1), float(rep1), _ptr(gamma), _ptr(beta), int(groups),
                  float(eps), _ptr(scale), _ptr(shift), N, _stream())
    return scale, shift


def norm_apply(src0, scale, shift, src1=None, relu=False, pool=False, out=None):
    """bf16 NDHWC normalise (+ReLU) (+concat of the nearest-upsampled src1) (+MaxPool3d(2))."""
    _need_cuda(src0, src1, scale, shift)
    src0 = src0.contiguous()
    N, D, H, W, C0 = src0.shape
    C1 = D1 = H1 = W1 = 0
    if src1 is not None:
        src1 = src1.contiguous()
        _, D1, H1, W1, C1 = src1.shape
    oshape = (N, D // 2, H // 2, W // 2, C0) if pool else (N, D, H, W, C0 + C1)
    if out is None:
        out = torch.empty(oshape, dtype=act_dtype(), device=src0.device)
    assert tuple(out.shape) == oshape and out.is_contiguous()
    with torch.cuda.device(src0.device):
        _lib.call("km_norm_apply", _ptr(src0), C0, _ptr(src1), C1, D1, H1, W1, _ptr(scale),
                  _ptr(shift), _ptr(out), N, D, H, W, int(relu), int(pool), _stream())
    return out


def maxpool2_stats(x):
    _need_cuda(x)
    x = x.contiguous()
    N, D, H, W, Cc = x.shape
    out = torch.empty((N, D // 2, H // 2, W // 2, Cc), dtype=act_dtype(), device=x.device)
    stats = torch.empty((red_nparts(), N, Cc, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_maxpool2_stats", _ptr(x), _ptr(out), _ptr(stats), N, Cc, D, H, W, _stream())
    return out, stats


def conv3d_stem(x, w, bias=None, in_scale=None, in_shift=None, out_scale=None, out_shift=None,
                relu_pre=False, relu_post=False, store=True, want_stats=True):
    """x: fp32 (N,1,D,H,W); w: fp32 (Cout,1,3,3,3).  v = [relu_pre](conv(in_scale*x+in_shift)+bias);
    returns (bf16 (N,D,H,W,Cout) of [relu_post](out_scale*v+out_shift) or None, partial stats of v
    or None).  store=False is the statistics pass (nothing is written but the partials)."""
    _need_cuda(x, w, bias, in_scale, in_shift, out_scale, out_shift)
    x, w, bias = _f32c(x), _f32c(w), _f32c(bias)
    in_scale, in_shift = _f32c(in_scale), _f32c(in_shift)
    out_scale, out_shift = _f32c(out_scale), _f32c(out_shift)
    N, Cin, D, H, W = x.shape
    assert Cin == 1
    Cout = w.shape[0]
    out = torch.empty((N, D, H, W, Cout), dtype=act_dtype(), device=x.device) if store else None
    stats = None
    if want_stats:
        nparts = _lib.query("km_stem_nparts", N, D, H, W)
        stats = torch.empty((nparts, N, Cout, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_conv3d_stem", _ptr(x), _ptr(w), _ptr(bias), _ptr(in_scale), _ptr(in_shift),
                  _ptr(out_scale), _ptr(out_shift), _ptr(out), _ptr(stats), N, Cout, D, H, W,
                  int(relu_pre), int(relu_post), _stream())
    return out, stats


def ndhwc_to_ncdhw(x):
    _need_cuda(x)
    x = x.contiguous()
    N, D, H, W, Cc = x.shape
    out = torch.empty((N, Cc, D, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_ndhwc_bf16_to_ncdhw_f32", _ptr(x), _ptr(out), N, Cc, D, H, W, _stream())
    return out


def ncdhw_to_ndhwc(x):
    _need_cuda(x)
    x = _f32c(x)
    N, Cc, D, H, W = x.shape
    out = torch.empty((N, D, H, W, Cc), dtype=act_dtype(), device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("km_ncdhw_f32_to_ndhwc_bf16", _ptr(x), _ptr(out), N, Cc, D, H, W, _stream())
    return out
