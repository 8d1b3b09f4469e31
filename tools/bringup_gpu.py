"""GPU bring-up checks: every kernel against plain torch ops on the same device.

Each group runs in its own subprocess (a device-side trap poisons the CUDA context) under a
timeout, and the script keeps going after a failure.  Usage on the GPU box:
    python tools/bringup_gpu.py [group ...]      # groups: warp com fit tps loss misc conv
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["warp", "com", "fit", "tps", "loss", "misc", "conv", "conv_big", "convcom"]


def _report(name, got, ref, tol):
    import torch
    err = (got.double() - ref.double()).abs().max().item()
    scale = ref.double().abs().max().item()
    ok = err <= tol * max(1.0, scale)
    print(f"  [{'ok' if ok else 'FAIL'}] {name}: max|err|={err:.3e} (ref max {scale:.3e}, tol {tol:g})",
          flush=True)
    return ok


def run_warp():
    import torch
    import torch.nn.functional as F
    from keymorph_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    for (C, Di, Hi, Wi, Do, Ho, Wo) in [(1, 32, 32, 32, 32, 32, 32), (3, 17, 20, 23, 16, 12, 20),
                                        (2, 8, 9, 10, 5, 6, 7)]:
        x = torch.randn(2, C, Di, Hi, Wi, device=dev)
        grid = (torch.rand(2, Do, Ho, Wo, 3, device=dev) * 2.4 - 1.2)
        for mode in ("bilinear", "nearest"):
            ref = F.grid_sample(x, grid, mode=mode, padding_mode="border", align_corners=False)
            refc = F.grid_sample(x.cpu(), grid.cpu(), mode=mode, padding_mode="border",
                                 align_corners=False)
            got = ops.grid_sample3d(x, grid, mode)
            ok &= _report(f"grid_sample {mode} C={C} vs cuda", got, ref, 1e-5 if mode == "bilinear" else 0)
            nbad = (got.cpu() != refc).sum().item()
            print(f"      vs cpu: mismatching elements {nbad} / {refc.numel()}, max diff "
                  f"{(got.cpu() - refc).abs().max().item():.3e}")
    # affine flow field vs restatement
    mat = torch.eye(4, device=dev)[None, :3].repeat(2, 1, 1) + 0.1 * torch.randn(2, 3, 4, device=dev)
    D, H, W = 12, 16, 20
    lin = [torch.linspace(-1, 1, s, device=dev) for s in (D, H, W)]
    g = torch.stack(torch.meshgrid(*lin, indexing="ij"), -1).reshape(1, -1, 3).repeat(2, 1, 1)
    gh = torch.cat([g, torch.ones_like(g[..., :1])], -1)
    ref = torch.bmm(mat, gh.permute(0, 2, 1)).permute(0, 2, 1).reshape(2, D, H, W, 3).flip(-1)
    got = ops.flow_field_affine(mat, (D, H, W))
    ok &= _report("flow_field_affine", got, ref, 2e-6)
    # fused warp + loss (affine / grid)
    for C in (1, 5, 18):
        mov = torch.rand(2, C, D, H, W, device=dev)
        fix = torch.rand(2, C, D, H, W, device=dev)
        ref_w = F.grid_sample(mov, ref, mode="bilinear", padding_mode="border", align_corners=False)
        out, sums = ops.warp_loss(mov, fix, mat34=mat)
        ok &= _report(f"warp_loss affine C={C} warped", out, ref_w, 2e-5)
        rs = torch.stack([((ref_w - fix) ** 2).flatten(2).sum(-1), (ref_w * fix).flatten(2).sum(-1),
                          (ref_w ** 2).flatten(2).sum(-1), (fix ** 2).flatten(2).sum(-1)], -1)
        ok &= _report(f"warp_loss affine C={C} sums", sums, rs, 1e-5)
        out2, sums2 = ops.warp_loss(mov, fix, grid=ref, mode="nearest")
        ref_n = F.grid_sample(mov, ref, mode="nearest", padding_mode="border", align_corners=False)
        ok &= _report(f"warp_loss grid nearest C={C}", out2, ref_n, 0)
    return ok


def run_com():
    import torch
    from keymorph_b200 import ops
    torch.manual_seed(0)
    ok = True
    for shape in [(2, 4, 16, 16, 16), (1, 3, 9, 10, 11), (1, 2, 64, 64, 64)]:
        h = torch.randn(*shape, device="cuda")
        v = torch.relu(h)
        N, K, D, H, W = shape
        lz, ly, lx = (torch.linspace(0, 1, s, device="cuda") for s in (D, H, W))
        mx, my, mz = v.sum((2, 3)), v.sum((2, 4)), v.sum((3, 4))
        cx = (lx * mx).sum(-1) / (mx.sum(-1) + 1e-8)
        cy = (ly * my).sum(-1) / (my.sum(-1) + 1e-8)
        cz = (lz * mz).sum(-1) / (mz.sum(-1) + 1e-8)
        ref = torch.stack([cz, cy, cx], -1) * 2 - 1
        got, mass = ops.com3d(h, ij=True, return_mass=True)
        ok &= _report(f"com3d {shape}", got, ref, 1e-5)
        ok &= _report(f"com3d mass {shape}", mass, v.flatten(2).sum(-1), 1e-5)
    return ok


def run_fit():
    import torch
    from keymorph_b200 import ops
    torch.manual_seed(0)
    ok = True
    N, K = 3, 64
    x = torch.rand(N, K, 3, device="cuda", dtype=torch.float64) * 2 - 1
    A = torch.eye(3, device="cuda", dtype=torch.float64) + 0.2 * torch.randn(N, 3, 3, device="cuda", dtype=torch.float64)
    t = 0.1 * torch.randn(N, 1, 3, device="cuda", dtype=torch.float64)
    y = x @ A.transpose(1, 2) + t + 0.01 * torch.randn(N, K, 3, device="cuda", dtype=torch.float64)
    for w in (None, torch.rand(N, K, device="cuda", dtype=torch.float64)):
        X = torch.cat([x, torch.ones_like(x[..., :1])], -1).transpose(1, 2)  # (N,4,K)
        Y = y.transpose(1, 2)
        Wm = torch.diag_embed(w) if w is not None else torch.eye(K, device="cuda", dtype=torch.float64)[None]
        ref = Y @ Wm @ X.transpose(1, 2) @ torch.linalg.inv(X @ Wm @ X.transpose(1, 2))
        A44, Ainv, st = ops.fit_affine(x.float(), y.float(), None if w is None else w.float())
        ok &= _report(f"fit_affine w={'yes' if w is not None else 'no'}", A44[:, :3], ref, 1e-5)
        ok &= _report("fit_affine inverse", torch.bmm(A44.double(), Ainv.double()),
                      torch.eye(4, device="cuda", dtype=torch.float64)[None].repeat(N, 1, 1), 1e-5)
        print("      status", st.tolist())
    # rigid: exact rotation recovery
    ang = torch.tensor([0.3, -1.0, 2.0], dtype=torch.float64)
    Rz = torch.stack([torch.stack([ang.cos(), -ang.sin(), torch.zeros(3, dtype=torch.float64)], -1),
                      torch.stack([ang.sin(), ang.cos(), torch.zeros(3, dtype=torch.float64)], -1),
                      torch.tensor([[0, 0, 1.0]] * 3, dtype=torch.float64)], 1).cuda()
    y = x @ Rz.transpose(1, 2) + t
    A44, Ainv, st = ops.fit_rigid(x.float(), y.float())
    ok &= _report("fit_rigid R", A44[:, :3, :3], Rz, 1e-5)
    ok &= _report("fit_rigid T", A44[:, :3, 3], t[:, 0], 1e-5)
    # reference KATs (test/test.py:259-413): collinear and coplanar point sets
    p1 = torch.tensor([[0, 0, 0], [0, 0, 0.1], [0, 0, 0.2], [0, 0, 0.3]], device="cuda").float()[None]
    p2 = torch.tensor([[0, 0, 0.1], [0, 0, 0.2], [0, 0, 0.3], [0, 0, 0.4]], device="cuda").float()[None]
    A44, Ainv, st = ops.fit_rigid(p1, p2)
    true = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0.1], [0, 0, 0, 1.0]], device="cuda")[None]
    ok &= _report("rigid KAT collinear translation", A44, true, 1e-5)
    p1 = torch.tensor([[0.1, -0.1, 0.1], [0.3, -0.2, 0.2], [0.5, -0.3, 0.3], [0.7, -0.4, 0.4]], device="cuda")[None]
    p2 = torch.tensor([[0.3, 0, 0], [0.5, -0.1, 0.1], [0.7, -0.2, 0.2], [0.9, -0.3, 0.3]], device="cuda")[None]
    A44, Ainv, st = ops.fit_rigid(p1, p2)
    true = torch.tensor([[1, 0, 0, 0.2], [0, 1, 0, 0.1], [0, 0, 1, -0.1], [0, 0, 0, 1.0]], device="cuda")[None]
    ok &= _report("rigid KAT collinear 3d translation", A44, true, 1e-5)
    p1 = torch.tensor([[1, 0, 0], [0, -1, 0], [-1, 0, 0], [0, 1, 0]], device="cuda").float()[None]
    p2 = 0.5 * torch.tensor([[0, -1, 0], [-1, 0, 0], [0, 1, 0], [1, 0, 0]], device="cuda").float()[None]
    A44, Ainv, st = ops.fit_rigid(p1, p2)
    true = torch.tensor([[0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]], device="cuda")[None]
    ok &= _report("rigid KAT coplanar rotation", A44, true, 1e-5)
    # singular affine (all z = 0): status must be set
    A44, Ainv, st = ops.fit_affine(p1, p2)
    print("      singular affine status", st.tolist(), "(expect nonzero)")
    ok &= bool(st.item() != 0)
    return ok


def _tps_ref64(c_src, c_dst, lam, pts):
    import torch
    c, t = c_src.double(), c_dst.double()
    N, K, _ = c.shape
    d = torch.sqrt(((c[:, :, None] - c[:, None]) ** 2).sum(-1) + 1e-6)
    U = d ** 2 * torch.log(d + 1e-6)
    A = torch.zeros(N, K + 4, K + 4, dtype=torch.float64, device=c.device)
    A[:, :K, :K] = U + lam.double().view(N, 1, 1) * torch.eye(K, dtype=torch.float64, device=c.device)
    P = torch.cat([torch.ones(N, K, 1, dtype=torch.float64, device=c.device), c], -1)
    A[:, :K, K:] = P
    A[:, K:, :K] = P.transpose(1, 2)
    v = torch.zeros(N, K + 4, 3, dtype=torch.float64, device=c.device)
    v[:, :K] = t
    theta = torch.linalg.solve(A, v)
    p = pts.double()
    dd = torch.sqrt(((c[:, :, None] - p[:, None]) ** 2).sum(-1) + 1e-6)
    Up = dd ** 2 * torch.log(dd + 1e-6)
    out = torch.bmm(Up.transpose(1, 2), theta[:, :K]) + \
        torch.bmm(torch.cat([torch.ones_like(p[..., :1]), p], -1), theta[:, K:])
    return theta, out


def run_tps():
    import time
    import torch
    from keymorph_b200 import _lib, ops
    torch.manual_seed(0)
    ok = True
    for K, spread, lam in [(16, 0.6, 0.0), (128, 0.6, 0.0), (128, 0.6, 1.0), (512, 0.6, 0.0),
                           (512, 0.05, 0.0), (512, 0.6, 10.0)]:
        N = 2
        c_src = (torch.rand(N, K, 3, device="cuda") * 2 - 1) * spread if spread > 0.1 else \
            torch.randn(N, K, 3, device="cuda") * spread
        c_dst = c_src + 0.05 * torch.randn(N, K, 3, device="cuda")
        lmb = torch.full((N,), lam, device="cuda")
        pts = torch.rand(N, 4096, 3, device="cuda") * 2 - 1
        theta64, out64 = _tps_ref64(c_src, c_dst, lmb, pts)
        theta, st = ops.tps_fit(c_src, c_dst, lmb)
        torch.cuda.synchronize()
        t0 = time.time()
        theta, st = ops.tps_fit(c_src, c_dst, lmb)
        torch.cuda.synchronize()
        dt = time.time() - t0
        got = ops.points_transform_tps(c_src, theta, pts)
        e_theta = (theta.double() - theta64).abs().max().item() / theta64.abs().max().item()
        e_pts = (got.double() - out64).abs().max().item()
        # the reference's own fp32 path (torch ops, fp32) for comparison
        c, t = c_src, c_dst
        d = torch.sqrt(((c[:, :, None] - c[:, None]) ** 2).sum(-1) + 1e-6)
        U = d ** 2 * torch.log(d + 1e-6)
        A = torch.zeros(N, K + 4, K + 4, device="cuda")
        A[:, :K, :K] = U + lmb.view(N, 1, 1) * torch.eye(K, device="cuda")
        P = torch.cat([torch.ones(N, K, 1, device="cuda"), c], -1)
        A[:, :K, K:] = P
        A[:, K:, :K] = P.transpose(1, 2)
        v = torch.zeros(N, K + 4, 3, device="cuda")
        v[:, :K] = t
        th32 = torch.linalg.solve(A.cpu(), v.cpu()).cuda()
        dd = torch.sqrt(((c[:, :, None] - pts[:, None]) ** 2).sum(-1) + 1e-6)
        Up = dd ** 2 * torch.log(dd + 1e-6)
        out32 = torch.bmm(Up.transpose(1, 2), th32[:, :K]) + \
            torch.bmm(torch.cat([torch.ones_like(pts[..., :1]), pts], -1), th32[:, K:])
        e_ref = (out32.double() - out64).abs().max().item()
        good = e_pts <= 2 * e_ref + 1e-5
        ok &= good
        print(f"  [{'ok' if good else 'FAIL'}] tps K={K} spread={spread} lam={lam}: rel theta err {e_theta:.2e}, "
              f"points err vs fp64 {e_pts:.2e} (reference-fp32 path err {e_ref:.2e}), fit {dt*1e3:.2f} ms, "
              f"status {st.tolist()}", flush=True)
        # dense flow field: fast vs accurate radial basis
        D = 24
        lin = torch.linspace(-1, 1, D, device="cuda")
        g = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3).repeat(N, 1, 1)
        _, g64 = _tps_ref64(c_src, c_dst, lmb, g)
        ref_grid = g64.reshape(N, D, D, D, 3).flip(-1)
        for fast in (1, 0):
            _lib.call("km_set_option", _lib.KM_OPT_TPS_FAST, fast)
            gg = ops.flow_field_tps(c_src, theta, (D, D, D))
            e = (gg.double() - ref_grid).abs().max().item()
            print(f"      flow_field_tps fast={fast}: max err vs fp64 {e:.2e}")
        _lib.call("km_set_option", _lib.KM_OPT_TPS_FAST, 1)
    return ok


def run_loss():
    import torch
    from keymorph_b200 import ops
    torch.manual_seed(0)
    ok = True
    for (N, C, S) in [(1, 1, 24), (2, 14, 20), (1, 33, 12)]:
        p = torch.rand(N, C, S, S, S, device="cuda")
        t = torch.rand(N, C, S, S, S, device="cuda")
        s = ops.pair_stats(p, t)
        rs = torch.stack([((p - t) ** 2).flatten(2).sum(-1), (p * t).flatten(2).sum(-1),
                          (p ** 2).flatten(2).sum(-1), (t ** 2).flatten(2).sum(-1)], -1)
        ok &= _report(f"pair_stats soft N={N} C={C}", s, rs, 1e-5)
        lab = ops.argmax_channels(p)
        ref_lab = torch.argmax(p, 1)
        nb = (lab.long() != ref_lab).sum().item()
        print(f"  [{'ok' if nb == 0 else 'FAIL'}] argmax labels mismatches: {nb}")
        ok &= nb == 0
        sh = ops.pair_stats(p, t, hard=True)
        oh = torch.zeros_like(p).scatter_(1, ref_lab[:, None], 1.0)
        rh = torch.stack([((oh - t) ** 2).flatten(2).sum(-1), (oh * t).flatten(2).sum(-1),
                          (oh ** 2).flatten(2).sum(-1), (t ** 2).flatten(2).sum(-1)], -1)
        ok &= _report(f"pair_stats hard N={N} C={C}", sh, rh, 1e-5)
    return ok


def run_misc():
    import torch
    import torch.nn.functional as F
    from keymorph_b200 import ops
    torch.manual_seed(0)
    ok = True
    dev = "cuda"
    # stem conv with input GroupNorm(1 group)
    for (N, Cout, D, H, W) in [(1, 16, 16, 16, 32), (2, 32, 9, 11, 35)]:
        x = torch.rand(N, 1, D, H, W, device=dev)
        w = torch.randn(Cout, 1, 3, 3, 3, device=dev) * 0.2
        gam, bet = torch.rand(1, device=dev) + 0.5, torch.randn(1, device=dev) * 0.1
        st = ops.volume_stats(x)
        sc, sh = ops.norm_finalize(st, D * H * W, gam, bet, 1)
        out, stats = ops.conv3d_stem(x, w, None, sc.reshape(-1), sh.reshape(-1), relu_pre=True)
        ref = F.relu(F.conv3d(F.group_norm(x, 1, gam, bet, 1e-5), w, padding=1))
        ok &= _report(f"stem conv Cout={Cout}", ops.ndhwc_to_ncdhw(out), ref, 1e-2)
        s = stats.double().sum(0)
        ok &= _report("stem stats sum", s[..., 0], ref.flatten(2).sum(-1), 5e-3)
    # norm finalize/apply (GroupNorm 8) + pool + concat
    N, C, D, H, W = 2, 32, 8, 8, 16
    xr = torch.randn(N, C, D, H, W, device=dev)
    xb = ops.ncdhw_to_ndhwc(xr)
    xq = ops.ndhwc_to_ncdhw(xb)
    ok &= _report("layout roundtrip", xq, xr, 1e-2)
    gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    st = ops.channel_stats(xb)
    sc, sh = ops.norm_finalize(st, D * H * W, gam, bet, 8)
    y = ops.norm_apply(xb, sc, sh)
    ref = F.group_norm(xq, 8, gam, bet, 1e-5)
    ok &= _report("groupnorm apply", ops.ndhwc_to_ncdhw(y), ref, 2e-2)
    pooled, pst = ops.maxpool2_stats(xb)
    refp = F.max_pool3d(xq, 2)
    ok &= _report("maxpool2", ops.ndhwc_to_ncdhw(pooled), refp, 0)
    ok &= _report("maxpool2 stats", pst.double().sum(0)[..., 0], refp.flatten(2).sum(-1), 1e-4)
    ok &= _report("maxpool2 stats sq", pst.double().sum(0)[..., 1], (refp ** 2).flatten(2).sum(-1), 1e-4)
    # concat + upsample: src1 at half resolution with 16 channels
    x1r = torch.randn(N, 16, D // 2, H // 2, W // 2, device=dev)
    x1b = ops.ncdhw_to_ndhwc(x1r)
    x1q = ops.ndhwc_to_ncdhw(x1b)
    cat = torch.cat([xq, F.interpolate(x1q, size=(D, H, W), mode="nearest")], 1)
    g2, b2 = torch.rand(48, device=dev) + 0.5, torch.randn(48, device=dev)
    st1 = ops.channel_stats(x1b)
    sc, sh = ops.norm_finalize(st, D * H * W, g2, b2, 8, stats1=st1, count1=(D * H * W) // 8, rep1=8.0)
    y = ops.norm_apply(xb, sc, sh, src1=x1b)
    ref = F.group_norm(cat, 8, g2, b2, 1e-5)
    ok &= _report("groupnorm concat+upsample", ops.ndhwc_to_ncdhw(y), ref, 2e-2)
    # instance norm + relu + pool
    sc, sh = ops.norm_finalize(st, D * H * W, None, None, C)
    y = ops.norm_apply(xb, sc, sh, relu=True, pool=True)
    ref = F.max_pool3d(F.relu(F.instance_norm(xq)), 2)
    ok &= _report("instnorm+relu+pool", ops.ndhwc_to_ncdhw(y), ref, 2e-2)
    return ok


def _conv_case(N, Cin, Cout, D, H, W, taps, relu, bias):
    import torch
    import torch.nn.functional as F
    from keymorph_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(Cin * 1000 + Cout + D)
    x = torch.randn(N, Cin, D, H, W, device=dev, generator=g)
    k = 3 if taps == 27 else 1
    w = torch.randn(Cout, Cin, k, k, k, device=dev, generator=g) / (Cin * taps) ** 0.5
    b = torch.randn(Cout, device=dev, generator=g) if bias else None
    xb = ops.ncdhw_to_ndhwc(x)
    wp = ops.pack_weights(w)
    xq = ops.ndhwc_to_ncdhw(xb)
    wq = w.to(ops.act_dtype()).float()
    out, stats, _ = ops.conv3d_tc(xb, wp, bias=b, relu=relu, want_stats=True)
    torch.cuda.synchronize()
    ref = F.conv3d(xq, wq, b, padding=k // 2)
    if relu:
        ref = F.relu(ref)
    got = ops.ndhwc_to_ncdhw(out)
    ok = _report(f"conv3d_tc N={N} {Cin}->{Cout} {D}x{H}x{W} taps={taps} relu={relu} bias={bias}", got, ref, 1.5e-2)
    refq = ref.to(ops.act_dtype()).float()
    s = stats.double().sum(0)
    ok &= _report("   stats sum", s[..., 0], refq.flatten(2).sum(-1), 2e-3)
    ok &= _report("   stats sumsq", s[..., 1], (refq ** 2).flatten(2).sum(-1), 2e-3)
    return ok


def run_conv():
    from keymorph_b200 import _lib
    ok = True
    cases = [
        (1, 64, 64, 8, 8, 8, 27, True, False),     # SW128, single tile row
        (1, 32, 32, 8, 8, 16, 27, True, False),    # SW64
        (1, 16, 32, 8, 16, 16, 27, True, False),   # SW32
        (2, 64, 128, 16, 16, 16, 27, False, True),
        (1, 128, 256, 8, 8, 8, 27, True, False),
        (1, 384, 128, 8, 8, 8, 27, True, False),
        (1, 192, 64, 16, 16, 16, 27, True, False),
        (1, 96, 32, 8, 8, 32, 27, True, False),
        (1, 64, 256, 16, 16, 16, 1, False, True),
        (1, 256, 512, 8, 8, 8, 27, False, True),   # two channel blocks
        (1, 64, 64, 10, 12, 20, 27, True, False),  # clipped bricks
        (1, 32, 64, 64, 64, 64, 27, True, False),
    ]
    cases += [(1, 64, 64, 4, 4, 4, 27, True, False), (2, 128, 128, 4, 4, 4, 27, True, False),
              (1, 64, 64, 8, 24, 40, 27, True, False), (1, 32, 64, 5, 9, 17, 27, True, False)]
    for force, axis in ((0, 2), (0, 1), (1, 2)):
        _lib.call("km_set_option", _lib.KM_OPT_CONV_FORCE_GENERIC, force)
        _lib.call("km_set_option", _lib.KM_OPT_CONV_HALO_AXIS, axis)
        print(f"  -- force generic path = {force}, halo axis = {axis}")
        for c in cases:
            try:
                ok &= _conv_case(*c)
            except Exception as e:  # noqa: BLE001
                print(f"  [FAIL] conv case {c}: {type(e).__name__}: {e}", flush=True)
                return False
    _lib.call("km_set_option", _lib.KM_OPT_CONV_FORCE_GENERIC, 0)
    return ok


def run_conv_big():
    import time
    import torch
    from keymorph_b200 import ops
    ok = _conv_case(1, 16, 32, 128, 128, 128, 27, True, False)
    ok &= _conv_case(1, 64, 64, 128, 128, 128, 27, True, False)
    # timing of the heavy layers
    for (Cin, Cout, S) in [(16, 32, 256), (32, 32, 128), (32, 64, 128), (64, 64, 64), (64, 128, 64),
                           (128, 128, 32), (128, 256, 32), (384, 128, 64), (128, 128, 64),
                           (192, 64, 128), (64, 64, 128)]:
        x = torch.randn(1, S, S, S, Cin, device="cuda").to(ops.act_dtype())
        wp = (torch.randn(27, Cout, Cin, device="cuda") / (27 * Cin) ** 0.5).to(ops.act_dtype())
        for _ in range(2):
            ops.conv3d_tc(x, wp, relu=True, want_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.conv3d_tc(x, wp, relu=True, want_stats=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fl = 2.0 * 27 * Cin * Cout * S ** 3
        print(f"  conv {Cin}->{Cout} @{S}^3: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def run_convtime():
    """Per-layer timing of the bench workload's convolutions (N=2), with the data-path options."""
    import torch
    from keymorph_b200 import _lib, ops
    layers = [(16, 32, 256, 27), (32, 32, 128, 27), (32, 64, 128, 27), (64, 64, 64, 27), (64, 128, 64, 27),
              (128, 128, 32, 27), (128, 256, 32, 27), (384, 128, 64, 27), (128, 128, 64, 27),
              (192, 64, 128, 27), (64, 64, 128, 27), (64, 256, 128, 1)]
    for label, opts in (("default (halo axis y)", {}), ("halo axis x", {_lib.KM_OPT_CONV_HALO_AXIS: 1}),
                        ("streamed weights", {_lib.KM_OPT_CONV_NO_RESIDENT_WEIGHTS: 1})):
        _lib.call("km_set_option", _lib.KM_OPT_CONV_HALO_AXIS, 2)
        _lib.call("km_set_option", _lib.KM_OPT_CONV_MAX_BRICKS, 4)
        for k, v in opts.items():
            _lib.call("km_set_option", k, v)
        print(f"  -- {label}")
        total = 0.0
        for (Cin, Cout, S, taps) in layers:
            x = torch.randn(2, S, S, S, Cin, device="cuda").to(ops.act_dtype())
            wp = (torch.randn(taps, Cout, Cin, device="cuda") / (taps * Cin) ** 0.5).to(ops.act_dtype())
            bias = torch.zeros(Cout, device="cuda") if taps == 1 else None
            kw = dict(want_com=True, store=False, bias=bias) if taps == 1 else dict(relu=True, want_stats=True)
            for _ in range(2):
                ops.conv3d_tc(x, wp, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.conv3d_tc(x, wp, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            total += ms
            fl = 2.0 * taps * Cin * Cout * S ** 3 * 2
            print(f"  conv {Cin}->{Cout} @{S}^3 x2 taps={taps}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        print(f"  total {total:.3f} ms")
        for k in opts:
            _lib.call("km_set_option", k, 0)
    return True


def run_convcom():
    import torch
    import torch.nn.functional as F
    from keymorph_b200 import ops
    ok = True
    for (N, Cin, K, D, H, W) in [(1, 64, 64, 16, 16, 16), (2, 64, 128, 8, 16, 32), (1, 32, 512, 16, 16, 16)]:
        g = torch.Generator(device="cuda").manual_seed(K)
        x = torch.randn(N, Cin, D, H, W, device="cuda", generator=g)
        w = torch.randn(K, Cin, 1, 1, 1, device="cuda", generator=g) / Cin ** 0.5
        b = torch.randn(K, device="cuda", generator=g) * 0.1
        xb, wp = ops.ncdhw_to_ndhwc(x), ops.pack_weights(w)
        xq, wq = ops.ndhwc_to_ncdhw(xb), w.to(ops.act_dtype()).float()
        heat = F.conv3d(xq, wq, b)
        ref, refm = ops.com3d(heat, ij=True, return_mass=True)
        for store in (False, True):
            out, _, com = ops.conv3d_tc(xb, wp, bias=b, want_com=True, store=store)
            pts, mass = ops.com_finalize(com, return_mass=True)
            ok &= _report(f"conv1x1+CoM K={K} store={store} points", pts, ref, 1e-4)
            ok &= _report(f"conv1x1+CoM K={K} store={store} mass", mass, refm, 1e-4)
            if store:
                ok &= _report("   heat", ops.ndhwc_to_ncdhw(out), heat, 1.5e-2)
    return ok


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--group":
        fn = globals()["run_" + sys.argv[2]]
        ok = fn()
        print(f"group {sys.argv[2]}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    groups = sys.argv[1:] or GROUPS
    results = {}
    for g in groups:
        print(f"== {g} ==", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], timeout=420)
            results[g] = r.returncode
        except subprocess.TimeoutExpired:
            print(f"group {g}: TIMEOUT", flush=True)
            results[g] = "timeout"
    print("summary:", results, flush=True)


if __name__ == "__main__":
    main()
