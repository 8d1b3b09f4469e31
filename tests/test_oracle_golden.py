"""The CPU oracle (oracle/keymorph_oracle.py) against (a) the reference's own known-answer tests
(test/test.py, restated here) and (b) golden vectors produced by the reference itself
(oracle/gen_golden.py).  No GPU involved."""
import os

import numpy as np
import pytest
import torch
from scipy import ndimage
from torch.testing import assert_close

from oracle import keymorph_oracle as O


# ---------------------------------------------------------------- reference KATs: CenterOfMass3d
def _blob(shape, at, sigma=5):
    img = np.zeros(shape)
    img[at] = 1
    return torch.tensor(ndimage.gaussian_filter(img, sigma)).float()


def test_com_kats():
    """test/test.py:117-253."""
    pt = torch.zeros(3, 3, 3)
    pt[1, 1, 1] = 1
    assert_close(O.center_of_mass3d(pt.view(1, 1, 3, 3, 3), ij=False), torch.zeros(1, 1, 3))
    three = torch.zeros(3, 3, 3)
    three[0, 0, 0] = three[1, 1, 1] = three[2, 2, 2] = 1
    assert_close(O.center_of_mass3d(three.view(1, 1, 3, 3, 3), ij=False), torch.zeros(1, 1, 3))
    assert_close(O.center_of_mass3d(_blob((101, 101, 101), (50, 50, 50))[None, None], ij=False),
                 torch.zeros(1, 1, 3))
    assert_close(O.center_of_mass3d(_blob((101, 51, 51), (50, 25, 25))[None, None], ij=False),
                 torch.zeros(1, 1, 3))
    off = _blob((101, 101, 101), (50, 25, 25))[None, None]
    assert_close(O.center_of_mass3d(off, ij=False), torch.tensor([[[-0.5, -0.5, 0.0]]]))
    two = torch.stack([_blob((101, 101, 101), (50, 25, 25)), _blob((101, 101, 101), (25, 50, 50))])[:, None]
    assert_close(O.center_of_mass3d(two, ij=False), torch.tensor([[[-0.5, -0.5, 0]], [[0, 0, -0.5]]]))
    assert_close(O.center_of_mass3d(two, ij=True), torch.tensor([[[0, -0.5, -0.5]], [[-0.5, 0, 0]]]))


# ---------------------------------------------------------------- reference KATs: rigid / affine
RIGID_KATS = [
    # (points_m, points_f, weights, expected transform_matrix)  test/test.py:259-413
    ([[0, 0, 0], [0, 0, 0.1], [0, 0, 0.2], [0, 0, 0.3]], [[0, 0, 0.1], [0, 0, 0.2], [0, 0, 0.3], [0, 0, 0.4]],
     None, [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0.1], [0, 0, 0, 1]]),
    ([[0.1, -0.1, 0.1], [0.3, -0.2, 0.2], [0.5, -0.3, 0.3], [0.7, -0.4, 0.4]],
     [[0.3, 0, 0], [0.5, -0.1, 0.1], [0.7, -0.2, 0.2], [0.9, -0.3, 0.3]],
     None, [[1, 0, 0, 0.2], [0, 1, 0, 0.1], [0, 0, 1, -0.1], [0, 0, 0, 1]]),
    ([[1, 0, 0], [0, -1, 0], [-1, 0, 0], [0, 1, 0]], [[0, -1, 0], [-1, 0, 0], [0, 1, 0], [1, 0, 0]],
     None, [[0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]),
    ([[1, 0, 0], [0, -1, 0], [-1, 0, 0], [0, 1, 0]], [[0, -0.5, 0], [-0.5, 0, 0], [0, 0.5, 0], [0.5, 0, 0]],
     None, [[0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]),
    ([[1, 0, 0], [0, -1, 0], [-1, 0, 0], [0, 1, 0]], [[0, -0.5, 0], [-0.5, 0, 0], [0, 0.5, 0], [0.5, 0, 0]],
     [1, 1, 1, 1], [[0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]),
]


@pytest.mark.parametrize("pm,pf,w,expected", RIGID_KATS)
def test_rigid_kats(pm, pf, w, expected):
    pm, pf = torch.tensor(pm).float()[None], torch.tensor(pf).float()[None]
    w = None if w is None else torch.tensor(w).float()[None]
    tm, inv = O.aligner_matrices(pm, pf, w, "rigid")
    assert_close(tm, torch.tensor(expected).float()[None])


def test_rigid_forward_inverse_symmetry():
    """test/test.py:279-299."""
    a = torch.tensor(RIGID_KATS[0][0]).float()[None]
    b = torch.tensor(RIGID_KATS[0][1]).float()[None]
    f12, i12 = O.aligner_matrices(a, b, None, "rigid")
    f21, i21 = O.aligner_matrices(b, a, None, "rigid")
    assert_close(f12, i21)
    assert_close(f21, i12)


def test_affine_singular_raises():
    """test/test.py:462-480 (test_affine_2 errors in the reference: all z == 0)."""
    pm = torch.tensor([[1, 0, 0], [0, -1, 0], [-1, 0, 0], [0, 1, 0]]).float()[None]
    pf = torch.tensor([[0, -1, 0], [-1, 0, 0], [0, 1, 0], [1, 0, 0]]).float()[None]
    with pytest.raises(Exception):
        O.aligner_matrices(pm, pf, None, "affine")


# ---------------------------------------------------------------- golden vectors from the reference
def test_golden_aligners(golden):
    g = golden("aligners")
    pm, pf, w = g["points_m"], g["points_f"], g["w"]
    shape = tuple(int(s) for s in g["shape"][2:])
    for kind in ("affine", "rigid"):
        for tag, ww in (("", None), ("_w", w)):
            tm, inv = O.aligner_matrices(pm, pf, ww, kind)
            assert_close(tm, g[f"{kind}{tag}_matrix"], rtol=0, atol=0)
            assert_close(inv, g[f"{kind}{tag}_inverse"], rtol=0, atol=0)
            assert_close(O.affine_flow_field(inv, shape), g[f"{kind}{tag}_grid"], rtol=0, atol=0)
            assert_close(O.transform_points(tm, pm), g[f"{kind}{tag}_points_a"], rtol=0, atol=0)
            assert_close(O.transform_points(inv, pf), g[f"{kind}{tag}_points_inv"], rtol=0, atol=0)


def test_golden_tps(golden):
    g = golden("tps")
    pm, pf, w = g["points_m"], g["points_f"], g["w"]
    shape = tuple(int(s) for s in g["shape"][2:])
    for lam, tag, ww in ((0.0, "lam0", None), (0.1, "lam0.1", None), (10.0, "lam10", None),
                         (0.1, "lam0.1_w", w)):
        lmbda = torch.tensor([lam])
        assert_close(O.tps_fit(pf, pm, lmbda, ww), g[f"{tag}_inverse_theta"], rtol=0, atol=0)
        assert_close(O.tps_fit(pm, pf, lmbda, ww), g[f"{tag}_theta"], rtol=0, atol=0)
        # the reference evaluates 4 sub-grids; chunking does not change per-voxel arithmetic but
        # may change the BLAS blocking, hence a tolerance at rounding level
        assert_close(O.tps_flow_field(pm, pf, lmbda, shape, ww), g[f"{tag}_grid"], rtol=1e-5, atol=2e-5)
        assert_close(O.tps_forward_points(pm, pf, lmbda, pm, ww), g[f"{tag}_points_a"], rtol=1e-5, atol=1e-5)


def test_golden_warp_and_losses(golden):
    g = golden("warp_loss")
    for mode in ("bilinear", "nearest"):
        assert_close(O.align_img(g["grid"], g["x"], mode), g[mode], rtol=0, atol=0)
        mine = torch.from_numpy(O.grid_sample3d_numpy(g["x"].numpy(), g["grid"].numpy(), mode))
        if mode == "nearest":
            assert torch.equal(mine, g[mode])          # pure gather: bit exact
        else:
            assert_close(mine, g[mode], rtol=1e-6, atol=1e-6)
    p, t = g["seg_pred"], g["seg_target"]
    assert_close(O.mse_loss(p, t), g["mse"], rtol=0, atol=0)
    for hard in (0, 1):
        for ign in (0, 1):
            assert_close(O.dice_loss(p, t, bool(hard), bool(ign)), g[f"dice_h{hard}_i{ign}"], rtol=0, atol=0)
            assert_close(O.dice_loss(p, t, bool(hard), bool(ign), True), g[f"dice_regions_h{hard}_i{ign}"],
                         rtol=0, atol=0)


def test_golden_com(golden):
    g = golden("com")
    assert_close(O.center_of_mass3d(g["heat"], ij=True), g["points_ij"], rtol=0, atol=0)
    assert_close(O.center_of_mass3d(g["heat"], ij=False), g["points_xy"], rtol=0, atol=0)


def test_golden_augment(golden):
    g = golden("augment")
    params = tuple(float(v) for v in g["params"])
    img, seg = O.affine_augment(g["img"], params, seg=g["seg"])
    assert_close(img, g["img_aug"], rtol=1e-6, atol=1e-6)
    assert (seg != g["seg_aug"]).float().mean() < 1e-3   # nearest: a coordinate may sit on a tie


def test_golden_augment_anisotropic(golden):
    """oracle AND the product's host-side matrix builder against the reference's AffineDeformation3d."""
    from keymorph_b200.augmentation import AffineDeformation3d
    g = golden("augment_aniso")
    M = O.affine_matrix_3d_params(g["scale"], g["offset"], g["theta"], g["shear"])
    assert_close(M, g["matrix"], rtol=1e-6, atol=1e-6)
    params = (g["scale"], g["offset"], g["theta"], g["shear"])
    Mp = AffineDeformation3d(device="cpu").build_affine_matrix(1, params)
    assert_close(Mp, g["matrix"], rtol=1e-6, atol=1e-6)
    Mb = AffineDeformation3d(device="cpu").build_affine_matrix(3, params)
    assert Mb.shape == (3, 4, 4) and torch.equal(Mb[2], Mp[0])
    img, seg, pts = O.deform(g["matrix"], g["img"], g["seg"], g["points"])
    assert_close(img, g["img_aug"], rtol=1e-5, atol=1e-5)
    assert (seg != g["seg_aug"]).float().mean() < 5e-3   # nearest: a coordinate may sit on a tie
    assert_close(pts, g["points_aug"], rtol=1e-5, atol=1e-6)


def test_golden_group_metrics(golden):
    """Hausdorff / fast_dice / pairwise group metrics of the oracle against the reference's own values
    (which come from scipy.ndimage's erosion and exact EDT)."""
    g = golden("group_metrics")
    segs, C = g["segs"], g["segs"].shape[1]
    hard = torch.nn.functional.one_hot(segs.argmax(1), C).permute(0, 4, 1, 2, 3).float()
    assert abs(O.hausdorff_distance(hard[0:1], hard[1:2]) - float(g["hausd_01"])) < 1e-9
    assert abs(O.hausdorff_distance(hard[0:2], hard[1:3]) - float(g["hausd_batch"])) < 1e-9
    blobs = [O.hausdorff_distance(hard[i:i + 1, k:], hard[j:j + 1, k:])
             for (i, j) in ((0, 1), (0, 2), (1, 2)) for k in (1, 2, 3)]
    np.testing.assert_allclose(blobs, g["hausd_blobs"].numpy(), rtol=1e-12)
    assert abs(O.fast_dice(segs[0:1].numpy(), segs[1:2].numpy()) - float(g["fast_dice_01"])) < 1e-9
    assert_close(O.avg_pairwise(g["imgs"], O.mse_loss), g["mse_pairwise"], rtol=1e-6, atol=0)
    assert_close(O.avg_pairwise(segs, O.dice_loss), g["softdice_pairwise"], rtol=1e-6, atol=0)
    assert_close(O.avg_pairwise(segs, lambda a, b: O.dice_loss(a, b, hard=True)), g["harddice_pairwise"],
                 rtol=1e-6, atol=0)
    assert abs(O.avg_pairwise(hard, O.hausdorff_distance) - float(g["multi_hausd"])) < 1e-9
    assert abs(O.avg_pairwise(hard.numpy(), O.fast_dice) - float(g["multi_dice"])) < 1e-9
    jd = np.mean([O.jdstd(g["grids"][i:i + 1].permute(0, 4, 1, 2, 3)) for i in range(3)])
    assert abs(jd - float(g["avg_jdstd"])) < 1e-7


@pytest.mark.parametrize("shape,sampling", [((10, 12, 14), (1.25, 1.25, 10.0)), ((7, 20, 9), (1.0, 2.0, 0.5)),
                                            ((16, 16, 16), (1.0, 1.0, 1.0))])
def test_oracle_hausdorff_vs_scipy_and_properties(shape, sampling):
    """The scipy-free restatement against the reference's own recipe on its dependency (binary_erosion +
    distance_transform_edt, keymorph/loss_ops.py:121-141), plus symmetry and d(A, A) = 0."""
    gen = torch.Generator().manual_seed(sum(shape))
    a = (torch.rand(shape, generator=gen) > 0.6).numpy()
    b = (torch.rand(shape, generator=gen) > 0.7).numpy()
    a[0, 0, 0] = b[-1, -1, -1] = True                      # never empty; corners exercise the border rule
    conn = ndimage.generate_binary_structure(3, 1)
    sa, sb = a & ~ndimage.binary_erosion(a, conn), b & ~ndimage.binary_erosion(b, conn)
    assert np.array_equal(O.surface_mask(a), sa) and np.array_equal(O.surface_mask(b), sb)
    dta, dtb = ndimage.distance_transform_edt(~sa, sampling), ndimage.distance_transform_edt(~sb, sampling)
    ref = max(dta[sb].max(), dtb[sa].max())
    got = O.hausdorff_distance(a[None, None], b[None, None], sampling)
    assert abs(got - ref) < 1e-9
    assert abs(O.hausdorff_distance(b[None, None], a[None, None], sampling) - got) < 1e-12
    assert O.hausdorff_distance(a[None, None], a[None, None], sampling) == 0.0


def test_eval_output_layout(tmp_path):
    """File names, dtypes and write-once behaviour of scripts/pairwise_register_eval.py:368-458 (host side)."""
    import json
    from keymorph_b200 import evaluation as E
    g = torch.Generator().manual_seed(1)
    img = torch.rand(1, 1, 4, 5, 6, generator=g)
    seg = torch.rand(1, 3, 4, 5, 6, generator=g)
    grid = torch.rand(1, 4, 5, 6, 3, generator=g)
    pts = torch.rand(1, 7, 3, generator=g)
    metrics = {"mse": 0.5, "harddiceroi": [0.1, 0.2]}
    w = E.save_pair_outputs(tmp_path, 3, "T1", "T2", "rot0", "affine", metrics, img, img * 2, img * 3, grid=grid,
                            seg_f=seg, seg_m=seg, seg_a=seg, points_f=pts, points_m=pts, points_a=pts,
                            points_weights=pts[..., 0])
    names = sorted(os.path.basename(p) for p in w)
    assert names == sorted([
        "metrics-rot0-affine.json", "img_f_3-T1.npy", "img_m_3-T2-rot0.npy", "img_a_3-T1-T2-rot0-affine.npy",
        "grid_3-T1-T2-rot0-affine.npy", "seg_f_3-T1.npy", "seg_m_3-T2-rot0.npy", "seg_a_3-T1-T2-rot0-affine.npy",
        "points_f_3-T1.npy", "points_m_3-T2-rot0.npy", "points_a_3-T1-T2-rot0-affine.npy",
        "points_weights_3-T1-T2-rot0-affine.npy"])
    assert json.load(open(tmp_path / "metrics-rot0-affine.json")) == metrics
    assert np.load(tmp_path / "img_a_3-T1-T2-rot0-affine.npy").shape == (1, 4, 5, 6)
    lab = np.load(tmp_path / "seg_a_3-T1-T2-rot0-affine.npy")
    assert lab.shape == (1, 4, 5, 6) and lab.dtype == np.int64
    assert np.array_equal(lab, seg.numpy().argmax(1))
    assert np.load(tmp_path / "grid_3-T1-T2-rot0-affine.npy").shape == (4, 5, 6, 3)
    assert np.load(tmp_path / "points_weights_3-T1-T2-rot0-affine.npy").shape == (7,)
    # a second alignment type of the same pair re-uses the fixed / moving files
    w2 = E.save_pair_outputs(tmp_path, 3, "T1", "T2", "rot0", "tps_0", metrics, img * 9, img, img, grid=None)
    assert sorted(os.path.basename(p) for p in w2) == ["img_a_3-T1-T2-rot0-tps_0.npy", "metrics-rot0-tps_0.json"]
    assert np.array_equal(np.load(tmp_path / "img_f_3-T1.npy"), img[0].numpy())


def _seeded(cls_name, **kw):
    import keymorph_b200 as kb
    torch.manual_seed(23)
    if cls_name == "trunc":
        return kb.TruncatedUNet3D(1, 16, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                                  num_levels=4, is_segmentation=False, conv_padding=1)
    if cls_name == "unet":
        return kb.UNet3D(1, 16, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                         is_segmentation=False, conv_padding=1)
    return kb.ConvNet(3, 1, 16, norm_type="instance")


def test_golden_backbones(golden):
    """Functional backbones of the oracle == the reference modules (same seeded weights)."""
    g = golden("truncunet_k16")
    sd = _seeded("trunc").state_dict()
    h32 = O.unet3d_forward(sd, O.gaussian_phantom(32, 1000), 4, 1)
    assert_close(h32, g["heat32"], rtol=1e-5, atol=1e-5)
    assert_close(O.center_of_mass3d(h32), g["points32"], rtol=1e-5, atol=1e-5)
    h64 = O.unet3d_forward(sd, O.gaussian_phantom(64, 1001), 4, 1)
    assert_close(O.center_of_mass3d(h64), g["points64"], rtol=1e-5, atol=1e-5)

    g = golden("unet_k16")
    h64 = O.unet3d_forward(_seeded("unet").state_dict(), O.gaussian_phantom(64, 1001), 4, 0)
    assert_close(O.center_of_mass3d(h64), g["points64"], rtol=1e-5, atol=1e-5)
    assert_close(h64[:, :, ::8, ::8, ::8], g["heat64_sub"], rtol=1e-4, atol=1e-4)

    g = golden("convnet_k16")
    h = O.convnet_forward(_seeded("conv").state_dict(), O.gaussian_phantom(128, 1002))
    assert_close(h, g["heat"], rtol=1e-4, atol=1e-4)
    assert_close(O.center_of_mass3d(h), g["points"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("power", [False, True])
def test_golden_forward(golden, power):
    g = golden("forward32_power" if power else "forward32")
    sd = _seeded("trunc").state_dict()
    types = ["rigid", "affine", "tps_10", "tps_0.1"]
    res = O.keymorph_forward("truncatedunet", sd, g["img_f"], g["img_m"], types,
                             weight_keypoints="power" if power else None)
    for t in types:
        r = res[t]
        assert_close(r["points_f"], g[f"{t}_points_f"], rtol=1e-5, atol=1e-5)
        assert_close(r["points_m"], g[f"{t}_points_m"], rtol=1e-5, atol=1e-5)
        tol = 2e-4 if t.startswith("tps") else 5e-5
        assert_close(r["points_a"], g[f"{t}_points_a"], rtol=tol, atol=tol)
        assert_close(r["grid"][:, ::2, ::2, ::2], g[f"{t}_grid"], rtol=tol, atol=tol)
        if "matrix" in r:
            assert_close(r["matrix"], g[f"{t}_matrix"], rtol=tol, atol=tol)
        if power:
            assert_close(r["points_weights"], g[f"{t}_weights"], rtol=1e-4, atol=1e-6)
        img_a = O.align_img(r["grid"], g["img_m"])
        assert_close(img_a[:, :, ::2, ::2, ::2], g[f"{t}_img_a"], rtol=1e-3, atol=1e-3)


def test_golden_groupwise(golden):
    g = golden("groupwise32")
    sd = _seeded("trunc").state_dict()
    subj = g["subjects"]
    pts = torch.cat([O.center_of_mass3d(O.unet3d_forward(sd, subj[i:i + 1], 4, 1)) for i in range(len(subj))])
    for t in ("rigid", "affine", "tps_1"):
        assert_close(pts, g[f"{t}_points_m"], rtol=1e-5, atol=1e-5)
        cur, mean = O.groupwise_points(pts, t, 3)
        assert_close(cur, g[f"{t}_points_a"], rtol=1e-4, atol=1e-4)
        for i in range(len(subj)):
            grid = O.register_points(mean, pts[i:i + 1], t, subj.shape[2:], None, False)["grid"]
            assert_close(grid[:, ::2, ::2, ::2], g[f"{t}_grid_{i}"], rtol=2e-4, atol=2e-4)


def test_golden_jacobian(golden):
    """Oracle restatement of loss_ops._jacobian_determinant / jdstd / jdlessthan0 (no scipy) against the
    reference's own outputs: the determinant field must agree to fp64 rounding."""
    import numpy as np
    g = golden("jacobian")
    for tag, key in (("norm", "disp"), ("vox", "disp_vox")):
        jd = O.jacobian_determinant(g[key])
        np.testing.assert_allclose(jd, g[f"jd_{tag}"].numpy(), rtol=1e-12, atol=1e-12)
        assert abs(O.jdstd(g[key]) - float(g[f"jdstd_{tag}"])) < 1e-12
        assert O.jdlessthan0(g[key]) == int(g[f"jdneg_{tag}"])
    assert int(g["jdneg_vox"]) > 0      # the folded field really has non-positive determinants


def _example_inputs(g):
    C = int(g["num_classes"])
    img_f, img_m = g["img_f_u8"].float() / 255, g["img_m_u8"].float() / 255
    oh = lambda lab: torch.nn.functional.one_hot(lab[:, 0].long(), C).permute(0, 4, 1, 2, 3).float()  # noqa: E731
    return img_f, img_m, oh(g["lab_f"]), oh(g["lab_m"]), C


def test_golden_example_pair_config1(golden):
    """BASELINE config 1 (bundled example_data_half pair, reduced to 64^3): the oracle pipeline on real
    anatomy against the reference's own forward / align_img / MSELoss / DiceLoss / jdstd outputs."""
    import keymorph_b200 as kb
    g = golden("example_pair64")
    img_f, img_m, seg_f, seg_m, _ = _example_inputs(g)
    torch.manual_seed(23)
    net = kb.TruncatedUNet3D(1, 32, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                             num_levels=4, is_segmentation=False, conv_padding=1)
    types = ["rigid", "affine", "tps_1"]
    res = O.keymorph_forward("truncatedunet", net.state_dict(), img_f, img_m, types)
    for t in types:
        r = res[t]
        assert_close(r["points_f"], g[f"{t}_points_f"], rtol=1e-5, atol=1e-5)
        assert_close(r["points_m"], g[f"{t}_points_m"], rtol=1e-5, atol=1e-5)
        assert_close(r["grid"][:, ::4, ::4, ::4], g[f"{t}_grid"], rtol=1e-4, atol=1e-4)
        img_a, seg_a = O.align_img(r["grid"], img_m), O.align_img(r["grid"], seg_m)
        assert_close(O.mse_loss(img_a, img_f), g[f"{t}_mse"], rtol=1e-3, atol=1e-6)
        assert_close(O.dice_loss(seg_a, seg_f), g[f"{t}_softdice"], rtol=1e-3, atol=1e-5)
        assert_close(O.dice_loss(seg_a, seg_f, hard=True), g[f"{t}_harddice"], rtol=2e-3, atol=1e-4)
        assert abs(O.jdstd(r["grid"].permute(0, 4, 1, 2, 3)) - float(g[f"{t}_jdstd"])) < 1e-6


def test_nifti_reader_reproduces_the_fixture(golden):
    """keymorph_b200.hostio.load_volume (NIfTI-1 + canonical orientation + block resize + rescale, no
    nibabel / torchio) on the reference's bundled files; skipped where /root/reference is absent."""
    import os
    import pytest
    from keymorph_b200 import hostio
    d = "/root/reference/example_data_half"
    if not os.path.isdir(d):
        pytest.skip("reference example data not available on this machine")
    g = golden("example_pair64")
    img, aff = hostio.load_volume(os.path.join(d, "img_m", "IXI_001_128x128x128.nii.gz"), size=64)
    lab, _ = hostio.load_volume(os.path.join(d, "seg_m", "IXI_001_128x128x128.nii.gz"), size=64, labels=True)
    assert img.shape == (1, 1, 64, 64, 64) and float(img.min()) == 0.0 and float(img.max()) == 1.0
    assert torch.equal((img * 255).round().to(torch.uint8), g["img_f_u8"])
    assert torch.equal(lab, g["lab_f"]) and int(lab.max()) == 13
    # canonical (RAS+) orientation: the file's sform diag(-1,-1,1) is flipped on axes 0 and 1
    assert aff[0, 0] > 0 and aff[1, 1] > 0 and aff[2, 2] > 0
    raw, a0 = hostio.read_nifti(os.path.join(d, "img_m", "IXI_001_128x128x128.nii.gz"))
    assert raw.shape == (256, 256, 256) and a0[0, 0] == -1.0 and a0[1, 1] == -1.0


# ---------------------------------------------------------------- rigid reflection case
def test_golden_rigid_reflection(golden):
    """Mirror-related keypoint sets: det(V U^T) < 0, the reference negates the last ROW of V
    (keymorph/keypoint_aligners.py:199-206).  Matrices produced by the reference itself."""
    g = golden("aligners_reflection")
    for case in range(3):
        pm, pf, w = g[f"c{case}_points_m"], g[f"c{case}_points_f"], g[f"c{case}_w"]
        for tag, ww in (("", None), ("_w", w)):
            tm, inv = O.aligner_matrices(pm, pf, ww, "rigid")
            assert_close(inv, g[f"c{case}_rigid{tag}_inverse"], rtol=1e-5, atol=1e-5)
            assert_close(tm, g[f"c{case}_rigid{tag}_matrix"], rtol=1e-5, atol=1e-5)
            assert_close(O.transform_points(tm, pm), g[f"c{case}_rigid{tag}_points_a"], rtol=1e-5, atol=1e-5)
            # the reference's result is a proper rotation, but NOT the least-squares optimal one
            assert abs(float(torch.det(inv[0, :3, :3])) - 1.0) < 1e-4
        tm, _ = O.aligner_matrices(pm, pf, None, "affine")
        assert_close(tm, g[f"c{case}_affine_matrix"], rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------- real-world coordinates (a16)
def _rot2(r):
    return torch.tensor([[np.cos(r), -np.sin(r), 0], [np.sin(r), np.cos(r), 0], [0, 0, 1]]).float()[None]


REALWORLD_KATS = {
    # reference test/test.py:550-719 (2-D known answers; a 5x5 grid, -1 -> -0.5 and 1 -> 4.5)
    "norm": [[1, 0], [0, -1], [-1, 0], [0, 1]],
    "voxel": [[4.5, 2.0], [2.0, -0.5], [-0.5, 2.0], [2.0, 4.5]],
    "real": [[2, -4.5], [-0.5, -2], [2, 0.5], [4.5, -2]],       # voxel rotated by -90 degrees
}


def realworld_kat_cases(mod):
    """(name, got, expected) for the six convert_points_* KATs, evaluated with module `mod`."""
    t = {k: torch.tensor(v).float().view(1, 4, 2) for k, v in REALWORLD_KATS.items()}
    size, aff = torch.tensor([[5, 5]]).float(), _rot2(-np.pi / 2)
    return [("norm2voxel", mod.convert_points_norm2voxel(t["norm"], size), t["voxel"]),
            ("voxel2norm", mod.convert_points_voxel2norm(t["voxel"], size), t["norm"]),
            ("voxel2real", mod.convert_points_voxel2real(t["voxel"], aff), t["real"]),
            ("real2voxel", mod.convert_points_real2voxel(t["real"], aff), t["voxel"]),
            ("norm2real", mod.convert_points_norm2real(t["norm"], aff, size), t["real"]),
            ("real2norm", mod.convert_points_real2norm(t["real"], aff, size), t["norm"])]


def test_realworld_kats():
    for name, got, want in realworld_kat_cases(O):
        assert_close(got, want, rtol=1e-6, atol=1e-6, msg=name)


def test_golden_realworld(golden):
    g = golden("realworld")
    aff_m, aff_f, sm, sf, pts = g["aff_m"], g["aff_f"], g["shape_m"][None], g["shape_f"][None], g["pts"]
    vox = O.convert_points_norm2voxel(pts, sm)
    assert_close(vox, g["norm2voxel"], rtol=0, atol=0)
    assert_close(O.convert_points_voxel2norm(vox, sm), g["voxel2norm"], rtol=0, atol=0)
    real = O.convert_points_voxel2real(vox, aff_m)
    assert_close(real, g["voxel2real"], rtol=0, atol=0)
    assert_close(O.convert_points_real2voxel(real, aff_m), g["real2voxel"], rtol=0, atol=0)
    assert_close(O.convert_points_norm2real(pts, aff_m, sm), g["norm2real"], rtol=0, atol=0)
    assert_close(O.convert_points_real2norm(real, aff_f, sf), g["real2norm"], rtol=0, atol=0)
    pm, pf = g["points_m"], g["points_f"]
    for tag, t in (("affine", "affine"), ("rigid", "rigid"), ("tps1", "tps_1"), ("tps0", "tps_0")):
        r = O.register_points_real_world(pf, pm, t, sf, sm, aff_f, aff_m, (18, 16, 20))
        tol = 2e-3 if tag == "tps0" else 1e-4      # lambda = 0 in scanner units: cond(A) ~ 1e7 in fp32
        if "matrix" in r:
            assert_close(r["matrix"], g[f"{tag}_matrix"], rtol=1e-4, atol=1e-4)
        assert_close(r["grid"], g[f"{tag}_grid"], rtol=0, atol=tol)
        assert_close(r["points_a"], g[f"{tag}_points_a"], rtol=0, atol=tol)


def test_golden_example_pair_config1_128(golden):
    """BASELINE config 1 at its STATED size (128^3, K = 128): oracle pipeline vs the reference's outputs."""
    import keymorph_b200 as kb
    g = golden("example_pair128")
    img_f, img_m, seg_f, seg_m, _ = _example_inputs(g)
    torch.manual_seed(23)
    net = kb.TruncatedUNet3D(1, 128, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                             num_levels=4, is_segmentation=False, conv_padding=1)
    r = O.keymorph_forward("truncatedunet", net.state_dict(), img_f, img_m, ["affine"])["affine"]
    assert_close(r["points_f"], g["affine_points_f"], rtol=1e-5, atol=1e-5)
    assert_close(r["points_m"], g["affine_points_m"], rtol=1e-5, atol=1e-5)
    assert_close(r["matrix"], g["affine_matrix"], rtol=1e-3, atol=1e-3)
    assert_close(r["grid"][:, ::8, ::8, ::8], g["affine_grid"], rtol=1e-3, atol=1e-3)
    img_a = O.align_img(r["grid"], img_m)
    assert_close(O.mse_loss(img_a, img_f), g["affine_mse"], rtol=2e-3, atol=1e-6)


def test_stock_run_eval_harness_with_the_reference_model(tmp_path, golden):
    """The harness that drives the reference's unmodified run_eval (tests/stock_eval_harness.py) with the
    REFERENCE model on the CPU: validates the harness itself (the GPU suite runs it on keymorph_b200)."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference not installed (oracle/build_ref.py)")
    import stock_eval_harness as H
    g = golden("example_pair64")
    img_f, img_m = g["img_f_u8"].float() / 255, g["img_m_u8"].float() / 255
    model = refshim.build_reference_model(32)
    metrics, save_dir = H.run_stock_eval(model, img_f, img_m, g["lab_f"], g["lab_m"], tmp_path, "cpu",
                                         aligns=("affine",))
    assert {p.name for p in save_dir.iterdir()} == H.expected_files(("affine",))
    key = "mse:img_m/IXI_001:img_m/IXI_002:rot0:affine"
    # (the script resamples the moving image through its own affine_augment(rot0) first, so the number is
    # close to, not equal to, the fixture's un-augmented one)
    assert abs(metrics[key][0] - float(g["affine_mse"])) < 0.1 * float(g["affine_mse"])
