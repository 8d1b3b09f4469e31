"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (alanqrwang/keymorph, read-only at
/root/reference) on seeded inputs.  Runs only in the build container (the GPU box has no
/root/reference); the fixtures it writes are committed and pin oracle/keymorph_oracle.py.

    python oracle/gen_golden.py            # rewrites tests/golden/

The reference's package __init__ eagerly imports nibabel / skimage / h5py / torchio / matplotlib,
none of which is installed or used on the hot path; they are replaced by MagicMock modules.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("KEYMORPH_REFERENCE", "/root/reference")
_MOCKED = {"nibabel", "skimage", "h5py", "torchio", "matplotlib"}


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _MOCKED:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


def import_reference():
    sys.meta_path.insert(0, _MockFinder())
    sys.path.insert(0, REF)
    import keymorph  # noqa: F401
    return keymorph


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, keys {sorted(conv)}")


def gen_jacobian():
    """Jacobian-determinant metrics (keymorph/loss_ops.py:161-247) of an affine grid and of a folded
    (non-diffeomorphic) field, computed by the reference functions themselves."""
    from keymorph import loss_ops
    from keymorph.keypoint_aligners import TPS
    g = torch.Generator().manual_seed(7)
    pts_f = torch.rand(1, 24, 3, generator=g) * 1.6 - 0.8
    pts_m = pts_f + 0.15 * torch.randn(1, 24, 3, generator=g)
    grid = TPS(pts_m, pts_f, torch.tensor([0.0])).get_flow_field((1, 1, 14, 12, 16))     # (1,14,12,16,3)
    disp = grid.permute(0, 4, 1, 2, 3).contiguous()
    # the reference is handed the NORMALISED grid (pairwise_register_eval.py:337-338); a second case in
    # voxel units exercises determinants <= 0
    disp_vox = disp * torch.tensor([8.0, 6.0, 7.0]).view(1, 3, 1, 1, 1) * torch.randn(1, 3, 14, 12, 16, generator=g).sign()
    out = {"disp": disp, "disp_vox": disp_vox}
    for tag, d in (("norm", disp), ("vox", disp_vox)):
        out[f"jdstd_{tag}"] = np.float64(loss_ops.jdstd(d.numpy()))
        out[f"jdneg_{tag}"] = np.int64(loss_ops.jdlessthan0(d.numpy()))
        out[f"jd_{tag}"] = loss_ops._jacobian_determinant(d.numpy())
    save("jacobian", **out)


def gen_example_pair():
    """BASELINE config 1 at fixture size: the bundled example_data_half pair (real T1 anatomy + 14-label
    segmentations, read with keymorph_b200.hostio -- no nibabel / torchio here), block-averaged to 64^3
    and quantised to uint8 so that the fixture is small and every consumer sees identical floats,
    through the REFERENCE model (seeded TruncatedUNet3D, K = 32) and the reference's align_img /
    MSELoss / DiceLoss / jdstd."""
    sys.path.insert(0, ROOT)
    from keymorph import loss_ops, utils
    from keymorph.model import KeyMorph
    from keymorph.unet3d.model import TruncatedUNet3D
    from keymorph_b200 import hostio
    d = os.path.join(REF, "example_data_half")
    f_img, _ = hostio.load_volume(os.path.join(d, "img_m", "IXI_001_128x128x128.nii.gz"), size=64)
    m_img, _ = hostio.load_volume(os.path.join(d, "img_m", "IXI_002_128x128x128.nii.gz"), size=64)
    f_seg, _ = hostio.load_volume(os.path.join(d, "seg_m", "IXI_001_128x128x128.nii.gz"), size=64, labels=True)
    m_seg, _ = hostio.load_volume(os.path.join(d, "seg_m", "IXI_002_128x128x128.nii.gz"), size=64, labels=True)
    f_u8 = (f_img * 255).round().to(torch.uint8)
    m_u8 = (m_img * 255).round().to(torch.uint8)
    img_f, img_m = f_u8.float() / 255, m_u8.float() / 255
    C = int(max(f_seg.max(), m_seg.max())) + 1
    oh = lambda lab: torch.nn.functional.one_hot(lab[:, 0].long(), C).permute(0, 4, 1, 2, 3).float()  # noqa: E731
    seg_f, seg_m = oh(f_seg), oh(m_seg)
    torch.manual_seed(23)
    net = TruncatedUNet3D(1, 32, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                          num_levels=4, is_segmentation=False, conv_padding=1)
    model = KeyMorph(torch.nn.DataParallel(net), 32, 3).eval()
    types = ["rigid", "affine", "tps_1"]
    with torch.no_grad():
        res = model(img_f, img_m, transform_type=types, return_aligned_points=True)
    out = {"img_f_u8": f_u8, "img_m_u8": m_u8, "lab_f": f_seg, "lab_m": m_seg, "num_classes": np.int64(C)}
    for t in types:
        r = res[t]
        img_a = utils.align_img(r["grid"], img_m)
        seg_a = utils.align_img(r["grid"], seg_m)
        out[f"{t}_points_f"], out[f"{t}_points_m"] = r["points_f"], r["points_m"]
        out[f"{t}_grid"] = r["grid"][:, ::4, ::4, ::4]
        if "matrix" in r:
            out[f"{t}_matrix"] = r["matrix"]
        out[f"{t}_mse"] = loss_ops.MSELoss()(img_a, img_f)
        out[f"{t}_softdice"] = loss_ops.DiceLoss()(seg_a, seg_f)
        out[f"{t}_harddice"] = loss_ops.DiceLoss(hard=True)(seg_a, seg_f)
        out[f"{t}_jdstd"] = np.float64(loss_ops.jdstd(r["grid"].permute(0, 4, 1, 2, 3).numpy()))
        out[f"{t}_jdneg"] = np.int64(loss_ops.jdlessthan0(r["grid"].permute(0, 4, 1, 2, 3).numpy()))
    save("example_pair64", **out)


def gen_augment_aniso():
    """Anisotropic augmentation (keymorph/augmentation.py:81-178) through the reference's own
    AffineDeformation3d on the CPU: matrix, moved points, bilinear image and nearest segmentation for
    one volume (the reference's affine_grid supports batch size 1 only)."""
    from keymorph.augmentation import AffineDeformation3d
    g = torch.Generator().manual_seed(5)
    scale = torch.empty(1, 3).uniform_(0.8, 1.2, generator=g)
    offset = torch.empty(1, 3).uniform_(-0.2, 0.2, generator=g)
    theta = torch.empty(1, 3).uniform_(-3.1416, 3.1416, generator=g)
    shear = torch.empty(1, 6).uniform_(-0.1, 0.1, generator=g)
    params = (scale, offset, theta, shear)
    img = torch.rand(1, 2, 12, 16, 20, generator=g)
    seg = torch.randint(0, 5, (1, 1, 12, 16, 20), generator=g).float()
    pts = torch.rand(1, 9, 3, generator=g) * 2 - 1
    aug = AffineDeformation3d(device="cpu")
    save("augment_aniso", scale=scale, offset=offset, theta=theta, shear=shear, img=img, seg=seg, points=pts,
         matrix=aug.build_affine_matrix(1, params), img_aug=aug(img, params=params, interp_mode="bilinear"),
         seg_aug=aug(seg, params=params, interp_mode="nearest"), points_aug=aug.deform_points(pts, params))


def gen_group_metrics():
    """Evaluation metrics of SURVEY.md 8f-3 computed by the reference itself (keymorph/loss_ops.py):
    hausdorff_distance / fast_dice on a pair of soft 4-class maps, and the O(G^2) group metrics
    (MSEPairwiseLoss, MultipleAvgSegPairwiseMetric, MultipleAvgGridMetric) on a group of three."""
    from keymorph import loss_ops
    from keymorph.keypoint_aligners import TPS
    g = torch.Generator().manual_seed(31)
    G, C, shape = 3, 4, (12, 14, 16)
    zz, yy, xx = torch.meshgrid(*[torch.linspace(-1, 1, n) for n in shape], indexing="ij")
    segs, imgs, grids = [], [], []
    for i in range(G):
        c = torch.rand(C, 3, generator=g) * 1.2 - 0.6
        r = torch.rand(C, generator=g) * 0.4 + 0.3
        logits = torch.stack([-(((zz - c[k, 0]) ** 2 + (yy - c[k, 1]) ** 2 + (xx - c[k, 2]) ** 2) / r[k] ** 2)
                              for k in range(C)])
        logits[0] = -1.0                                   # channel 0 = background shell around the blobs
        segs.append(torch.softmax(4 * logits, 0))
        imgs.append(torch.rand(1, *shape, generator=g))
        pf = torch.rand(1, 12, 3, generator=g) * 1.6 - 0.8
        pm = pf + 0.1 * torch.randn(1, 12, 3, generator=g)
        grids.append(TPS(pm, pf, torch.tensor([0.0])).get_flow_field((1, 1) + shape)[0])
    segs, imgs, grids = torch.stack(segs), torch.stack(imgs), torch.stack(grids)
    hard = torch.nn.functional.one_hot(segs.argmax(1), C).permute(0, 4, 1, 2, 3).float()
    names = ["dice", "harddice", "harddiceroi", "softdice", "hausd"]
    seg_m = loss_ops.MultipleAvgSegPairwiseMetric()(hard, names)
    grid_m = loss_ops.MultipleAvgGridMetric()(grids, ["jdstd", "jdlessthan0"])
    out = {"segs": segs, "imgs": imgs, "grids": grids,
           "hausd_01": np.float64(loss_ops.hausdorff_distance(hard[0:1], hard[1:2])),
           "hausd_batch": np.float64(loss_ops.hausdorff_distance(hard[0:2], hard[1:3])),
           "hausd_blobs": np.array([loss_ops.hausdorff_distance(hard[i:i + 1, k:], hard[j:j + 1, k:])
                                    for (i, j) in ((0, 1), (0, 2), (1, 2)) for k in (1, 2, 3)]),
           "fast_dice_01": np.float64(loss_ops.fast_dice(segs[0:1].numpy(), segs[1:2].numpy())),
           "mse_pairwise": loss_ops.MSEPairwiseLoss()(imgs),
           "softdice_pairwise": loss_ops.SoftDicePairwiseLoss()(segs),
           "harddice_pairwise": loss_ops.HardDicePairwiseLoss()(segs),
           "avg_jdstd": np.float64(loss_ops.AvgJDStd()(grids)),
           "avg_jdneg": np.float64(loss_ops.AvgJDLessThan0()(grids))}
    for k, v in seg_m.items():
        out[f"multi_{k}"] = v if isinstance(v, torch.Tensor) else np.float64(v)
    for k, v in grid_m.items():
        out[f"multi_{k}"] = np.float64(v)
    save("group_metrics", **out)


def gen_example_pair128():
    """BASELINE config 1 AT ITS STATED SIZE: the bundled example_data_half pair at 128^3 (block mean of the
    256^3 files, quantised to uint8: 0.2 MB per volume compressed), TruncatedUNet3D, K = 128, through the
    REFERENCE model and the reference's align_img / MSELoss / DiceLoss / jdstd (scripts/register.py:40-118
    -> pairwise_register_eval.py:116-171).  Grids are stored on a stride-8 lattice."""
    sys.path.insert(0, ROOT)
    from keymorph import loss_ops, utils
    from keymorph.model import KeyMorph
    from keymorph.unet3d.model import TruncatedUNet3D
    from keymorph_b200 import hostio
    d = os.path.join(REF, "example_data_half")
    S, K = 128, 128
    f_img, _ = hostio.load_volume(os.path.join(d, "img_m", "IXI_001_128x128x128.nii.gz"), size=S)
    m_img, _ = hostio.load_volume(os.path.join(d, "img_m", "IXI_002_128x128x128.nii.gz"), size=S)
    f_seg, _ = hostio.load_volume(os.path.join(d, "seg_m", "IXI_001_128x128x128.nii.gz"), size=S, labels=True)
    m_seg, _ = hostio.load_volume(os.path.join(d, "seg_m", "IXI_002_128x128x128.nii.gz"), size=S, labels=True)
    f_u8 = (f_img * 255).round().to(torch.uint8)
    m_u8 = (m_img * 255).round().to(torch.uint8)
    img_f, img_m = f_u8.float() / 255, m_u8.float() / 255
    C = int(max(f_seg.max(), m_seg.max())) + 1
    oh = lambda lab: torch.nn.functional.one_hot(lab[:, 0].long(), C).permute(0, 4, 1, 2, 3).float()  # noqa: E731
    seg_f, seg_m = oh(f_seg), oh(m_seg)
    torch.manual_seed(23)
    net = TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                          num_levels=4, is_segmentation=False, conv_padding=1)
    model = KeyMorph(torch.nn.DataParallel(net), K, 3).eval()
    types = ["rigid", "affine", "tps_1"]
    with torch.no_grad():
        res = model(img_f, img_m, transform_type=types, return_aligned_points=True)
    out = {"img_f_u8": f_u8, "img_m_u8": m_u8, "lab_f": f_seg, "lab_m": m_seg, "num_classes": np.int64(C)}
    for t in types:
        r = res[t]
        img_a = utils.align_img(r["grid"], img_m)
        seg_a = utils.align_img(r["grid"], seg_m)
        out[f"{t}_points_f"], out[f"{t}_points_m"], out[f"{t}_points_a"] = r["points_f"], r["points_m"], r["points_a"]
        out[f"{t}_grid"] = r["grid"][:, ::8, ::8, ::8]
        out[f"{t}_img_a"] = img_a[:, :, ::8, ::8, ::8]
        if "matrix" in r:
            out[f"{t}_matrix"] = r["matrix"]
        out[f"{t}_mse"] = loss_ops.MSELoss()(img_a, img_f)
        out[f"{t}_softdice"] = loss_ops.DiceLoss()(seg_a, seg_f)
        out[f"{t}_harddice"] = loss_ops.DiceLoss(hard=True)(seg_a, seg_f)
        out[f"{t}_jdstd"] = np.float64(loss_ops.jdstd(r["grid"].permute(0, 4, 1, 2, 3).numpy()))
        out[f"{t}_jdneg"] = np.int64(loss_ops.jdlessthan0(r["grid"].permute(0, 4, 1, 2, 3).numpy()))
    save("example_pair128", **out)


def gen_reflection():
    """Mirror-related point sets (an L/R-flipped volume against the original): det(V U^T) < 0 and the
    reference's reflection step (keymorph/keypoint_aligners.py:199-206, last ROW of V negated) decides the
    rigid matrix.  Unweighted and weighted, plus the affine fit of the same points."""
    from keymorph.keypoint_aligners import AffineKeypointAligner, RigidKeypointAligner
    g = torch.Generator().manual_seed(17)
    out = {}
    for case in range(3):
        K = (12, 40, 7)[case]
        pm = torch.rand(1, K, 3, generator=g) * 1.4 - 0.7
        mirror = torch.tensor([1.0, 1.0, -1.0]) if case != 1 else torch.tensor([-1.0, 1.0, 1.0])
        pf = pm * mirror + 0.02 * torch.randn(1, K, 3, generator=g) + torch.tensor([0.03, -0.02, 0.01])
        w = torch.rand(1, K, generator=g)
        w = w / w.sum()
        out[f"c{case}_points_m"], out[f"c{case}_points_f"], out[f"c{case}_w"] = pm, pf, w
        for wtag, ww in (("", None), ("_w", w)):
            al = RigidKeypointAligner(pm, pf, w=ww, dim=3)
            out[f"c{case}_rigid{wtag}_matrix"] = al.transform_matrix
            out[f"c{case}_rigid{wtag}_inverse"] = al.inverse_transform_matrix
            out[f"c{case}_rigid{wtag}_det"] = torch.det(al.inverse_transform_matrix[:, :3, :3])
            out[f"c{case}_rigid{wtag}_points_a"] = al.get_forward_transformed_points(pm)
        al = AffineKeypointAligner(pm, pf, dim=3)
        out[f"c{case}_affine_matrix"] = al.transform_matrix
    save("aligners_reflection", **out)


def gen_realworld():
    """Real-world-coordinate alignment (keymorph/utils.py:243-354,
    keypoint_aligners.py:22-147,219-274,435-465, model.py:164-170): the six convert_points_* helpers on
    random 3-D inputs, the three aligners with align_in_real_world_coords=True (matrices / theta, flow
    field, forward and inverse points), and KeyMorph.forward(align_keypoints_in_real_world_coords=True)."""
    from keymorph import utils
    from keymorph.keypoint_aligners import TPS, AffineKeypointAligner, RigidKeypointAligner
    from keymorph.model import KeyMorph
    from keymorph.unet3d.model import TruncatedUNet3D
    from oracle.keymorph_oracle import affine_augment, gaussian_phantom
    g = torch.Generator().manual_seed(29)
    K = 20

    def rand_affine(spacing, origin, rot):
        a = torch.eye(4)
        c, s_ = np.cos(rot), np.sin(rot)
        R = torch.tensor([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
        a[:3, :3] = R @ torch.diag(torch.tensor(spacing))
        a[:3, 3] = torch.tensor(origin)
        return a[None]

    aff_m = rand_affine([1.0, 1.2, 0.8], [-60.0, -70.0, -50.0], 0.1)
    aff_f = rand_affine([0.9, 1.0, 1.1], [-64.0, -60.0, -66.0], -0.05)
    shape_m = torch.tensor([16.0, 20.0, 24.0])
    shape_f = torch.tensor([18.0, 16.0, 20.0])
    pts = torch.rand(1, K, 3, generator=g) * 2 - 1
    out = {"aff_m": aff_m, "aff_f": aff_f, "shape_m": shape_m, "shape_f": shape_f, "pts": pts}
    vox = utils.convert_points_norm2voxel(pts, shape_m[None])
    out["norm2voxel"] = vox
    out["voxel2norm"] = utils.convert_points_voxel2norm(vox, shape_m[None])
    real = utils.convert_points_voxel2real(vox, aff_m)
    out["voxel2real"] = real
    out["real2voxel"] = utils.convert_points_real2voxel(real, aff_m)
    out["norm2real"] = utils.convert_points_norm2real(pts, aff_m, shape_m[None])
    out["real2norm"] = utils.convert_points_real2norm(real, aff_f, shape_f[None])
    # aligners: keypoints of the SAME anatomy seen through two different voxel grids
    pm = torch.rand(1, K, 3, generator=g) * 1.2 - 0.6
    real_m = utils.convert_points_norm2real(pm, aff_m, shape_m[None])
    lin = torch.eye(3) + 0.05 * torch.randn(3, 3, generator=g)
    real_f = real_m @ lin.T + torch.tensor([2.0, -1.5, 1.0]) + 0.3 * torch.randn(1, K, 3, generator=g)
    pf = utils.convert_points_real2norm(real_f, aff_f, shape_f[None])
    out["points_m"], out["points_f"] = pm, pf
    gshape = (1, 1, 18, 16, 20)
    common = dict(align_in_real_world_coords=True, aff_m=aff_m, aff_f=aff_f, shape_m=shape_m[None],
                  shape_f=shape_f[None], dim=3)
    for tag, make in (("affine", lambda: AffineKeypointAligner(pm, pf, **common)),
                      ("rigid", lambda: RigidKeypointAligner(pm, pf, **common)),
                      ("tps1", lambda: TPS(pm, pf, torch.tensor([1.0]), **common)),
                      ("tps0", lambda: TPS(pm, pf, torch.tensor([0.0]), **common))):
        al = make()
        if tag in ("affine", "rigid"):
            out[f"{tag}_matrix"] = al.transform_matrix
            out[f"{tag}_inverse"] = al.inverse_transform_matrix
        else:
            out[f"{tag}_inverse_theta"] = al.inverse_theta
        out[f"{tag}_grid"] = al.get_flow_field(gshape, compute_on_subgrids=True)
        out[f"{tag}_points_a"] = al.get_forward_transformed_points(pm)
        out[f"{tag}_points_inv"] = al.get_inverse_transformed_points(pf)
    # the whole pipeline with align_keypoints_in_real_world_coords=True (32^3, identity-like affines)
    torch.manual_seed(23)
    net = TruncatedUNet3D(1, 16, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                          num_levels=4, is_segmentation=False, conv_padding=1)
    model = KeyMorph(torch.nn.DataParallel(net), 16, 3, align_keypoints_in_real_world_coords=True).eval()
    img_f = gaussian_phantom(32, 1000)
    img_m = affine_augment(gaussian_phantom(32, 1000), (0.05, 0.03, 0.15, 0.01))
    a_f = rand_affine([2.0, 2.0, 2.0], [-32.0, -32.0, -32.0], 0.0)
    a_m = rand_affine([2.0, 2.2, 1.9], [-30.0, -35.0, -31.0], 0.02)
    with torch.no_grad():
        res = model(img_f, img_m, transform_type=["rigid", "affine", "tps_1"], return_aligned_points=True,
                    aff_f=a_f, aff_m=a_m)
    out.update(fw_img_f=img_f, fw_img_m=img_m, fw_aff_f=a_f, fw_aff_m=a_m)
    for t, r in res.items():
        out[f"fw_{t}_points_f"], out[f"fw_{t}_points_m"] = r["points_f"], r["points_m"]
        out[f"fw_{t}_points_a"] = r["points_a"]
        out[f"fw_{t}_grid"] = r["grid"][:, ::2, ::2, ::2]
        if "matrix" in r:
            out[f"fw_{t}_matrix"] = r["matrix"]
    save("realworld", **out)


def main():
    import_reference()
    sys.path.insert(0, ROOT)
    if "--only-group-metrics" in sys.argv:
        gen_group_metrics()
        return
    if "--only-augment" in sys.argv:
        gen_augment_aniso()
        return
    if "--only-jacobian" in sys.argv:
        gen_jacobian()
        return
    if "--only-example" in sys.argv:
        gen_example_pair()
        return
    only = {"--only-example128": gen_example_pair128, "--only-reflection": gen_reflection,
            "--only-realworld": gen_realworld}
    picked = [fn for flag, fn in only.items() if flag in sys.argv]
    if picked:
        for fn in picked:
            fn()
        return
    from keymorph import layers, loss_ops, utils
    from keymorph.augmentation import affine_augment
    from keymorph.keypoint_aligners import TPS, AffineKeypointAligner, RigidKeypointAligner
    from keymorph.model import KeyMorph
    from keymorph.net import ConvNet
    from keymorph.unet3d.model import TruncatedUNet3D, UNet3D
    from oracle.keymorph_oracle import gaussian_phantom

    torch.set_num_threads(8)

    # ---------------------------------------------------------------- aligners
    g = torch.Generator().manual_seed(11)
    K = 32
    pm = torch.rand(1, K, 3, generator=g) * 1.2 - 0.6
    lin = torch.eye(3) + 0.15 * torch.randn(3, 3, generator=g)
    pf = pm @ lin.T + 0.05 * torch.randn(1, K, 3, generator=g) + torch.tensor([0.05, -0.1, 0.02])
    w = torch.rand(1, K, generator=g)
    w = w / w.sum()
    shape = (1, 1, 8, 10, 12)
    out = {"points_m": pm, "points_f": pf, "w": w, "shape": np.array(shape)}
    for tag, cls in (("affine", AffineKeypointAligner), ("rigid", RigidKeypointAligner)):
        for wtag, ww in (("", None), ("_w", w)):
            al = cls(pm, pf, w=ww, dim=3)
            out[f"{tag}{wtag}_matrix"] = al.transform_matrix
            out[f"{tag}{wtag}_inverse"] = al.inverse_transform_matrix
            out[f"{tag}{wtag}_grid"] = al.get_flow_field(shape)
            out[f"{tag}{wtag}_points_a"] = al.get_forward_transformed_points(pm)
            out[f"{tag}{wtag}_points_inv"] = al.get_inverse_transformed_points(pf)
    save("aligners", **out)

    # ---------------------------------------------------------------- TPS
    K = 24
    pm = torch.rand(1, K, 3, generator=g) * 1.2 - 0.6
    pf = pm + 0.08 * torch.randn(1, K, 3, generator=g)
    w = torch.rand(1, K, generator=g)
    w = w / w.sum()
    out = {"points_m": pm, "points_f": pf, "w": w, "shape": np.array(shape)}
    for lam in (0.0, 0.1, 10.0):
        for wtag, ww in (("", None), ("_w", w)):
            if ww is not None and lam != 0.1:
                continue
            tps = TPS(pm, pf, torch.tensor([lam]), w=ww, dim=3, num_subgrids=4)
            tag = f"lam{lam:g}{wtag}"
            out[f"{tag}_inverse_theta"] = tps.inverse_theta
            out[f"{tag}_grid"] = tps.get_flow_field(shape, compute_on_subgrids=True)
            out[f"{tag}_points_a"] = tps.get_forward_transformed_points(pm)
            out[f"{tag}_theta"] = tps.theta
    save("tps", **out)

    # ---------------------------------------------------------------- warp + losses
    x = torch.randn(2, 3, 9, 10, 11, generator=g)
    grid = torch.rand(2, 7, 8, 12, 3, generator=g) * 2.4 - 1.2
    # exact half-way coordinates exercise round-half-even in nearest mode
    grid[0, 0, 0, :4, 0] = torch.tensor([-1 + 1.0 / 11, -1 + 3.0 / 11, -1 + 5.0 / 11, 1.0])
    out = {"x": x, "grid": grid,
           "bilinear": utils.align_img(grid, x, "bilinear"),
           "nearest": utils.align_img(grid, x, "nearest")}
    lab_p = torch.randint(0, 5, (2, 1, 6, 7, 8), generator=g)
    lab_t = torch.randint(0, 5, (2, 1, 6, 7, 8), generator=g)
    seg_t = torch.nn.functional.one_hot(lab_t)[:, 0].permute(0, 4, 1, 2, 3).float()
    soft_p = torch.softmax(torch.randn(2, 5, 6, 7, 8, generator=g) +
                           2 * torch.nn.functional.one_hot(lab_p)[:, 0].permute(0, 4, 1, 2, 3).float(), 1)
    out.update(seg_pred=soft_p, seg_target=seg_t)
    out["mse"] = loss_ops.MSELoss()(soft_p, seg_t)
    for hard in (False, True):
        for ign in (False, True):
            out[f"dice_h{int(hard)}_i{int(ign)}"] = loss_ops.DiceLoss(hard=hard)(soft_p, seg_t, ign_first_ch=ign)
            out[f"dice_regions_h{int(hard)}_i{int(ign)}"] = loss_ops.DiceLoss(hard=hard, return_regions=True)(
                soft_p, seg_t, ign_first_ch=ign)
    save("warp_loss", **out)

    # ---------------------------------------------------------------- CoM on a random heat map
    heat = torch.randn(2, 5, 9, 12, 16, generator=g)
    save("com", heat=heat, points_ij=layers.CenterOfMass3d(indexing="ij")(heat),
         points_xy=layers.CenterOfMass3d(indexing="xy")(heat))

    # ---------------------------------------------------------------- augmentation
    img = gaussian_phantom(16, 5)
    seg = (img > 0.3).float()
    params = (0.1, 0.05, 0.3, 0.02)
    a_img, a_seg = affine_augment(img, params, seg=seg)
    save("augment", img=img, seg=seg, params=np.array(params), img_aug=a_img, seg_aug=a_seg)

    # ---------------------------------------------------------------- backbones (seed 23 init)
    def state_checksums(net):
        sd = net.state_dict()
        keys = sorted(sd)
        return keys, np.array([float(sd[k].double().abs().sum()) for k in keys])

    img32 = gaussian_phantom(32, 1000)
    img64 = gaussian_phantom(64, 1001)
    com = layers.CenterOfMass3d(indexing="ij")

    torch.manual_seed(23)
    tnet = TruncatedUNet3D(1, 16, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                           num_levels=4, is_segmentation=False, conv_padding=1).eval()
    keys, sums = state_checksums(tnet)
    with torch.no_grad():
        h32 = tnet(img32)
        h64 = tnet(img64)
    save("truncunet_k16", keys=np.array(keys), checksums=sums, heat32=h32, points32=com(h32),
         points64=com(h64), heat64_sub=h64[:, :, ::4, ::4, ::4])

    torch.manual_seed(23)
    unet = UNet3D(1, 16, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                  is_segmentation=False, conv_padding=1).eval()
    keys, sums = state_checksums(unet)
    with torch.no_grad():
        h64 = unet(img64)
    save("unet_k16", keys=np.array(keys), checksums=sums, points64=com(h64),
         heat64_sub=h64[:, :, ::8, ::8, ::8])

    torch.manual_seed(23)
    cnet = ConvNet(3, 1, 16, norm_type="instance").eval()
    keys, sums = state_checksums(cnet)
    with torch.no_grad():
        h64 = cnet(gaussian_phantom(128, 1002))
    save("convnet_k16", keys=np.array(keys), checksums=sums, heat=h64, points=com(h64))

    # ---------------------------------------------------------------- KeyMorph.forward (32^3)
    torch.manual_seed(23)
    tnet = TruncatedUNet3D(1, 16, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                           num_levels=4, is_segmentation=False, conv_padding=1)
    for wk in (None, "power"):
        model = KeyMorph(torch.nn.DataParallel(tnet), 16, 3, weight_keypoints=wk).eval()
        img_f = gaussian_phantom(32, 1000)
        img_m = affine_augment(gaussian_phantom(32, 1000), (0.05, 0.03, 0.15, 0.01))
        types = ["rigid", "affine", "tps_10", "tps_0.1"]
        with torch.no_grad():
            res = model(img_f, img_m, transform_type=types, return_aligned_points=True)
        out = {"img_f": img_f, "img_m": img_m}
        for t in types:
            r = res[t]
            out[f"{t}_grid"] = r["grid"][:, ::2, ::2, ::2]
            out[f"{t}_points_f"] = r["points_f"]
            out[f"{t}_points_m"] = r["points_m"]
            out[f"{t}_points_a"] = r["points_a"]
            if "matrix" in r:
                out[f"{t}_matrix"] = r["matrix"]
            if r["points_weights"] is not None:
                out[f"{t}_weights"] = r["points_weights"]
            out[f"{t}_img_a"] = utils.align_img(r["grid"], img_m)[:, :, ::2, ::2, ::2]
        save("forward32" + ("_power" if wk else ""), **out)

    # ---------------------------------------------------------------- groupwise (directory mode)
    model = KeyMorph(torch.nn.DataParallel(tnet), 16, 3).eval()
    with tempfile.TemporaryDirectory() as d, torch.no_grad():
        subj = [affine_augment(gaussian_phantom(32, 1000), (0.02 * i, 0.02 * i - 0.03, 0.1 * i, 0.0))
                for i in range(4)]
        for i, s in enumerate(subj):
            np.savez(os.path.join(d, f"img_m_{i:03}.npz"), img=s.numpy())
        sd = os.path.join(d, "out")
        os.makedirs(sd)
        res = model.groupwise_register(d, transform_type=["rigid", "affine", "tps_1"], device="cpu", num_iters=3,
                                       log_to_console=False, save_dir=sd, save_results_to_disk=True)
        out = {"subjects": torch.cat(subj, 0)}
        for t in ("rigid", "affine", "tps_1"):
            out[f"{t}_points_m"] = res[t]["grouppoints_m"]
            out[f"{t}_points_a"] = res[t]["grouppoints_a"]
            for i in range(4):
                out[f"{t}_grid_{i}"] = np.load(os.path.join(sd, f"{t}_grid_{i:03}.npy"))[:, ::2, ::2, ::2]
        save("groupwise32", **out)
    gen_jacobian()
    gen_example_pair()
    gen_example_pair128()
    gen_reflection()
    gen_realworld()


if __name__ == "__main__":
    main()
