"""KeyMorph pipeline module with the surface of keymorph/model.py:22-530 (constructor, forward,
get_keypoints, groupwise_register, result-dict keys) running on the B200 kernels."""
from __future__ import annotations

import os
import re
import time

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .engine import backbone_engine
from .keypoint_aligners import TPS, AffineKeypointAligner, RigidKeypointAligner
from .layers import CenterOfMass3d
from .transformations import deferred_singular_checks
from .loss_ops import dice_from_sums
from .utils import str_or_float


# Staging of a graph's static inputs: Tensor.copy_ is a cudaMemcpyAsync, which the driver may hand to a copy
# engine -- where it queues behind the H2D / D2H transfers of the neighbouring registrations in a pipelined
# loop.  An elementwise kernel on the SMs (67 MB: 20 us) never waits for PCIe traffic.
GRAPH_STAGE_WITH_KERNEL = True


def _stage(dst, src):
    if GRAPH_STAGE_WITH_KERNEL:
        torch.mul(src, 1.0, out=dst)
    else:
        dst.copy_(src)


class KeyMorph(nn.Module):
    def __init__(self, backbone, num_keypoints, dim, keypoint_layer="com", max_train_keypoints=None,
                 use_amp=False, use_checkpoint=False, weight_keypoints=None,
                 align_keypoints_in_real_world_coords=False, max_rand_tps_lmbda=10,
                 fused_warp=False, cuda_graph=False):
        """Same arguments as keymorph/model.py:23-35.  `use_amp` / `use_checkpoint` are accepted for
        compatibility (the backbone always runs bf16 operands with fp32 accumulation, and nothing is
        checkpointed at inference).  `fused_warp=True` additionally returns `img_a` (and `seg_a`,
        `mse`, `dice` when segmentations are passed) from forward(); the stock
        scripts/pairwise_register_eval.py:137-138,156-157 uses them instead of calling align_img.
        `cuda_graph=True` captures the whole forward() of a given (shape, transform list) into one CUDA graph
        on its second call and replays it afterwards: see `_forward_graphed`."""
        super().__init__()
        if dim != 3:
            raise NotImplementedError("keymorph_b200 implements the 3-D path only")
        if keypoint_layer != "com":
            raise NotImplementedError("only the centre-of-mass keypoint layer is implemented "
                                      "(LinearRegressor* is broken in the reference, layers.py:15,27)")
        self.backbone = backbone
        self.num_keypoints = num_keypoints
        self.dim = dim
        self.keypoint_layer = CenterOfMass3d(indexing="ij")
        self.max_train_keypoints = max_train_keypoints
        self.use_amp = use_amp
        self.use_checkpoint = use_checkpoint
        self.max_rand_tps_lmbda = max_rand_tps_lmbda
        self.supported_transform_type = ["rigid", "affine", "tps"]
        assert weight_keypoints in [None, "variance", "power"]
        if weight_keypoints == "variance":
            raise NotImplementedError("weight_keypoints='variance' is never reached from the "
                                      "reference's forward() (model.py:183-191)")
        self.weight_keypoints = weight_keypoints
        self.align_keypoints_in_real_world_coords = align_keypoints_in_real_world_coords
        self.fused_warp = fused_warp
        self.cuda_graph = cuda_graph
        self._graphs = {}           # (shape, device, transforms, ...) -> captured forward (insertion order = age)
        self.max_graphs = 4         # captured graphs kept alive (each pins its activations: ~6 GB for a 256^3 pair)
        self._lmbda_cache = {}      # numeric TPS lambdas resident on the device (no H2D copy per call)

    # ------------------------------------------------------------------ keypoints
    def weight_by_power(self, feat1, feat2):
        """keymorph/model.py:95-109 on heat maps (used only when a caller holds feature maps)."""
        p1 = torch.relu(feat1).flatten(2).sum(-1)
        p2 = torch.relu(feat2).flatten(2).sum(-1)
        w = p1 * p2
        return w / w.sum(dim=1, keepdim=True)

    def _keypoints_and_mass(self, img, want_feat=False):
        eng = backbone_engine(self.backbone)
        if eng is None:
            # one backend only: a foreign nn.Module is NOT run through eager torch behind the caller's back
            raise ops._lib.KMError(
                f"keymorph_b200.KeyMorph runs its own backbones (TruncatedUNet3D, UNet3D, ConvNet from this "
                f"package; they load the reference's state dicts), not {type(self.backbone).__name__}: there is "
                "no eager-torch path.  For heat maps computed elsewhere use keymorph_b200.CenterOfMass3d.")
        return eng.keypoints(img, want_mass=True, want_feat=want_feat)

    def get_keypoints(self, img, return_feat=False):
        """keymorph/model.py:111-117."""
        pts, _, feat = self._keypoints_and_mass(img, want_feat=return_feat)
        return (pts, feat) if return_feat else pts

    def _convert_tps_lmbda(self, num_samples, tps_lmbda):
        """keymorph/model.py:119-132."""
        if tps_lmbda == "uniform":
            return torch.rand(num_samples) * self.max_rand_tps_lmbda
        if tps_lmbda == "loguniform":
            from scipy.stats import loguniform
            return torch.tensor(loguniform.rvs(1e-6, self.max_rand_tps_lmbda, size=num_samples))
        return torch.tensor(tps_lmbda).repeat(num_samples)

    def _tps_lmbda_on(self, num_samples, spec, device):
        """_convert_tps_lmbda on the device; numeric lambdas are uploaded once (a pageable H2D copy per call
        stalls the stream and cannot be captured into a CUDA graph)."""
        if isinstance(spec, str):
            return self._convert_tps_lmbda(num_samples, spec).to(device)
        key = (num_samples, float(spec), str(device))
        hit = self._lmbda_cache.get(key)
        if hit is None:
            hit = self._lmbda_cache[key] = self._convert_tps_lmbda(num_samples, spec).to(device)
        return hit

    @staticmethod
    def is_supported_transform_type(s):
        return s in ["affine", "rigid"] or bool(re.match(r"^tps_.*$", s))

    def _make_aligner(self, align_type, points_m, points_f, weights, tps_lmbda, aff_f=None,
                      aff_m=None, shape_f=None, shape_m=None, real_world=False, fit_forward=False,
                      fit_inverse=True):
        common = dict(points_m=points_m, points_f=points_f, w=weights, aff_f=aff_f, aff_m=aff_m,
                      shape_f=shape_f, shape_m=shape_m, dim=self.dim,
                      align_in_real_world_coords=real_world)
        if align_type == "rigid":
            return RigidKeypointAligner(**common)
        if align_type == "affine":
            return AffineKeypointAligner(**common)
        return TPS(lmbda=tps_lmbda, use_checkpoint=self.use_checkpoint, fit_forward=fit_forward,
                   fit_inverse=fit_inverse, **common)

    # ------------------------------------------------------------------ pairwise
    @torch.no_grad()
    def forward(self, img_f, img_m, transform_type="affine", **kwargs):
        """keymorph/model.py:142-289.  Returns {align_str: {grid, points_f, points_m,
        points_weights, tps_lmbda, time*, [matrix], [points_a]}}; unknown kwargs (seg_f, save_dir,
        num_resolutions_for_itkelastix, ...) are tolerated like in the reference."""
        return_aligned_points = kwargs["return_aligned_points"]
        if not isinstance(transform_type, (list, tuple)):
            transform_type = [transform_type]
        assert all(self.is_supported_transform_type(s) for s in transform_type), \
            "Invalid transform_type"
        if self.cuda_graph:
            key = self._graph_key(img_f, img_m, transform_type, kwargs)
            if key is not None:
                return self._forward_graphed(key, img_f, img_m, list(transform_type), kwargs)
        return self._forward_eager(img_f, img_m, transform_type, None, kwargs)

    def _forward_eager(self, img_f, img_m, transform_type, status_sink, kwargs, batch=None):
        """`batch`: img_f and img_m are the two halves of this tensor already (graph replay stages its inputs
        there), so the concatenation below is skipped."""
        return_aligned_points = kwargs["return_aligned_points"]
        if self.align_keypoints_in_real_world_coords:
            aff_f, aff_m = kwargs["aff_f"], kwargs["aff_m"]
            shape_m = torch.tensor(img_m.shape[2:]).to(img_m)
            shape_f = torch.tensor(img_f.shape[2:]).to(img_f)
        else:
            aff_f = aff_m = shape_f = shape_m = None
        assert img_f.shape[1] == 1, "Image dimension must be 1"
        assert img_m.shape[1] == 1, "Image dimension must be 1"

        start_time = time.time()
        nb = img_f.shape[0]
        if img_f.shape == img_m.shape and backbone_engine(self.backbone) is not None:
            # both volumes through the backbone as one batch (fills the small pyramid levels)
            pts, mass, _ = self._keypoints_and_mass(torch.cat([img_f, img_m], 0) if batch is None else batch)
            points_f, points_m = pts[:nb], pts[nb:]
            mass_f, mass_m = mass[:nb], mass[nb:]
        else:
            points_f, mass_f, _ = self._keypoints_and_mass(img_f)
            points_m, mass_m, _ = self._keypoints_and_mass(img_m)
        if self.weight_keypoints == "power":
            # keymorph/model.py:95-109: sum relu(feat) is exactly the CoM kernel's total mass
            weights = mass_f * mass_m
            weights = weights / weights.sum(dim=1, keepdim=True)
        else:
            weights = None
        keypoint_extract_time = time.time() - start_time

        result_dict = {}
        with deferred_singular_checks(status_sink):   # one read of the status flags, after the last launch
            self._align_all(result_dict, transform_type, img_f, img_m, points_f, points_m, weights, aff_f,
                            aff_m, shape_f, shape_m, return_aligned_points, keypoint_extract_time, kwargs)
        return result_dict

    # ------------------------------------------------------------------ CUDA-graph replay of forward()
    def _graph_key(self, img_f, img_m, transform_type, kwargs):
        """Key of the captured graph this call can replay, or None when the call has to run eagerly: the graph
        holds one fixed launch sequence over static buffers, so per-call tensors other than the two images
        (segmentations, label maps, real-world affines), random TPS lambdas and training mode are excluded."""
        if not (img_f.is_cuda and img_m.is_cuda and img_f.device == img_m.device and img_f.shape == img_m.shape
                and img_f.dtype == torch.float32 and img_m.dtype == torch.float32
                and not self.training and not self.align_keypoints_in_real_world_coords):
            return None
        for v in kwargs.values():
            if torch.is_tensor(v) or isinstance(v, np.ndarray):
                return None
        for s in transform_type:
            if s.startswith("tps") and isinstance(str_or_float(s[4:]), str):
                return None
        return (tuple(img_f.shape), img_f.device.index, tuple(transform_type),
                bool(kwargs["return_aligned_points"]))

    def _graph_signature(self):
        """Everything a captured launch sequence bakes in besides the key: parameter storage / versions (packed
        weights are rebuilt when they change) and the engine's schedule switches."""
        params = tuple((p.data_ptr(), p._version) for p in self.backbone.parameters())
        return (params, str(ops.act_dtype()), ops.USE_GN_FOLD, ops.gn_fold_stem_enabled(), ops.USE_GN_FOLD_TC_PAIR,
                ops.USE_PAIR_CONV, ops.USE_ZFOLD_PAIR, ops.zfold_pair_cin32_enabled(), ops.USE_COARSE_UPCONV,
                self.fused_warp, self.weight_keypoints)

    def _forward_graphed(self, key, img_f, img_m, transform_type, kwargs):
        """First call of a key: eager (packs weights, sets kernel attributes, fills every cache).  Second call:
        the same launch sequence is captured into a torch.cuda.CUDAGraph over static input buffers.  From then
        on: two device copies into the static inputs + one graph launch.  The returned tensors are the graph's
        static outputs: they are overwritten by the next forward() with the same key (clone what must outlive
        it), and the `time*` entries are those of the capture.  Singular-fit flags cannot be read inside a
        graph; they are checked after the replay (one small D2H read, as in the eager path)."""
        sig = self._graph_signature()
        ent = self._graphs.get(key)
        if ent is None or ent["sig"] != sig:
            self._graphs.pop(key, None)
            while len(self._graphs) >= max(1, self.max_graphs):      # evict the oldest capture (frees its pool)
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = {"sig": sig, "graph": None}
            return self._forward_eager(img_f, img_m, transform_type, None, kwargs)
        if ent["graph"] is None:
            if ent.get("failed"):
                return self._forward_eager(img_f, img_m, transform_type, None, kwargs)
            nb = img_f.shape[0]
            batch = torch.cat([img_f, img_m], 0)      # static input: the backbone's two-volume batch itself
            static_f, static_m = batch[:nb], batch[nb:]
            status = []
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(img_f.device)
            try:
                # thread_local: other threads' CUDA calls (NCCL watchdog, prefetch threads) must not abort the capture
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = self._forward_eager(static_f, static_m, transform_type, status, kwargs, batch=batch)
                    flags = torch.stack([st.reshape(-1).any() for st, _ in status]) if status else None
            except Exception as exc:        # a launch the runtime refuses to capture: stay eager, say why once
                ent["failed"] = f"{type(exc).__name__}: {exc}"
                import warnings
                warnings.warn(f"keymorph_b200: CUDA-graph capture of forward{key} failed ({ent['failed']}); "
                              "running eagerly")
                torch.cuda.synchronize(img_f.device)
                return self._forward_eager(img_f, img_m, transform_type, None, kwargs)
            ent.update(graph=graph, out=out, static_f=static_f, static_m=static_m, status=status, flags=flags)
        _stage(ent["static_f"], img_f)
        _stage(ent["static_m"], img_m)
        ent["graph"].replay()
        from . import transformations
        if ent["flags"] is not None and transformations.CHECK_SINGULAR:
            for bad, (st, what) in zip(ent["flags"].cpu().tolist(), ent["status"]):
                if bad:
                    raise torch.linalg.LinAlgError(f"{what}: the matrix is singular (status={st.tolist()})")
        return ent["out"]

    def graph_state(self):
        """{key: "captured" | "eager (<why>)" | "warm-up"} of the CUDA-graph cache (diagnostics, bench.py)."""
        return {k: ("captured" if e["graph"] is not None else
                    (f"eager ({e['failed']})" if e.get("failed") else "warm-up")) for k, e in self._graphs.items()}

    def _align_all(self, result_dict, transform_type, img_f, img_m, points_f, points_m, weights, aff_f, aff_m,
                   shape_f, shape_m, return_aligned_points, keypoint_extract_time, kwargs):
        for align_type_str in transform_type:
            start_time = time.time()
            if align_type_str.startswith("tps"):
                align_type = "tps"
                tps_lmbda = self._tps_lmbda_on(len(img_f), str_or_float(align_type_str[4:]), img_f.device)
            else:
                align_type, tps_lmbda = align_type_str, None
            aligner = self._make_aligner(align_type, points_m, points_f, weights, tps_lmbda, aff_f,
                                         aff_m, shape_f, shape_m,
                                         self.align_keypoints_in_real_world_coords,
                                         fit_forward=bool(return_aligned_points))
            fused = None
            if self.fused_warp and align_type in ("rigid", "affine") and img_m.shape == img_f.shape:
                # one pass generates the flow field, warps the moving image and takes the loss sums
                img_a, sums, grid = ops.warp_loss(img_m, img_f, mat34=aligner._grid_matrix(),
                                                  want_grid=True)
                fused = (img_a, sums)
            else:
                grid = aligner.get_flow_field(img_f.shape, compute_on_subgrids=not self.training)
            if return_aligned_points:
                points_a = aligner.get_forward_transformed_points(points_m)
            align_time = time.time() - start_time
            res = {
                "grid": grid,
                "points_f": points_f,
                "points_m": points_m,
                "points_weights": weights,
                "tps_lmbda": tps_lmbda,
                "time_keypoint_extract": keypoint_extract_time,
                "time_align": align_time,
                "time": keypoint_extract_time + align_time,
            }
            if align_type in ["rigid", "affine"]:
                res["matrix"] = aligner.transform_matrix
            if return_aligned_points:
                res["points_a"] = points_a
            if self.fused_warp:
                self._fused_outputs(res, grid, img_f, img_m, kwargs, fused)
            result_dict[align_type_str] = res

    def _fused_outputs(self, res, grid, img_f, img_m, kwargs, fused=None):
        """Warped image / segmentation and loss sums in the same pass that reads (or, for rigid /
        affine, writes) the grid (scripts/pairwise_register_eval.py:137-162,303-321 would otherwise
        call align_img and the losses separately)."""
        img_a, sums = fused if fused is not None else ops.warp_loss(img_m, img_f, grid=grid)
        res["img_a"] = img_a
        res["mse"] = (sums[..., 0].sum() / img_f.numel()).float()
        seg_f, seg_m = kwargs.get("seg_f"), kwargs.get("seg_m")
        lab_f, lab_m = kwargs.get("labels_f"), kwargs.get("labels_m")
        if lab_f is not None and lab_m is not None:
            # label-map fast path: integer label volumes instead of fp32 one-hot
            # channels; same soft / hard Dice as one_hot -> align_img -> DiceLoss
            soft, hard, lab_a = ops.warp_labels_dice(lab_m, lab_f, kwargs["num_classes"], grid=grid,
                                                     want_labels=True)
            res["labels_a"] = lab_a.reshape(lab_m.shape)
            res["softdice"] = dice_from_sums(soft)
            res["harddice"] = dice_from_sums(hard)
        if seg_f is not None and seg_m is not None:
            seg_a, ssums = ops.warp_loss(seg_m.float(), seg_f.float(), grid=grid)
            res["seg_a"] = seg_a
            res["softdice"] = dice_from_sums(ssums)

    def pairwise_register(self, *args, **kwargs):
        """Alias for forward() (the reference's alias, model.py:291-293, passes `self` twice and
        cannot be called; this one works)."""
        return self.forward(*args, **kwargs)

    # ------------------------------------------------------------------ groupwise
    def _groupwise_step(self, group_points, align_type, lmbda):
        """keymorph/model.py:331-394 with all subjects fitted in ONE batched kernel call."""
        mean_points = torch.mean(group_points, dim=0, keepdim=True)
        return self._register_to_mean(group_points, mean_points, align_type, lmbda), mean_points

    def _register_to_mean(self, group_points, mean_points, align_type, lmbda):
        G = group_points.shape[0]
        fixed = mean_points.expand(G, -1, -1).contiguous()
        lam = None if lmbda is None else lmbda.reshape(-1)[:1].repeat(G)
        aligner = self._make_aligner(align_type, group_points, fixed, None, lam, fit_forward=True,
                                     fit_inverse=False)
        return aligner.get_forward_transformed_points(group_points)

    @torch.no_grad()
    def groupwise_register(self, inputs, transform_type="affine", **kwargs):
        """keymorph/model.py:295-530.  `inputs`: directory of *.npz (key "img"), list of paths, or a
        tensor stack (N,1,D,H,W).  With save_results_to_disk and a directory input the grids are
        written as {save_dir}/{align}_grid_{i:03}.npy; otherwise they are returned as
        res["groupgrids"] (the reference's in-memory branch is broken)."""
        device = kwargs["device"]
        num_iters = kwargs["num_iters"]
        log_to_console = kwargs["log_to_console"]
        if isinstance(transform_type, str):
            transform_type = [transform_type]
        if isinstance(inputs, str):
            save_dir = kwargs["save_dir"]
            inputs = sorted(os.path.join(inputs, f) for f in os.listdir(inputs) if f.endswith(".npz"))
            if len(inputs) == 0:
                raise ValueError("No .npz files found")
        else:
            save_dir = None

        group_points = []
        img_m = None
        for i in range(len(inputs)):
            if isinstance(inputs[i], str):
                img_m = torch.tensor(np.load(inputs[i])["img"]).float()
            else:
                img_m = inputs[i:i + 1]
            img_m = img_m.to(device)
            group_points.append(self.get_keypoints(img_m).detach())
            if log_to_console:
                print(f"-> Extracted keypoints from subject {i+1}/{len(inputs)}")
        group_points = torch.cat(group_points, dim=0)

        result_dict = {}
        for align_type_str in transform_type:
            start_time = time.time()
            if align_type_str.startswith("tps"):
                align_type = "tps"
                tps_lmbda = self._tps_lmbda_on(len(img_m), str_or_float(align_type_str[4:]), img_m.device)
            else:
                align_type, tps_lmbda = align_type_str, None
            curr_points = group_points.clone()
            mean_points = None
            with deferred_singular_checks():     # one read of the fits' status flags per transform
                for j in range(num_iters):
                    curr_points, mean_points = self._groupwise_step(curr_points, align_type, tps_lmbda)
                    if log_to_console:
                        print(f"-> Iteration {j+1}/{num_iters}")
            res = {"time": time.time() - start_time, "grouppoints_m": group_points,
                   "grouppoints_a": curr_points}
            # grids: ORIGINAL points against the mean taken at the start of the last iteration
            grids = []
            for i in range(len(group_points)):
                pm = group_points[i:i + 1]
                lam = None if tps_lmbda is None else tps_lmbda.reshape(-1)[:1]
                aligner = self._make_aligner(align_type, pm, mean_points, None, lam)
                grid = aligner.get_flow_field(img_m.shape, compute_on_subgrids=True)
                if kwargs["save_results_to_disk"] and save_dir:
                    np.save(f"{save_dir}/{align_type_str}_grid_{i:03}.npy", grid.cpu().numpy())
                else:
                    grids.append(grid)
            if grids:
                res["groupgrids"] = torch.cat(grids, dim=0)
            result_dict[align_type_str] = res
        return result_dict
