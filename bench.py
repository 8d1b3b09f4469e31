"""Benchmark of the pairwise registration hot path (BASELINE.json: "pairwise 256^3 registrations/sec at
1/2/4/8 B200; grid_sample GB/s vs HBM peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload = BASELINE configs[1]: synthetic 256^3 pair, affine, 256 keypoints, TruncatedUNet3D, bf16.
One step = one pairwise registration: two 256^3 fp32 volumes -> backbone on both -> centre-of-mass keypoints
-> fit -> flow field -> warped moving image + MSE.
  value        registrations/s, volumes resident in HBM (CUDA events around the K steps, max over ranks), through
               KeyMorph(..., fused_warp=True, cuda_graph=True).forward: the call replays the CUDA graph it captured
               on its second invocation (identical kernels; `execution` reports the eager number beside it, and
               --eager or a failed capture makes the eager path the headline)
  e2e          the same call fed from pinned HOST memory: H2D of both volumes and D2H of the warped image
               + MSE of every step inside the timed region
  roofline     the tcgen05 convolution kernels (tensor bound), timed live with CUDA events around every C-ABI call
               of an EAGER pass of the same model; flops = the MMAs executed; peak = the burst figure of
               MEASURED_PEAKS.json when the timed region is shorter than 1 s, else the sustained one
  tps_config3  BASELINE configs[2] (the north-star target: 256^3 pair, TPS lambda=0, 512 keypoints): value,
               e2e and the roofline of the dense TPS field kernel (issue / MUFU bound) and of the fused warp
  gpu_baseline the UNMODIFIED reference (oracle/_ref, stock torch CUDA ops: cuDNN conv3d, ATen grid_sampler_3d,
               host LAPACK for TPS) on the same GPU, fp32 and its own AMP mode (fp16 autocast), both configs
  groupwise    BASELINE configs[4]: 32 synthetic 256^3 subjects, tps_0, K=512, 5 iterations, sharded over the
               ranks; mean keypoints by NCCL all-reduce (5 collectives) and by one all-gather + redundant
               local iteration (1 collective), both timed, results compared
  cpu_baseline the reference itself (oracle/_ref; the oracle port only if that is absent) on the host cores
--impl reference times the reference's own CPU implementation (rank 0 only) on the headline workload, full
256^3 volumes, every step a complete registration.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S = 256
METRIC = "pairwise 256^3 registrations/sec"
UNIT = "registrations/s"
CONFIGS = {
    "affine": {"K": 256, "transform": "affine",
               "workload": "synthetic 256^3 pair, affine, 256 keypoints, TruncatedUNet3D(levels 4, truncated 1), "
                           "fp16 operands / fp32 accumulate (tcgen05 kind::f16; the BASELINE-named bf16 operand "
                           "type is measured beside it in bf16_operands)"},
    "tps": {"K": 512, "transform": "tps_0",
            "workload": "synthetic 256^3 pair, TPS lambda=0, 512 keypoints, TruncatedUNet3D(levels 4, truncated 1), "
                        "fp16 operands / fp32 accumulate"},
}
GROUP_SUBJECTS, GROUP_ITERS = 32, 5
# dram__bytes_read.sum + dram__bytes_write.sum per step, parsed from the committed ncu launch list of this
# workload (tools/summarize_traffic.py output); None when the profile is missing
TRAFFIC_PROFILE = os.path.join(ROOT, "profiles", "r02_ncu_launches_dram_traffic.txt")


def config_dict(world):
    return {"workload": CONFIGS["affine"]["workload"], "pairs_per_step_per_gpu": 1,
            "parallelism": f"pairs sharded x{world}, no collective on the data path",
            "l2": "per-step working set (>5 GB of activations) exceeds the 126 MB L2"}


def conv_layers(S, K, n_img, executed=True):
    """(name, Cin, Cout, edge, taps, FLOPs) of every conv on the path, keymorph/unet3d/buildingblocks.py
    :171-181 channel rule; FLOPs = 2*taps*Cin*Cout*edge^3 per image."""
    L = [("enc0.c1", 1, 16, S, 27), ("enc0.c2", 16, 32, S, 27), ("enc1.c1", 32, 32, S // 2, 27),
         ("enc1.c2", 32, 64, S // 2, 27), ("enc2.c1", 64, 64, S // 4, 27), ("enc2.c2", 64, 128, S // 4, 27),
         ("enc3.c1", 128, 128, S // 8, 27), ("enc3.c2", 128, 256, S // 8, 27),
         ("dec0.c1", 384, 128, S // 4, 27), ("dec0.c2", 128, 128, S // 4, 27),
         ("dec1.c1", 192, 64, S // 2, 27), ("dec1.c2", 64, 64, S // 2, 27), ("final", 64, K, S // 2, 1)]
    from keymorph_b200 import ops
    if executed and ops.USE_COARSE_UPCONV and ops.USE_GN_FOLD and ops.USE_ZFOLD_PAIR:
        # EXECUTED flops: the upsampled channels of the decoders' first convs run as 8 pre-summed taps per output
        # voxel on the coarse lattice (conv_up2.cu), the skip channels as the usual 27
        i = [l[0] for l in L].index("dec1.c1")
        L[i:i + 1] = [("dec1.c1", 64, 64, S // 2, 27), ("dec1.c1.up", 128, 64, S // 2, 8)]
        if ops.USE_PAIR_CONV:
            i = [l[0] for l in L].index("dec0.c1")
            L[i:i + 1] = [("dec0.c1", 128, 128, S // 4, 27), ("dec0.c1.up", 256, 128, S // 4, 8)]
    return [(n, ci, co, e, t, 2.0 * t * ci * co * e ** 3 * n_img) for (n, ci, co, e, t) in L]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tc_burst": p["bf16_tflops"], "tc_sustained": p["bf16_tflops_sustained"],
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:  # noqa: BLE001
        return {"hbm": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1400.0, "sm_max_mhz": 1965.0,
                "src": "fallback (B200_PROFILING.md)"}


def traffic_from_profile():
    """{kernel name: DRAM bytes per step} from the committed ncu launch list (None if absent)."""
    out = {}
    try:
        with open(TRAFFIC_PROFILE) as f:
            for line in f:
                p = line.split()
                if len(p) >= 6 and p[1].endswith("%") and p[2].startswith("x"):
                    name = p[5].split("<")[0]          # template arguments dropped: instantiations are summed
                    out[name] = out.get(name, 0.0) + (float(p[3]) + float(p[4])) * 1e6
    except OSError:
        return None
    return out or None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# synthetic inputs: Gaussian-blob phantoms; moving = a second phantom under a fixed affine
def host_pair(rank):
    import torch
    from oracle import keymorph_oracle as O          # phantom generator only (outside every timed region)
    img_f = O.gaussian_phantom(S, 1000 + rank)
    base_m = O.gaussian_phantom(S, 2000 + rank)
    return img_f, base_m, torch.inverse(O.affine_matrix_3d(0.1, 0.05, 0.3, 0.02))


def device_phantom(size, seed, dev, n_blobs=12):
    """oracle.gaussian_phantom's recipe evaluated on the device (fp32, separable outer products)."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    ctr = (torch.rand(n_blobs, 3, generator=gen, dtype=torch.float64) * 1.2 - 0.6).float().to(dev)
    sig = (torch.rand(n_blobs, generator=gen, dtype=torch.float64) * 0.25 + 0.1).float().to(dev)
    amp = torch.rand(n_blobs, generator=gen, dtype=torch.float64).float().to(dev)
    lin = torch.linspace(-1, 1, size, device=dev)
    g = torch.exp(-0.5 * ((lin[None, None, :] - ctr[:, :, None]) / sig[:, None, None]) ** 2)     # (B,3,S)
    vol = torch.einsum("b,bz,by,bx->zyx", amp, g[:, 0], g[:, 1], g[:, 2])
    return (vol / vol.max())[None, None].contiguous()


def seeded_backbone(K):
    import torch
    import keymorph_b200 as kb
    torch.manual_seed(23)
    return kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                              num_levels=4, is_segmentation=False, conv_padding=1).eval()


# ------------------------------------------------------------------------------------------------------
# reference arms (CPU): the reference itself from oracle/_ref, else the oracle port
def reference_step_fn(name, device="cpu", amp=False):
    """-> (callable(img_f, img_m) -> mse, kind) running one complete registration of config `name`
    through the UNMODIFIED reference (keymorph/model.py:142-289, utils.py:14-21, loss_ops.py:9-14)."""
    import torch
    from oracle import refshim
    cfg = CONFIGS[name]
    if refshim.available():
        refshim.import_reference()
        from keymorph.keypoint_aligners import TPS
        from keymorph.loss_ops import MSELoss
        from keymorph.utils import align_img
        model = refshim.build_reference_model(cfg["K"], device=device, use_amp=amp)
        mse_fn = MSELoss()

        def step(img_f, img_m):
            with torch.no_grad():
                if name == "tps" and device == "cpu":
                    # the stock forward() builds TPS(num_subgrids=4): one K x M x 3 fp32 temporary of 25.8 GB,
                    # several alive; BASELINE.md section 4.3: same per-voxel arithmetic with 64 sub-grids
                    pf, pm = model.get_keypoints(img_f), model.get_keypoints(img_m)
                    al = TPS(pm, pf, torch.tensor([0.0]), num_subgrids=64)
                    grid = al.get_flow_field(img_f.shape, compute_on_subgrids=True)
                    al.get_forward_transformed_points(pm)
                else:
                    grid = model(img_f, img_m, transform_type=cfg["transform"], return_aligned_points=True)[
                        cfg["transform"]]["grid"]
                return mse_fn(align_img(grid, img_m), img_f)
        step.model = model
        return step, "reference"
    from oracle import keymorph_oracle as O
    sd = {k: v.clone() for k, v in seeded_backbone(cfg["K"]).state_dict().items()}

    def step(img_f, img_m):
        with torch.no_grad():
            r = O.keymorph_forward("truncatedunet", sd, img_f, img_m, cfg["transform"])[cfg["transform"]]
            return O.mse_loss(O.align_img(r["grid"], img_m), img_f)
    return step, "port"


def cpu_pair():
    import torch
    from oracle import keymorph_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    img_f, base_m, minv = host_pair(0)
    img_m = O.align_img(O.affine_flow_field(minv, (S, S, S)), base_m)
    return img_f, img_m


def run_reference(args):
    """--impl reference: the reference's CPU implementation, all host threads, full 256^3 volumes,
    warmup + steps complete registrations of the headline workload; then ONE registration of config 3."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    img_f, img_m = cpu_pair()
    step, kind = reference_step_fn("affine")
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        mse = float(step(img_f, img_m))
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.mean(times)
    value = 1.0 / sec
    sample = (f"{args.steps} complete 256^3 registrations (2 backbone passes, CoM, affine fit, flow field, "
              f"align_img, MSE), torch {torch.__version__} CPU fp32, {torch.get_num_threads()} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world), "device": "cpu", "mse": mse,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_tps:
        try:
            tstep, tkind = reference_step_fn("tps")
            t0 = time.perf_counter()
            tm = float(tstep(img_f, img_m))
            dt = time.perf_counter() - t0
            line["tps_config3"] = {"value": 1.0 / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "kind": tkind, "mse": tm,
                                   "workload": CONFIGS["tps"]["workload"],
                                   "sample": "ONE complete registration, no warm-up; TPS(num_subgrids=64) so that the "
                                             "reference's K x M x 3 temporaries fit in host RAM (BASELINE.md 4.3)"}
        except Exception as e:  # noqa: BLE001
            line["tps_config3"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    emit(line)


# ------------------------------------------------------------------------------------------------------
class Tracer:
    """CUDA events around chosen C-ABI calls on the launching stream (installed as _lib.TRACE)."""

    def __init__(self, names, stream):
        import torch
        self.torch = torch
        self.t = {n: [] for n in names}
        self.pending = {}
        self.stream = stream

    def __call__(self, name, phase):
        if name in self.t:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record(self.stream)
            if phase == "pre":
                self.pending[name] = ev
            else:
                self.t[name].append((self.pending.pop(name), ev))

    def ms(self, *names, per=1):
        return sum(a.elapsed_time(b) for n in names for a, b in self.t.get(n, [])) / per

    def count(self, *names, per=1):
        return sum(len(self.t.get(n, [])) for n in names) // per


CONV_CALLS = ("km_conv3d_tc", "km_conv3d_tc_pair", "km_conv3d_zfold_pair", "km_conv3d_zfold_pair_gn",
              "km_conv3d_zfold_pair_gn_cat", "km_conv3d_tc_pair_gn", "km_conv3d_up2_gn", "km_conv3d_zfold_pair_gn_add", "km_conv3d_tc_pair_gn_add")
ZF_CALLS = ("km_conv3d_zfold", "km_conv3d_zfold_gn")
TRACED = CONV_CALLS + ZF_CALLS + ("km_conv1x1_com", "km_conv3d_stem", "km_warp_loss", "km_flow_field_tps", "km_tps_fit")


def measure_pairwise(model, name, img_f, img_m, steps, warmup, ctx):
    """Device-resident timing of config `name`: -> dict(ms_step (max over ranks), tracer, launches, mse)."""
    import torch
    from keymorph_b200 import _lib
    t = CONFIGS[name]["transform"]

    def step():
        return model(img_f, img_m, transform_type=t, return_aligned_points=True)[t]

    for _ in range(warmup):
        step()
    ctx["sync"]()
    tr = Tracer(TRACED, torch.cuda.current_stream())
    l0 = _lib.launch_count
    _lib.TRACE = tr
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = step()
    e1.record()
    ctx["sync"]()
    _lib.TRACE = None
    return {"ms_step": ctx["max"](e0.elapsed_time(e1) / steps), "tr": tr, "launches": _lib.launch_count - l0,
            "mse": float(r["mse"].item()), "points": torch.cat([r["points_f"], r["points_m"]]).cpu()}


def measure_e2e(model, name, host_f, host_m, steps, ctx, with_grid=False, clone_outputs=False):
    """The public call fed from pinned HOST buffers.  Every step: H2D of both volumes (prefetched on a side
    stream), D2H of the warped image (+ the flow field with with_grid) and of the MSE into pinned host
    memory on a second side stream; the host reads step i's results after step i+1 has been enqueued."""
    import torch
    from keymorph_b200.hostio import prefetch_to_device
    dev = ctx["dev"]
    t = CONFIGS[name]["transform"]
    d2h = torch.cuda.Stream(dev)
    compute = torch.cuda.current_stream(dev)
    ring = [{"img": torch.empty((1, 1, S, S, S), dtype=torch.float32).pin_memory(),
             "grid": torch.empty((1, S, S, S, 3), dtype=torch.float32).pin_memory() if with_grid else None,
             "mse": torch.empty((), dtype=torch.float32).pin_memory(), "ev": torch.cuda.Event(), "keep": None}
            for _ in range(2)]

    def run(n):
        pending, last = None, None
        for i, (f, m) in enumerate(prefetch_to_device([(host_f, host_m)] * n, dev)):
            r = model(f, m, transform_type=t, return_aligned_points=True)[t]
            if clone_outputs:     # a replayed CUDA graph overwrites its static outputs at the next call
                r = {k: r[k].clone() for k in (("img_a", "mse", "grid") if with_grid else ("img_a", "mse"))}
            slot = ring[i & 1]
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                slot["img"].copy_(r["img_a"], non_blocking=True)
                if with_grid:
                    slot["grid"].copy_(r["grid"], non_blocking=True)
                slot["mse"].copy_(r["mse"], non_blocking=True)
                slot["ev"].record(d2h)
            slot["keep"] = r                      # the device tensors stay alive until their copy has run
            if pending is not None:
                ring[pending]["ev"].synchronize()
                last = (float(ring[pending]["mse"]), float(ring[pending]["img"][0, 0, S // 2, S // 2, S // 2]))
                ring[pending]["keep"] = None
            pending = i & 1
        ring[pending]["ev"].synchronize()
        last = (float(ring[pending]["mse"]), float(ring[pending]["img"][0, 0, S // 2, S // 2, S // 2]))
        torch.cuda.synchronize()
        return last

    run(2)
    ctx["sync"]()
    t0 = time.perf_counter()
    last = run(steps)
    ms = ctx["max"]((time.perf_counter() - t0) * 1e3 / steps)
    d2h_bytes = S ** 3 * 4 + 4 + (S ** 3 * 12 if with_grid else 0)
    return {"value": ctx["world"] * 1e3 / ms, "unit": UNIT, "ms_per_step": ms,
            "h2d_bytes_per_step": 2 * S ** 3 * 4, "d2h_bytes_per_step": d2h_bytes,
            "d2h": "warped image (img_a) + MSE" + (" + flow field" if with_grid else
                                                   "; the flow field stays on the device (it is a function of the "
                                                   "returned matrix / spline parameters; --e2e-grid copies it too)"),
            "loss": last[0], "img_a_center": last[1]}


def measure_graph(gm, name, img_f, img_m, steps, warmup, ctx):
    """Device-resident timing of forward() replayed from its CUDA graph (KeyMorph(cuda_graph=True): call 1 eager,
    call 2 captures, every later call = 2 staging copies + 1 graph launch)."""
    import torch
    t = CONFIGS[name]["transform"]
    for _ in range(max(warmup, 3)):
        r = gm(img_f, img_m, transform_type=t, return_aligned_points=True)[t]
    state = "; ".join(gm.graph_state().values())
    ctx["sync"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = gm(img_f, img_m, transform_type=t, return_aligned_points=True)[t]
    e1.record()
    ctx["sync"]()
    captured = ctx["max"](0.0 if state == "captured" else 1.0) == 0.0      # on every rank
    return {"ms_step": ctx["max"](e0.elapsed_time(e1) / steps), "state": state, "captured": captured,
            "mse": float(r["mse"].item())}


def conv_rooflines(tr, name, steps, ms_step, peaks, timed_s):
    """Rooflines of the backbone kernels of one step from the traced launches."""
    K = CONFIGS[name]["K"]
    burst = timed_s < 1.0
    tc_peak = peaks["tc_burst"] if burst else peaks["tc_sustained"]
    which = "burst" if burst else "sustained"
    layers = {l[0]: l[5] for l in conv_layers(S, K, 2)}
    alg = {l[0]: l[5] for l in conv_layers(S, K, 2, executed=False)}
    alg_flops = sum(v for k, v in alg.items() if k not in ("enc0.c1", "enc0.c2", "final"))
    conv_ms, zf_ms = tr.ms(*CONV_CALLS, per=steps), tr.ms(*ZF_CALLS, per=steps)
    com_ms, stem_ms = tr.ms("km_conv1x1_com", per=steps), tr.ms("km_conv3d_stem", per=steps)
    zf_flops = layers["enc0.c2"] if zf_ms > 0 else 0.0
    com_flops = layers["final"] if com_ms > 0 else 0.0
    tc_flops = sum(v for k, v in layers.items() if k != "enc0.c1") - zf_flops - com_flops
    traffic = traffic_from_profile() or {}
    conv_traffic = sum(traffic.get(k, 0.0) for k in ("conv_tc_kernel", "conv_tc2_kernel", "conv_zf2_kernel", "conv_up2_kernel")) or None
    ach = tc_flops / (conv_ms * 1e-3) / 1e12
    main = {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc2_kernel / conv_zf2_kernel / conv_up2_kernel (tcgen05 3x3x3 convolutions; EXECUTED flops)",
            "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak, "traffic": conv_traffic,
            "flops_per_step": tc_flops, "algorithmic_flops_per_step": alg_flops,
            "algorithmic_tflops": alg_flops / (conv_ms * 1e-3) / 1e12,
            "flops_note": "achieved / frac count the MMAs EXECUTED; the reference's 27-tap arithmetic on the upsampled "
                          "decoder channels (algorithmic_flops) is done with 8 pre-summed taps on the coarse lattice", "launches_per_step": tr.count(*CONV_CALLS, per=steps),
            "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / ms_step,
            "peak_source": f"{peaks['src']}: {which} bf16 figure (timed region {timed_s:.2f} s)",
            "frac_of_sustained": ach / peaks["tc_sustained"], "frac_of_burst": ach / peaks["tc_burst"],
            "traffic_note": "DRAM read+write bytes of the same launches in one step, parsed from the committed ncu "
                            "launch list profiles/r02_ncu_launches_dram_traffic.txt (null if absent)"}
    other = []
    if zf_ms > 0:
        a = zf_flops / (zf_ms * 1e-3) / 1e12
        other.append({"bound": "tensor", "kernel": "conv_zf_kernel (16->32 @256^3, dz folded into N, pool fused)",
                      "achieved": a, "peak": tc_peak, "unit": "TFLOP/s", "frac": a / tc_peak, "kernel_ms_per_step": zf_ms,
                      "traffic": traffic.get("conv_zf_kernel")})
    if com_ms > 0:
        a = com_flops / (com_ms * 1e-3) / 1e12
        other.append({"bound": "tensor", "kernel": "com_tc_kernel (final 1x1x1 conv + ReLU + centre of mass)",
                      "achieved": a, "peak": tc_peak, "unit": "TFLOP/s", "frac": a / tc_peak, "kernel_ms_per_step": com_ms,
                      "hbm_gbs": 2 * (S // 2) ** 3 * 64 * 2 / (com_ms * 1e-3) / 1e9, "traffic": traffic.get("com_tc_kernel")})
    if stem_ms > 0:
        # the 1 -> 16 stem is HBM-bound by construction (24 FLOP/B): 4 B/voxel read per pass + 32 B/voxel written once
        nst = tr.count("km_conv3d_stem", per=steps)      # passes over the batch of two volumes
        b = 2 * S ** 3 * (4.0 * nst + 32.0) if nst else 0.0
        other.append({"bound": "hbm", "kernel": f"conv_stem_mma_kernel x{nst} (1->16 @256^3, mma.sync TF32)",
                      "achieved": b / (stem_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                      "frac": b / (stem_ms * 1e-3) / 1e9 / peaks["hbm"], "kernel_ms_per_step": stem_ms,
                      "traffic": traffic.get("conv_stem_mma_kernel")})
    bb_ms = conv_ms + zf_ms + com_ms + stem_ms
    allf = sum(layers.values())
    other.append({"bound": "tensor", "kernel": "whole backbone (stem + conv_zf + conv_tc* + com_tc)",
                  "achieved": allf / (bb_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                  "frac": allf / (bb_ms * 1e-3) / 1e12 / tc_peak, "kernel_ms_per_step": bb_ms})
    return main, other


def gpu_baseline(dev, img_f, img_m, steps=3, engine_points=None):
    """The reference ON THIS GPU through stock torch CUDA ops: fp32 and use_amp=True."""
    import torch
    out = {"impl": "unmodified reference (oracle/_ref) on cuda: cuDNN conv3d, ATen grid_sampler_3d, host LAPACK "
                   "solve for TPS (keymorph/keypoint_aligners.py:276-320)", "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version(), "unit": UNIT}
    try:
        from oracle import refshim
        if not refshim.available():
            return {"unavailable": "oracle/_ref missing (run oracle/build_ref.py in the build container)"}
        for name in ("affine", "tps"):
            for amp in (False, True):
                key = f"{name}_{'fp16_autocast' if amp else 'fp32'}"
                try:
                    step, _ = reference_step_fn(name, device=dev, amp=amp)
                    for _ in range(2):
                        mse = step(img_f, img_m)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        mse = step(img_f, img_m)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    out[key] = {"value": 1e3 / ms, "ms_per_step": ms, "steps": steps, "mse": float(mse),
                                "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
                    if engine_points and name in engine_points and hasattr(step, "model"):
                        # accuracy of the engine's keypoints against THIS run of the reference (fp32: the
                        # parity target; fp16 autocast: the reference's own reduced-precision mode)
                        with torch.no_grad(), torch.amp.autocast(device_type="cuda", enabled=amp, dtype=torch.float16):
                            rp = torch.cat([step.model.get_keypoints(img_f), step.model.get_keypoints(img_m)]).float().cpu()
                        if not amp:
                            out[f"{name}_reference_fp32_points"] = rp
                        for tag, pts in engine_points[name].items():
                            d = (pts - rp).abs()
                            out[key][f"keypoint_err_{tag}_vs_this_run"] = {"max": float(d.max()), "mean": float(d.mean())}
                        if amp and f"{name}_reference_fp32_points" in out:
                            d = (rp - out[f"{name}_reference_fp32_points"]).abs()
                            out[key]["own_drift_vs_reference_fp32"] = {"max": float(d.max()), "mean": float(d.mean())}
                except Exception as e:  # noqa: BLE001
                    out[key] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
                del step
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats(dev)
    except Exception as e:  # noqa: BLE001
        out["unavailable"] = f"{type(e).__name__}: {e}"[:300]
    for k in [k for k in out if k.endswith("_reference_fp32_points")]:
        del out[k]
    return out


def groupwise_block(model, ctx):
    """BASELINE configs[4]: GROUP_SUBJECTS synthetic 256^3 subjects, tps_0, K=512, GROUP_ITERS iterations,
    subjects sharded over the ranks (keymorph/model.py:295-530; scripts/groupwise_register_eval.py:378-405)."""
    import torch
    import torch.distributed as dist
    from keymorph_b200 import ops, parallel
    dev, world, rank = ctx["dev"], ctx["world"], ctx["rank"]
    mine = list(parallel.shard_range(GROUP_SUBJECTS, rank, world))
    subjects = []
    for i in mine:
        base = device_phantom(S, 3000 + i, dev)
        a = 0.02 * ((i % 7) - 3)
        from oracle import keymorph_oracle as O
        minv = torch.inverse(O.affine_matrix_3d(a, 0.5 * a, 2.0 * a, 0.0)).to(dev)[:, :3]
        subjects.append(ops.warp_loss(base, None, mat34=minv)[0])
    subjects = torch.cat(subjects)
    out = {"workload": f"{GROUP_SUBJECTS} synthetic 256^3 subjects, tps_0, 512 keypoints, {GROUP_ITERS} iterations, "
                       f"{len(mine)} subjects on rank 0 of {world}", "scaling": "strong", "n_gpus": world}
    results = {}
    for mode in ("allreduce", "allgather"):
        for rep in range(2):                 # first repetition = warm-up (allocator, NCCL channel setup)
            ctx["sync"]()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            pts = parallel.extract_keypoints(model, subjects)
            ev[1].record()
            cur, mean, ncoll = parallel.groupwise_iterate_points(model, pts, "tps_0", GROUP_ITERS, mode=mode)
            ev[2].record()
            checksum = parallel.groupwise_grids(model, pts, mean, "tps_0", subjects, consume="warp")
            ev[3].record()
            ctx["sync"]()
        t = [ctx["max"](ev[i].elapsed_time(ev[i + 1])) for i in range(3)]
        total = ctx["max"](ev[0].elapsed_time(ev[3]))
        results[mode] = (cur, mean)
        out[mode] = {"ms_total": total, "ms_keypoints": t[0], "ms_iterations": t[1], "ms_grids_and_warps": t[2],
                     "collectives": ncoll, "subjects_per_s": GROUP_SUBJECTS * 1e3 / total, "warp_checksum": checksum}
    # The all-gather scheme IS the single-process computation (torch.mean over the same (G,K,3) tensor on every
    # rank); the all-reduce scheme differs from it by the summation order of the mean only.  Keypoint-level
    # difference, and what tps_0 (cond(A) ~ 1e6 on clustered keypoints) makes of it in one subject's flow field.
    e_mean = (results["allreduce"][1] - results["allgather"][1]).abs().max()
    e_pts = (results["allreduce"][0] - results["allgather"][0]).abs().max()
    g_ar = parallel.groupwise_grids(model, pts[:1], results["allreduce"][1], "tps_0", subjects[:1])
    g_ag = parallel.groupwise_grids(model, pts[:1], results["allgather"][1], "tps_0", subjects[:1])
    out["sharded_allreduce_vs_single_process"] = {
        "mean_points_maxabs": ctx["max"](float(e_mean)), "aligned_points_maxabs": ctx["max"](float(e_pts)),
        "flow_field_maxabs_subject0": ctx["max"](float((g_ar - g_ag).abs().max())),
        "flow_field_maxabs_subject0_interior": ctx["max"](float(
            (g_ar - g_ag)[:, S // 4:3 * S // 4, S // 4:3 * S // 4, S // 4:3 * S // 4].abs().max())),
        "note": "all-gather scheme == single process bit for bit (gloo test); 0 at N = 1.  tps_0 interpolates 512 "
                "CLUSTERED keypoints exactly (random-init weights): its far field amplifies a 5e-5 change of the "
                "targets by orders of magnitude; the affine field of the same two means is compared beside it"}
    a_ar = parallel.groupwise_grids(model, pts[:1], results["allreduce"][1], "affine", subjects[:1])
    a_ag = parallel.groupwise_grids(model, pts[:1], results["allgather"][1], "affine", subjects[:1])
    out["sharded_allreduce_vs_single_process"]["affine_flow_field_maxabs_subject0"] = ctx["max"](
        float((a_ar - a_ag).abs().max()))
    del g_ar, g_ag, a_ar, a_ag
    # latency of the collective itself: (K*3+1) floats, NCCL all-reduce, 200 back-to-back calls
    if world > 1:
        buf = torch.zeros(512 * 3 + 1, device=dev)
        for _ in range(20):
            dist.all_reduce(buf)
        ctx["sync"]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            dist.all_reduce(buf)
        e1.record()
        ctx["sync"]()
        out["allreduce_6KB_us"] = ctx["max"](e0.elapsed_time(e1) / 200 * 1e3)
    else:
        out["allreduce_6KB_us"] = None
    return out


# ------------------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist

    import keymorph_b200 as kb
    from keymorph_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1 and not args.no_numa:
        # one process per GPU: keep it, and the pinned buffers it allocates from here on, on the GPU's NUMA node
        from keymorph_b200.parallel import bind_cpu_to_gpu
        numa_cpus = bind_cpu_to_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ctx = {"dev": dev, "world": world, "rank": rank, "sync": sync_all, "max": max_ranks}
    peaks = measured_peaks()

    img_f_cpu, base_m_cpu, minv = host_pair(rank)
    img_f_host = img_f_cpu.pin_memory()
    img_m = ops.warp_loss(base_m_cpu.to(dev), None, mat34=minv.to(dev)[:, :3])[0]
    img_m_host = img_m.cpu().pin_memory()
    img_f = img_f_host.to(dev)

    models = {}
    for name, cfg in CONFIGS.items():
        if name == "tps" and args.no_tps:
            continue
        models[name] = kb.KeyMorph(torch.nn.DataParallel(seeded_backbone(cfg["K"]).to(dev)), cfg["K"], 3,
                                   fused_warp=True).eval()

    clocks = ClockSampler(local)
    clocks.start()
    # ---- headline: configs[1]
    # eager pass with per-call CUDA events (kernel times for the rooflines, launch count), then the same forward()
    # replayed from its CUDA graph: the public API's cuda_graph=True mode is the headline (no launch gaps, no
    # dependence on how fast the host enqueues ~100 launches); --eager keeps the eager number as the headline
    res = measure_pairwise(models["affine"], "affine", img_f, img_m, args.steps, args.warmup, ctx)
    eager_ms = res["ms_step"]
    timed_s = eager_ms * args.steps * 1e-3
    roofline, roofline_other = conv_rooflines(res["tr"], "affine", args.steps, eager_ms, peaks, timed_s)
    gmodels = {}
    for name, m in models.items():
        gmodels[name] = kb.KeyMorph(m.backbone, CONFIGS[name]["K"], 3, fused_warp=True, cuda_graph=True).eval()
    gres = None if args.eager else measure_graph(gmodels["affine"], "affine", img_f, img_m, args.steps, args.warmup, ctx)
    use_graph = gres is not None and gres["captured"]
    ms_step = gres["ms_step"] if use_graph else eager_ms
    value = world * 1e3 / ms_step
    warp_ms = res["tr"].ms("km_warp_loss", per=args.steps)
    warp_bytes = 24.0 * S ** 3     # grid written 12 + moving 4 + fixed 4 + warped written 4
    traffic = traffic_from_profile() or {}
    roofline_warp = {"bound": "hbm", "kernel": "fused affine warp + flow-field store + MSE", "achieved": warp_bytes / (warp_ms * 1e-3) / 1e9,
                     "peak": peaks["hbm"], "unit": "GB/s", "frac": warp_bytes / (warp_ms * 1e-3) / 1e9 / peaks["hbm"],
                     "traffic": traffic.get("warp_tile_kernel") or traffic.get("warp_loss_kernel"),
                     "bytes_per_launch": warp_bytes, "kernel_ms_per_step": warp_ms}
    e2e = measure_e2e(gmodels["affine"] if use_graph else models["affine"], "affine", img_f_host, img_m_host, args.steps,
                      ctx, with_grid=args.e2e_grid, clone_outputs=use_graph)
    clk = clocks.stop()
    dtype_name = "fp16" if kb.act_dtype() == torch.float16 else "bf16"
    engine_points = {"affine": {dtype_name: res["points"]}}
    # the same workload with the other operand type (BASELINE configs[1] names bf16)
    other = "bf16" if dtype_name == "fp16" else "fp16"
    kb.set_operand_dtype(other)
    try:
        rb = measure_pairwise(models["affine"], "affine", img_f, img_m, args.steps, args.warmup, ctx)
        rbg = measure_graph(gmodels["affine"], "affine", img_f, img_m, args.steps, args.warmup, ctx) if use_graph else None
        ob_ms = rbg["ms_step"] if rbg and rbg["captured"] else rb["ms_step"]
        other_block = {"value": world * 1e3 / ob_ms, "unit": UNIT, "ms_per_step": ob_ms, "mse": rb["mse"],
                       "eager_ms_per_step": rb["ms_step"],
                       "note": f"identical kernels and schedule with {other} activations / weights"}
        engine_points["affine"][other] = rb["points"]
    finally:
        kb.set_operand_dtype(dtype_name)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype_name, "data": "synthetic", "config": config_dict(world),
            "clocks": clk, "e2e": e2e, "gpu_launches": res["launches"], "roofline": roofline,
            "roofline_warp": roofline_warp, "roofline_other": roofline_other, "mse": res["mse"],
            f"{other}_operands": other_block,
            "cpu_affinity": ({"cpus_rank0": len(numa_cpus), "source": "NVML cpu affinity of the rank's GPU"}
                             if numa_cpus else None),
            "execution": {"mode": "cuda_graph_replay" if use_graph else "eager",
                          "api": "keymorph_b200.KeyMorph(..., fused_warp=True, cuda_graph=%s).forward" % use_graph,
                          "eager_value": world * 1e3 / eager_ms, "eager_ms_per_step": eager_ms,
                          "graph_state": gres["state"] if gres else "not requested (--eager)",
                          "graph_mse": gres["mse"] if gres else None,
                          "note": "value / e2e: forward() captured once per (shape, transform list) into one CUDA graph and "
                                  "replayed (2 staging copies + 1 graph launch per registration, identical kernels); the "
                                  "rooflines and gpu_launches come from the eager pass of the same model, timed with "
                                  "CUDA events around every C-ABI call"}}

    # ---- configs[2], the north-star target: TPS lambda = 0, 512 keypoints
    if "tps" in models:
        c2 = ClockSampler(local)
        c2.start()
        r3 = measure_pairwise(models["tps"], "tps", img_f, img_m, args.steps, args.warmup, ctx)
        g3 = measure_graph(gmodels["tps"], "tps", img_f, img_m, args.steps, args.warmup, ctx) if use_graph else None
        graph3 = g3 is not None and g3["captured"]
        ms3 = g3["ms_step"] if graph3 else r3["ms_step"]
        engine_points["tps"] = {dtype_name: r3["points"]}
        tr = r3["tr"]
        flow_ms, fit_ms = tr.ms("km_flow_field_tps", per=args.steps), tr.ms("km_tps_fit", per=args.steps)
        wl_ms = tr.ms("km_warp_loss", per=args.steps)
        terms = 512.0 * S ** 3
        clk3 = None
        e2e3 = measure_e2e(gmodels["tps"] if graph3 else models["tps"], "tps", img_f_host, img_m_host, args.steps, ctx,
                           with_grid=args.e2e_grid, clone_outputs=graph3)
        clk3 = c2.stop()
        f_sm = (clk3.get("sm_mhz") or peaks["sm_max_mhz"]) * 1e6
        # one MUFU (lg2) per radial-basis term, 16 MUFU lanes / clk / SM (XU pipe), 148 SMs
        xu_peak = 16 * 148 * f_sm / 1e9
        main3, other3 = conv_rooflines(tr, "tps", args.steps, r3["ms_step"], peaks, r3["ms_step"] * args.steps * 1e-3)
        line["tps_config3"] = {
            "workload": CONFIGS["tps"]["workload"], "value": world * 1e3 / ms3, "unit": UNIT,
            "ms_per_step": ms3, "execution": "cuda_graph_replay" if graph3 else "eager",
            "eager_value": world * 1e3 / r3["ms_step"], "eager_ms_per_step": r3["ms_step"], "e2e": e2e3, "gpu_launches": r3["launches"], "mse": r3["mse"], "clocks": clk3,
            "roofline": {"bound": "issue/XU (MUFU)", "kernel": "flow_tps_rows_kernel (dense TPS field, K radial-basis terms per voxel)",
                         "achieved": terms / (flow_ms * 1e-3) / 1e9, "peak": xu_peak, "unit": "G terms/s",
                         "frac": terms / (flow_ms * 1e-3) / 1e9 / xu_peak, "kernel_ms_per_step": flow_ms,
                         "terms_per_launch": terms, "traffic": traffic.get("flow_tps_rows_kernel"),
                         "peak_source": f"1 MUFU.LG2 per term x 16 lanes/clk/SM x 148 SMs x {f_sm / 1e6:.0f} MHz (median SM clock sampled "
                                        "during the run); the kernel writes 12 B/voxel and is nowhere near HBM"},
            "roofline_warp": {"bound": "hbm", "kernel": "fused warp (flow field read) + MSE", "achieved": warp_bytes / (wl_ms * 1e-3) / 1e9,
                              "peak": peaks["hbm"], "unit": "GB/s", "frac": warp_bytes / (wl_ms * 1e-3) / 1e9 / peaks["hbm"],
                              "bytes_per_launch": warp_bytes, "kernel_ms_per_step": wl_ms},
            "roofline_conv": main3, "tps_fit_ms_per_step": fit_ms,
            "backbone_ms_per_step": other3[-1]["kernel_ms_per_step"]}

    # ---- configs[4]: groupwise, the only collective on the path
    if "tps" in models and not args.no_groupwise:
        try:
            line["groupwise"] = groupwise_block(models["tps"], ctx)
        except Exception as e:  # noqa: BLE001
            line["groupwise"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        del models, gmodels
        torch.cuda.empty_cache()
        line["gpu_baseline"] = gpu_baseline(dev, img_f, img_m, engine_points=engine_points)
        gb = line["gpu_baseline"]
        for name, mine in (("affine", value), ("tps", line.get("tps_config3", {}).get("value"))):
            best = max([gb[k]["value"] for k in gb if k.startswith(name) and isinstance(gb[k], dict) and "value" in gb[k]],
                       default=None)
            if best and mine:
                gb[f"{name}_speedup_over_best_torch_gpu"] = mine / best
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            step, kind = reference_step_fn("affine")
            torch.set_num_threads(os.cpu_count() or 1)
            f_cpu, m_cpu = img_f_host.clone(), img_m_host.clone()
            t0 = time.perf_counter()
            cm = float(step(f_cpu, m_cpu))
            sec = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": "ONE complete registration of the same 256^3 pair (no warm-up, no slab), "
                                              f"torch CPU fp32, {torch.get_num_threads()} threads", "mse": cm}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's ORIGINAL stdout."""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    # libraries chat on stdout (NCCL prints its version banner there at the first communicator): keep fd 1 for
    # the JSON line only and send everything else to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-GPU reference leg")
    ap.add_argument("--no-tps", action="store_true", help="skip the configs[2] (TPS) block")
    ap.add_argument("--no-groupwise", action="store_true", help="skip the configs[4] (groupwise) block")
    ap.add_argument("--no-numa", action="store_true", help="do not bind each rank to its GPU's NUMA node (N > 1)")
    ap.add_argument("--eager", action="store_true", help="headline from the eager path (no CUDA-graph replay)")
    ap.add_argument("--e2e-grid", action="store_true", help="the e2e loop also copies the flow field to the host")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
