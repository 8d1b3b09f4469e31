"""Benchmark of the pairwise registration hot path (BASELINE.json: "pairwise 256^3
registrations/sec"; workload = configs[1]: synthetic 256^3 pair, affine, 256 keypoints, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pairwise registration: two 256^3 fp32 volumes -> TruncatedUNet3D backbone on both
-> centre-of-mass keypoints -> affine fit -> flow field -> warped moving image + MSE.
  value : registrations/s with the volumes already resident in HBM (CUDA events, max over ranks)
  e2e   : the same call with HOST (pinned) volumes: H2D of both volumes and D2H of the MSE of every
          step inside the timed region (the MSE of step i is read after step i+1 was enqueued)
  roofline     : the tcgen05 convolution kernel (tensor bound), timed live with CUDA events
  cpu_baseline : the oracle (CPU port of the reference path) on the host cores, bounded sample
--impl reference times that CPU port as its own arm (rank 0 only).
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

S, K = 256, 256
WORKLOAD = "synthetic 256^3 pair, affine, 256 keypoints, TruncatedUNet3D(levels 4, truncated 1), bf16 operands"
METRIC = "pairwise 256^3 registrations/sec"
UNIT = "registrations/s"
CPU_SLAB = 64      # the CPU arm times a 64x256x256 slab of each volume and scales by 256/64
# dram__bytes_read.sum + dram__bytes_write.sum per step from the ncu capture of this workload
# (profiles/r01_s3_ncu_launches_dram_traffic_default.txt); filled in by hand after each profiling pass
NCU_TRAFFIC = {"conv_tc_kernel": 5.639e9,    # 10 launches: 3.617 GB read + 2.023 GB written
               "conv_zf_kernel": 1.351e9,    # 1.100 GB read + 0.251 GB written (pooled output only)
               "com_tc_kernel": 0.540e9}     # 0.537 GB read


def conv_layers(S, K, n_img):
    """(name, Cin, Cout, edge, taps) of every conv on the path, keymorph/unet3d/buildingblocks.py
    :171-181 channel rule; FLOPs = 2*taps*Cin*Cout*edge^3 per image."""
    L = [("enc0.c1", 1, 16, S, 27), ("enc0.c2", 16, 32, S, 27), ("enc1.c1", 32, 32, S // 2, 27),
         ("enc1.c2", 32, 64, S // 2, 27), ("enc2.c1", 64, 64, S // 4, 27), ("enc2.c2", 64, 128, S // 4, 27),
         ("enc3.c1", 128, 128, S // 8, 27), ("enc3.c2", 128, 256, S // 8, 27),
         ("dec0.c1", 384, 128, S // 4, 27), ("dec0.c2", 128, 128, S // 4, 27),
         ("dec1.c1", 192, 64, S // 2, 27), ("dec1.c2", 64, 64, S // 2, 27), ("final", 64, K, S // 2, 1)]
    return [(n, ci, co, e, t, 2.0 * t * ci * co * e ** 3 * n_img) for (n, ci, co, e, t) in L]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
def cpu_sample_seconds(sd, steps=1, warmup=0):
    """Reference CPU path (oracle port, torch-CPU fp32, all host threads) on a 64-slab of the 256^3
    pair; returns seconds per FULL registration (slab time x 256/64)."""
    import torch
    from oracle import keymorph_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    f = O.gaussian_phantom(S, 1000)[:, :, :CPU_SLAB].contiguous()
    m = O.gaussian_phantom(S, 2000)[:, :, :CPU_SLAB].contiguous()
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            r = O.keymorph_forward("truncatedunet", sd, f, m, "affine")["affine"]
            img_a = O.align_img(r["grid"], m)
            mse = O.mse_loss(img_a, f).item()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert mse == mse
    return statistics.mean(times) * (S / CPU_SLAB), times


def seeded_state_dict():
    import torch
    import keymorph_b200 as kb
    torch.manual_seed(23)
    net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                             num_levels=4, is_segmentation=False, conv_padding=1).eval()
    return net


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    net = seeded_state_dict()
    sd = {k: v for k, v in net.state_dict().items()}
    sec, times = cpu_sample_seconds(sd, steps=args.steps, warmup=args.warmup)
    value = 1.0 / sec
    cores = os.cpu_count() or 1
    sample = (f"{CPU_SLAB}x{S}x{S} slab of each 256^3 volume through the full pipeline (2 backbone passes, CoM, "
              f"affine fit, flow field, warp, MSE); per-step time scaled x{S // CPU_SLAB}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist

    import keymorph_b200 as kb
    from keymorph_b200 import _lib, ops
    from oracle import keymorph_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    net = seeded_state_dict()
    sd_cpu = {k: v.clone() for k, v in net.state_dict().items()}
    model = kb.KeyMorph(torch.nn.DataParallel(net.to(dev)), K, 3, fused_warp=True).eval()

    # synthetic pair: Gaussian-blob phantoms, moving = affine-warped second phantom
    img_f_host = O.gaussian_phantom(S, 1000 + rank).pin_memory()
    base_m = O.gaussian_phantom(S, 2000 + rank).to(dev)
    Minv = torch.inverse(O.affine_matrix_3d(0.1, 0.05, 0.3, 0.02)).to(dev)
    img_m_host = ops.warp_loss(base_m, None, mat34=Minv[:, :3])[0].cpu().pin_memory()
    del base_m
    img_f = img_f_host.to(dev)
    img_m = img_m_host.to(dev)

    def step(f, m):
        r = model(f, m, transform_type="affine", return_aligned_points=True)["affine"]
        return r["mse"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(img_f, img_m)
    sync_all()

    # tracer: CUDA events around the kernels whose rooflines are reported
    traced = {"km_conv3d_tc": [], "km_conv3d_tc_pair": [], "km_conv3d_zfold_pair": [], "km_conv3d_zfold": [],
              "km_conv3d_zfold_gn": [], "km_conv3d_zfold_pair_gn": [], "km_conv3d_zfold_pair_gn_cat": [], "km_conv3d_tc_pair_gn": [], "km_conv1x1_com": [],
              "km_conv3d_stem": [], "km_warp_loss": []}
    stream = torch.cuda.current_stream()
    pending = {}

    def trace(name, phase):
        if name in traced:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            if phase == "pre":
                pending[name] = ev
            else:
                traced[name].append((pending.pop(name), ev))

    clocks = ClockSampler(local)
    clocks.start()
    launches0 = _lib.launch_count
    _lib.TRACE = trace
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mse = step(img_f, img_m)
    e1.record()
    sync_all()
    _lib.TRACE = None
    launches = _lib.launch_count - launches0
    ms_step = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step_max = t.item()
    value = world * 1e3 / ms_step_max

    # roofline of the dominant kernel (conv_tc_kernel, the general tcgen05 3x3x3 convolution): algorithmic
    # FLOPs of the layers it runs in one step / summed duration of its launches in one step.  The
    # other tensor-core kernels of the backbone are reported next to it with the same definition.
    hbm_peak, tc_peak, peak_src = measured_peaks()

    def per_step_ms(name):
        return sum(a.elapsed_time(b) for a, b in traced[name]) / args.steps

    # conv_tc_kernel (one SM per MMA) and conv_tc2_kernel (cta_group::2) are the same implicit GEMM
    conv_names = ("km_conv3d_tc", "km_conv3d_tc_pair", "km_conv3d_zfold_pair", "km_conv3d_zfold_pair_gn",
                  "km_conv3d_zfold_pair_gn_cat", "km_conv3d_tc_pair_gn")
    conv_ms = sum(per_step_ms(k) for k in conv_names)
    zf_ms = per_step_ms("km_conv3d_zfold") + per_step_ms("km_conv3d_zfold_gn")
    com_ms = per_step_ms("km_conv1x1_com")
    stem_ms, warp_ms = per_step_ms("km_conv3d_stem"), per_step_ms("km_warp_loss")
    layers = {l[0]: l[5] for l in conv_layers(S, K, 2)}
    zf_flops = layers["enc0.c2"] if zf_ms > 0 else 0.0
    com_flops = layers["final"] if com_ms > 0 else 0.0
    tc_flops = sum(v for k, v in layers.items() if k != "enc0.c1") - zf_flops - com_flops
    achieved = tc_flops / (conv_ms * 1e-3) / 1e12
    n_conv = sum(len(traced[k]) for k in conv_names) // args.steps
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc2_kernel / conv_zf2_kernel (cta_group::2)", "achieved": achieved, "peak": tc_peak,
                "unit": "TFLOP/s", "frac": achieved / tc_peak, "traffic": NCU_TRAFFIC.get("conv_tc_kernel"),
                "flops_per_step": tc_flops, "launches_per_step": n_conv, "kernel_ms_per_step": conv_ms,
                "share_of_step": conv_ms / ms_step, "peak_source": peak_src,
                "traffic_note": "DRAM bytes (read+write) summed over the kernel's launches of one step, ncu "
                                "capture in profiles/ (see DESIGN.md section 6)"}
    backbone_ms = conv_ms + zf_ms + com_ms + stem_ms
    all_flops = sum(layers.values())
    roofline_other = []
    if zf_ms > 0:
        roofline_other.append({"bound": "tensor", "kernel": "conv_zf_kernel (16->32 @256^3, dz folded into N, pool fused%s)" % (
                                   ", GroupNorm folded in" if traced["km_conv3d_zfold_gn"] else ""),
                               "achieved": zf_flops / (zf_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                               "frac": zf_flops / (zf_ms * 1e-3) / 1e12 / tc_peak, "kernel_ms_per_step": zf_ms,
                               "traffic": NCU_TRAFFIC.get("conv_zf_kernel")})
    if com_ms > 0:
        roofline_other.append({"bound": "tensor", "kernel": "com_tc_kernel (final 1x1x1 conv + ReLU + centre of mass)",
                               "achieved": com_flops / (com_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                               "frac": com_flops / (com_ms * 1e-3) / 1e12 / tc_peak, "kernel_ms_per_step": com_ms,
                               "hbm_gbs": 2 * (S // 2) ** 3 * 64 * 2 / (com_ms * 1e-3) / 1e9,
                               "traffic": NCU_TRAFFIC.get("com_tc_kernel")})
    roofline_other.append({"bound": "tensor", "kernel": "whole backbone (stem%s + conv_zf + conv_tc + com_tc)" % (
                               "" if traced["km_conv3d_zfold_gn"] else " x2"),
                           "achieved": all_flops / (backbone_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                           "frac": all_flops / (backbone_ms * 1e-3) / 1e12 / tc_peak,
                           "kernel_ms_per_step": backbone_ms})
    # the HBM-bound kernel of the path: ONE pass writes the affine flow field (12 B/voxel), gathers the
    # moving volume (4), reads the fixed volume (4), stores the warped volume (4) and reduces the
    # MSE sums (fused warp + loss, grid generated in registers)
    warp_bytes = 24.0 * S ** 3
    warp_gbs = warp_bytes / (warp_ms * 1e-3) / 1e9 if warp_ms > 0 else None
    roofline_warp = {"bound": "hbm", "kernel": "warp_loss_kernel<AFFINE> (+grid store, +MSE)", "achieved": warp_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": (warp_gbs / hbm_peak) if warp_gbs else None,
                     "traffic": None, "bytes_per_launch": warp_bytes, "kernel_ms_per_step": warp_ms}

    # end to end through the public API with HOST (pinned) buffers: every step copies both volumes
    # host->device (prefetched on a side stream by keymorph_b200.hostio) and reads the MSE back
    from keymorph_b200.hostio import prefetch_to_device
    sync_all()
    for f, m in prefetch_to_device([(img_f_host, img_m_host)] * 2, dev):
        step(f, m).item()
    sync_all()
    # every step's MSE is copied to pinned host memory and read; the read of step i happens after
    # step i+1 has been enqueued, so the host never idles the GPU while it waits for a scalar
    host_mse = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    t0 = time.perf_counter()
    loss, pending = None, None
    for i, (f, m) in enumerate(prefetch_to_device([(img_f_host, img_m_host)] * args.steps, dev)):
        host_mse[i & 1].copy_(step(f, m), non_blocking=True)     # D2H read of the step's result
        evs[i & 1].record()
        if pending is not None:
            evs[pending].synchronize()
            loss = float(host_mse[pending])
        pending = i & 1
    evs[pending].synchronize()
    loss = float(host_mse[pending])
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clk = clocks.stop()
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": world * 1e3 / t.item(), "unit": UNIT,
           "h2d_bytes_per_step": 2 * img_f_host.numel() * 4, "d2h_bytes_per_step": 4}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": 1, "parallelism": f"pairs sharded x{world}",
                       "l2": "per-step working set (>5 GB of activations) exceeds the 126 MB L2"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "roofline_warp": roofline_warp, "roofline_other": roofline_other, "mse": float(mse.item()), "loss_e2e": loss}

    if rank == 0 and world == 1 and not args.no_cpu:
        sec, _ = cpu_sample_seconds(sd_cpu, steps=1, warmup=0)
        line["cpu_baseline"] = {
            "value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"one {CPU_SLAB}x{S}x{S} slab of the same pair through the oracle pipeline, time scaled "
                      f"x{S // CPU_SLAB}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
