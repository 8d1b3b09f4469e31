"""CPU-side checks: the C-ABI library loads and exports every symbol include/km_b200.h declares,
the ctypes table matches the header, the host logic (state-dict layout, transform parsing, sharding,
groupwise exchange over gloo with world_size 2) behaves like the reference.  No GPU needed."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import keymorph_b200 as kb
from keymorph_b200 import _lib, parallel
from oracle import keymorph_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "km_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t|const char\*)\s+(km_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(2)] = n
    return out


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in km_b200.h but not exported"
    assert lib.km_version() >= 100


def test_ctypes_table_matches_header():
    declared = _header_functions()
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    for name, n in declared.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_library_is_plain_c_abi():
    """no torch / libstdc++ types leak through the boundary: only libc-level dependencies."""
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libc10" not in out


def test_no_cpu_fallback():
    x = torch.zeros(1, 1, 8, 8, 8)
    g = torch.zeros(1, 8, 8, 8, 3)
    with pytest.raises(_lib.KMError):
        kb.align_img(g, x)
    with pytest.raises(_lib.KMError):
        kb.CenterOfMass3d("ij")(x)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "keymorph_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert "keymorph_oracle" not in src, fn


def test_state_dict_layout_and_dataparallel_prefix():
    """scripts/script_utils.py:59-81: checkpoints carry DataParallel's 'module.' prefix."""
    net = kb.TruncatedUNet3D(1, 16, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                             num_levels=4, is_segmentation=False, conv_padding=1)
    keys = set(net.state_dict())
    assert "encoders.0.basic_module.SingleConv1.groupnorm.weight" in keys
    assert "decoders.1.basic_module.SingleConv2.conv.weight" in keys
    assert "final_conv.bias" in keys
    assert net.encoders[0].basic_module.SingleConv1.conv.weight.shape == (16, 1, 3, 3, 3)
    assert net.decoders[0].basic_module.SingleConv1.conv.weight.shape == (128, 384, 3, 3, 3)
    assert net.final_conv.weight.shape == (16, 64, 1, 1, 1)
    assert len(net.decoders) == 2
    model = kb.KeyMorph(torch.nn.DataParallel(net), 16, 3)
    sd = {"module." + k: torch.randn_like(v) for k, v in net.state_dict().items()}
    model.backbone.load_state_dict(sd, strict=True)
    assert torch.equal(net.final_conv.bias, sd["module.final_conv.bias"])
    cn = kb.ConvNet(3, 1, 16, norm_type="instance")
    assert set(cn.state_dict()) == {f"block{b}.conv.{p}" for b in range(1, 10) for p in ("weight", "bias")}


def test_transform_type_parsing_matches_reference():
    m = kb.KeyMorph(torch.nn.Identity(), 4, 3)
    for s, ok in (("rigid", True), ("affine", True), ("tps_0", True), ("tps_loguniform", True),
                  ("bspline", False), ("tps", False)):
        assert m.is_supported_transform_type(s) == ok
    assert torch.equal(m._convert_tps_lmbda(2, 0.5), torch.tensor([0.5, 0.5]))
    lam = m._convert_tps_lmbda(64, "uniform")
    assert lam.shape == (64,) and lam.min() >= 0 and lam.max() <= 10
    lam = m._convert_tps_lmbda(8, "loguniform")
    assert lam.min() >= 1e-6 and lam.max() <= 10
    with pytest.raises(KeyError):
        m(torch.zeros(1, 1, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8))   # return_aligned_points is required


def test_shard_range_partitions():
    for n, w in ((64, 8), (10, 3), (3, 8), (0, 4)):
        seen = []
        for r in range(w):
            seen += list(parallel.shard_range(n, r, w))
        assert seen == list(range(n))
        sizes = [len(parallel.shard_range(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from keymorph_b200 import parallel
from oracle import keymorph_oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
g = torch.Generator().manual_seed(5)
pts = torch.rand(6, 12, 3, generator=g) * 1.2 - 0.6        # all subjects, same on both ranks
mine = pts[list(parallel.shard_range(6, rank, 2))]
def reg(p, mean):
    out = torch.zeros_like(p)
    for i in range(len(p)):
        tm, _ = O.aligner_matrices(p[i:i+1], mean, None, "affine")
        out[i:i+1] = O.transform_points(tm, p[i:i+1])
    return out
cur, mean = parallel.groupwise_iterate(mine, reg, 3)
ref_cur, ref_mean = O.groupwise_points(pts, "affine", 3)
ref_mine = ref_cur[list(parallel.shard_range(6, rank, 2))]
assert torch.allclose(cur, ref_mine, atol=1e-5), (cur - ref_mine).abs().max()
assert torch.allclose(mean, ref_mean, atol=1e-6)
gm = parallel.global_mean_points(mine)
assert torch.allclose(gm, pts.mean(0, keepdim=True), atol=1e-6)
# the one-collective scheme: all-gather every subject's keypoints once, iterate locally on all of them;
# the mean is torch.mean over the same (G,K,3) tensor the single process sees -> identical bits.
# 7 subjects -> shards of 4 and 3 (padded all-gather)
pts7 = (torch.rand(1, 12, 3, generator=g) * 1.2 - 0.6) + 0.05 * torch.randn(7, 12, 3, generator=g)   # one anatomy, 7 subjects
idx7 = list(parallel.shard_range(7, rank, 2))
allp, off = parallel.gather_all_points(pts7[idx7])
assert torch.equal(allp, pts7) and off == idx7[0]
cur2, mean2 = parallel.groupwise_iterate(pts7[idx7], reg, 3, mode="allgather")
ref_cur7, ref_mean7 = O.groupwise_points(pts7, "affine", 3)
assert torch.equal(mean2, ref_mean7), (mean2 - ref_mean7).abs().max()
assert torch.allclose(cur2, ref_cur7[idx7], atol=1e-6)
cur3, mean3 = parallel.groupwise_iterate(pts7[idx7], reg, 3, mode="allreduce")
assert torch.allclose(mean3, mean2, atol=2e-6), (mean3 - mean2).abs().max()    # summation order differs
assert torch.allclose(cur3, cur2, atol=1e-5), (cur3 - cur2).abs().max()
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_groupwise_exchange_gloo_world2(tmp_path):
    """The N>1 path on CPU: two gloo ranks, subjects partitioned, all-reduced mean keypoints; the
    result equals the single-process oracle iteration (keymorph/model.py:331-444)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


# ---------------------------------------------------------------- NIfTI reader edge cases
def _write_nifti(path, data, sform=None, qform=None, slope=float("nan"), inter=float("nan"), pixdim=(1, 1, 1, 1)):
    import struct
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, *([data.ndim] + list(data.shape) + [1] * (7 - data.ndim)))
    struct.pack_into("<h", hdr, 70, {"float32": 16, "int16": 4, "int32": 8}[data.dtype.name])
    struct.pack_into("<8f", hdr, 76, *pixdim, 0, 0, 0, 0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, slope, inter)
    if qform is not None:
        struct.pack_into("<h", hdr, 252, 1)
        struct.pack_into("<6f", hdr, 256, *qform)
    if sform is not None:
        struct.pack_into("<h", hdr, 254, 1)
        struct.pack_into("<12f", hdr, 280, *np.asarray(sform, dtype=np.float32).reshape(-1))
    hdr[344:348] = b"n+1\0"
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(data.tobytes(order="F"))


def test_nifti_reader_edge_cases(tmp_path):
    from keymorph_b200 import hostio
    x = np.arange(4 * 6 * 8, dtype=np.float32).reshape(4, 6, 8)
    # NaN scl_slope / scl_inter mean "no scaling" (nibabel); qform-only headers are honoured
    _write_nifti(tmp_path / "a.nii", x, qform=(0, 0, 0, -10, -20, -30), pixdim=(1, 2, 3, 4))
    data, aff = hostio.read_nifti(str(tmp_path / "a.nii"))
    assert np.array_equal(data, x)
    assert np.allclose(aff, [[2, 0, 0, -10], [0, 3, 0, -20], [0, 0, 4, -30], [0, 0, 0, 1]])
    # 90 degrees about z with qfac = -1
    _write_nifti(tmp_path / "b.nii", x, qform=(0, 0, np.sin(np.pi / 4), 0, 0, 0), pixdim=(-1, 1, 1, 1))
    _, aff = hostio.read_nifti(str(tmp_path / "b.nii"))
    assert np.allclose(aff[:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, -1]], atol=1e-6)
    # a real scaling is applied
    _write_nifti(tmp_path / "s.nii", x, sform=np.eye(4)[:3], slope=2.0, inter=1.0)
    assert np.allclose(hostio.read_nifti(str(tmp_path / "s.nii"))[0], 2 * x + 1)
    # block-mean downsample: voxel size doubles AND the origin moves to the centre of the first block
    y = np.random.RandomState(0).rand(8, 8, 8).astype(np.float32)
    _write_nifti(tmp_path / "c.nii", y, sform=[[2, 0, 0, -5], [0, 2, 0, -6], [0, 0, 2, -7]])
    t, aff = hostio.load_volume(str(tmp_path / "c.nii"), size=4)
    assert t.shape == (1, 1, 4, 4, 4) and np.allclose(aff[:3, 3], [-4, -5, -6]) and np.allclose(np.diag(aff)[:3], 4)
    # the strided label pick keeps voxel 0 where it was; ids that do not fit uint8 are refused, not wrapped
    lab = (np.arange(512).reshape(8, 8, 8) % 7).astype(np.int32)
    _write_nifti(tmp_path / "l.nii", lab, sform=[[2, 0, 0, -5], [0, 2, 0, -6], [0, 0, 2, -7]])
    t, aff = hostio.load_volume(str(tmp_path / "l.nii"), size=4, labels=True)
    assert t.dtype == torch.uint8 and np.allclose(aff[:3, 3], [-5, -6, -7])
    _write_nifti(tmp_path / "m.nii", (lab * 400).astype(np.int32), sform=np.eye(4)[:3])
    with pytest.raises(ValueError):
        hostio.load_volume(str(tmp_path / "m.nii"), size=4, labels=True)


def test_deferred_singular_checks_are_per_thread():
    """transformations.deferred_singular_checks keeps its pending list in thread-local storage."""
    import threading
    from keymorph_b200 import transformations as T
    seen = {}

    def worker():
        seen["inner"] = getattr(T._TLS, "pending", None)

    with T.deferred_singular_checks():
        assert T._TLS.pending == []
        th = threading.Thread(target=worker)
        th.start()
        th.join()
    assert seen["inner"] is None and getattr(T._TLS, "pending", None) is None


def test_bind_cpu_to_gpu_is_best_effort():
    """parallel.bind_cpu_to_gpu narrows the CPU affinity to the GPU's NUMA node when NVML knows it and is a no-op
    (None, no exception, affinity untouched) otherwise -- e.g. in this container, which has no GPU."""
    import os

    from keymorph_b200.parallel import bind_cpu_to_gpu
    before = os.sched_getaffinity(0)
    got = bind_cpu_to_gpu(0)
    after = os.sched_getaffinity(0)
    if got is None:
        assert after == before
    else:
        assert set(got) == after and after < before
        os.sched_setaffinity(0, before)
    assert bind_cpu_to_gpu(10 ** 6) is None       # no such device: still no exception
    assert os.sched_getaffinity(0) == before


def test_bench_flop_model_and_traffic_profile():
    """bench.py's roofline inputs: the executed-flop model counts the decoders' upsampled channels with 8 taps
    instead of 27 (conv_up2.cu) and stays below the reference's 27-tap arithmetic by exactly that difference; the
    committed ncu launch list parses into per-kernel DRAM bytes for every tcgen05 conv kernel it names."""
    import bench
    ex = {l[0]: l for l in bench.conv_layers(256, 256, 2)}
    al = {l[0]: l for l in bench.conv_layers(256, 256, 2, executed=False)}
    assert set(al) == {"enc0.c1", "enc0.c2", "enc1.c1", "enc1.c2", "enc2.c1", "enc2.c2", "enc3.c1", "enc3.c2",
                       "dec0.c1", "dec0.c2", "dec1.c1", "dec1.c2", "final"}
    assert al["dec1.c1"][1:5] == (192, 64, 128, 27) and al["dec0.c1"][1:5] == (384, 128, 64, 27)
    assert ex["dec1.c1"][1:5] == (64, 64, 128, 27) and ex["dec1.c1.up"][1:5] == (128, 64, 128, 8)
    assert ex["dec0.c1"][1:5] == (128, 128, 64, 27) and ex["dec0.c1.up"][1:5] == (256, 128, 64, 8)
    f_ex, f_al = sum(l[5] for l in ex.values()), sum(l[5] for l in al.values())
    saved = 2.0 * 19 * (128 * 64 * 128 ** 3 + 256 * 128 * 64 ** 3) * 2       # 19 of 27 taps of the upsampled channels
    assert abs((f_al - f_ex) - saved) < 1e-6 * f_al
    assert abs(al["enc1.c1"][5] - 2.0 * 27 * 32 * 32 * 128 ** 3 * 2) < 1.0
    tr = bench.traffic_from_profile()
    assert tr is not None
    for k in ("conv_zf_kernel", "conv_zf2_kernel", "conv_up2_kernel", "conv_tc2_kernel", "conv_stem_mma_kernel"):
        assert tr.get(k, 0.0) > 1e8, k          # hundreds of MB per step each
    assert 0.5e9 < tr["conv_up2_kernel"] < 1.2e9     # reads the coarse tensors, writes the 16-bit partial sums
