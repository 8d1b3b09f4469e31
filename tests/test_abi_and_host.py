"""CPU-side checks: the C-ABI library loads and exports every symbol include/km_b200.h declares,
the ctypes table matches the header, the host logic (state-dict layout, transform parsing, sharding,
groupwise exchange over gloo with world_size 2) behaves like the reference.  No GPU needed."""
import ctypes
import os
import re
import subprocess
import sys

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
