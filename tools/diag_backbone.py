"""Accuracy diagnostics of the bf16 backbone on the GPU box: keypoint drift of the CUDA engine vs a
true-fp32 evaluation of the same network, next to the drift of the reference's own reduced-precision
paths (TF32 convolutions = torch default on CUDA, fp16 / bf16 autocast = `use_amp`), per volume size,
plus per-layer relative errors.  Usage: python tools/diag_backbone.py [S ...]"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import engine as E  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

DEV = "cuda"
K = 64


def layerwise_oracle(sd, x):
    """unet3d forward (truncated 1, 4 levels) capturing every SingleConv output."""
    outs = {}

    def sc(prefix, t):
        c = t.shape[1]
        g = 8 if c >= 8 else 1
        t = F.group_norm(t, g, sd[prefix + ".groupnorm.weight"], sd[prefix + ".groupnorm.bias"], 1e-5)
        return F.relu(F.conv3d(t, sd[prefix + ".conv.weight"], None, padding=1))

    feats = []
    for i in range(4):
        if i > 0:
            x = F.max_pool3d(x, 2)
        x = sc(f"encoders.{i}.basic_module.SingleConv1", x)
        outs[f"enc{i}.c1"] = x
        x = sc(f"encoders.{i}.basic_module.SingleConv2", x)
        outs[f"enc{i}.c2"] = x
        feats.insert(0, x)
    feats = feats[1:]
    for j in range(2):
        up = F.interpolate(x, size=feats[j].shape[2:], mode="nearest")
        x = torch.cat((feats[j], up), 1)
        x = sc(f"decoders.{j}.basic_module.SingleConv1", x)
        outs[f"dec{j}.c1"] = x
        x = sc(f"decoders.{j}.basic_module.SingleConv2", x)
        outs[f"dec{j}.c2"] = x
    outs["heat"] = F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])
    return outs


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
    torch.manual_seed(23)
    net = kb.TruncatedUNet3D(1, K, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8,
                             num_levels=4, is_segmentation=False, conv_padding=1).eval().to(DEV)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    model = kb.KeyMorph(net, K, 3).eval()
    eng = E.backbone_engine(net)
    for S in sizes:
        img = O.gaussian_phantom(S, 1000).to(DEV)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            ref_layers = layerwise_oracle(sd, img)
            ref = O.center_of_mass3d(ref_layers["heat"])
            torch.backends.cudnn.allow_tf32 = True
            tf32 = O.center_of_mass3d(O.unet3d_forward(sd, img, 4, 1))
            torch.backends.cudnn.allow_tf32 = False
            with torch.autocast("cuda", dtype=torch.float16):
                fp16 = O.center_of_mass3d(O.unet3d_forward(sd, img, 4, 1).float())
            with torch.autocast("cuda", dtype=torch.bfloat16):
                bf16 = O.center_of_mass3d(O.unet3d_forward(sd, img, 4, 1).float())
        eng.debug = {}
        ours, feat = model.get_keypoints(img, return_feat=True)
        dbg, eng.debug = eng.debug, None

        def e(p):
            d = (p.float() - ref).abs()
            return f"max {d.max().item():.2e} mean {d.mean().item():.2e}"
        print(f"S={S} K={K}: engine(bf16) {e(ours)} | torch tf32 {e(tf32)} | autocast fp16 {e(fp16)} | "
              f"autocast bf16 {e(bf16)}", flush=True)
        for name, t in dbg.items():
            r = ref_layers[name]
            got = kb.ops.ndhwc_to_ncdhw(t)[:, : r.shape[1]]
            rel = ((got - r).abs().mean() / r.abs().mean()).item()
            mx = ((got - r).abs().max() / r.abs().max()).item()
            print(f"   {name:8s} mean rel err {rel:.2e}  max err / max {mx:.2e}  (ref mean {r.mean().item():.3e})")
        r = ref_layers["heat"]
        rel = ((feat - r).abs().mean() / r.abs().mean()).item()
        print(f"   heat     mean rel err {rel:.2e}")
        del ref_layers, dbg


if __name__ == "__main__":
    main()
