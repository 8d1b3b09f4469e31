"""Layer-level error of the stem -> conv_zf pair on the device, three ways, against the fp64 CPU layer:
two-pass stem + normalised store (no fold), one-pass stem + GroupNorm folded into conv_zf (raw), the same
with the centred store.  python tools/fold_error_probe.py [size]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keymorph_b200 as kb  # noqa: E402
from keymorph_b200 import ops  # noqa: E402
from oracle import keymorph_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
torch.manual_seed(23)
net = kb.TruncatedUNet3D(1, 64, 1, final_sigmoid=False, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                         is_segmentation=False, conv_padding=1).eval()
sd = {k: v.clone().double() for k, v in net.state_dict().items()}
p = "encoders.0.basic_module."
x = torch.cat([O.gaussian_phantom(S, 1000), O.gaussian_phantom(S, 7)]).double()
x0 = F.group_norm(x, 1, sd[p + "SingleConv1.groupnorm.weight"], sd[p + "SingleConv1.groupnorm.bias"], 1e-5)
v = F.relu(F.conv3d(x0, sd[p + "SingleConv1.conv.weight"], padding=1))
a = F.group_norm(v, 8, sd[p + "SingleConv2.groupnorm.weight"], sd[p + "SingleConv2.groupnorm.bias"], 1e-5)
ref = F.relu(F.conv3d(a, sd[p + "SingleConv2.conv.weight"], padding=1)).float()
dev = "cuda"
xg = x.float().to(dev)
N, D, H, W = 2, S, S, S
m = net.to(dev)
sc0 = m.encoders[0].basic_module.SingleConv1
sc1 = m.encoders[0].basic_module.SingleConv2
st = ops.volume_stats(xg)
scale, shift = ops.norm_finalize(st, D * H * W, sc0.groupnorm.weight, sc0.groupnorm.bias, 1, sc0.groupnorm.eps)
w0, w1 = sc0.conv.weight.detach(), sc1.conv.weight.detach()
in_sc, in_sh = scale.reshape(-1), shift.reshape(-1)
g1 = sc1.groupnorm


def report(tag, out):
    o = ops.ndhwc_to_ncdhw(out).cpu()
    e = (o - ref).abs()
    print(f"{tag}: mean {e.mean():.3e} max {e.max():.3e} signed mean {(o - ref).mean():.2e} "
          f"(ref rms {ref.pow(2).mean().sqrt():.3f})")


_, st1 = ops.conv3d_stem(xg, w0, None, in_sc, in_sh, relu_pre=True, store=False)
s1, h1 = ops.norm_finalize(st1, D * H * W, g1.weight, g1.bias, g1.num_groups, g1.eps)
an, _ = ops.conv3d_stem(xg, w0, None, in_sc, in_sh, s1, h1, relu_pre=True, want_stats=False)
out, _ = ops.conv3d_zfold(an, ops.pack_weights_zfold(w1), relu=True, want_stats=True)
report("two-pass stem, normalised store", out)
araw, st2 = ops.conv3d_stem(xg, w0, None, in_sc, in_sh, relu_pre=True)
s2, h2 = ops.norm_finalize(st2, D * H * W, g1.weight, g1.bias, g1.num_groups, g1.eps)
print("scale/shift identical between passes:", torch.equal(s1, s2), torch.equal(h1, h2))
out, _ = ops.conv3d_zfold_gn(araw, w1, s2, h2, relu=True, want_stats=True)
report("one pass, folded, raw store", out)
c = st2.double().sum(0)[..., 0].float() / float(D * H * W)
acen, _ = ops.conv3d_stem(xg, w0, None, in_sc, in_sh, torch.ones_like(c), -c, relu_pre=True, want_stats=False)
out, _ = ops.conv3d_zfold_gn(acen, w1, s2, torch.addcmul(h2, s2, c), relu=True, want_stats=True)
report("one pass, folded, centred store (exact mean)", out)
