"""Backbone executor: runs a UNet3D / TruncatedUNet3D / ConvNet parameter container on the CUDA
kernels (16-bit NDHWC activations -- fp16 by default, bf16 on request: ops.set_operand_dtype --, fp32 accumulation) and returns keypoints directly.

Layer schedule of the UNet path (reference: keymorph/unet3d/model.py:115-151):

    volume_stats -> norm_finalize                     GroupNorm(1) of the 1-channel input
    conv3d_stem x2 (GN on load, ReLU; pass 1 = stats, enc0.SingleConv1   1 -> 16      mma.sync TF32
                    pass 2 = next GN folded, store)
    norm_finalize, norm_apply, conv3d_tc(ReLU,stats)  every other SingleConv          tcgen05
    maxpool2_stats                                    encoder transitions (stats of the pooled map)
    norm_apply(src0 = skip, src1 = x)                 decoder: GN + nearest-upsample + concat fused
    conv3d_tc(taps=1, bias, CoM)                      final 1x1x1 conv + ReLU + centre of mass; the
                                                      heat map is never written unless asked for

GroupNorm statistics are produced by the epilogue of the kernel that writes the tensor, so no
tensor is read just to be measured.  Inference only (no autograd through the kernels).
"""
from __future__ import annotations

import torch

from . import ops


def _unwrap(module):
    # torch.nn.DataParallel(network) as in scripts/register.py:264
    return module.module if isinstance(module, torch.nn.DataParallel) else module


def backbone_engine(module):
    """Engine cached on the parameter container (None for foreign modules)."""
    from .net import ConvNet
    from .unet3d import _UNetBase
    m = _unwrap(module)
    if isinstance(m, _UNetBase):
        if m._engine is None:
            object.__setattr__(m, "_engine", UNetEngine(m))
        return m._engine
    if isinstance(m, ConvNet):
        if m._engine is None:
            object.__setattr__(m, "_engine", ConvNetEngine(m))
        return m._engine
    return None


class _WeightCache:
    """Packed (bf16, [tap][Cout][Cin]) copies of the conv weights, refreshed when a parameter
    changes (load_state_dict, .to(device), in-place updates bump `_version`)."""

    def __init__(self):
        self._packed = {}

    def get(self, key, param, pad_out_to=None, zfold=False):
        sig = (param.data_ptr(), param._version, str(param.device), tuple(param.shape), pad_out_to, ops.act_dtype())
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        w = param.detach()
        if pad_out_to is not None and w.shape[0] % pad_out_to:
            extra = pad_out_to - w.shape[0] % pad_out_to
            w = torch.cat([w, w.new_zeros((extra,) + tuple(w.shape[1:]))], 0)
        packed = ops.pack_weights_zfold_pair(w) if zfold == "pair" else \
            (ops.pack_weights_zfold(w) if zfold else ops.pack_weights(w))
        self._packed[key] = (sig, packed)
        return packed


def _padded(vec, multiple):
    if vec is None or vec.shape[0] % multiple == 0:
        return vec
    return torch.cat([vec, vec.new_zeros(multiple - vec.shape[0] % multiple)], 0)


class _EngineBase:
    def __init__(self, module):
        self.m = module
        self.weights = _WeightCache()
        self.debug = None      # set to a dict to capture every layer's raw output (tools/diag_*)

    def _dbg(self, key, t):
        if self.debug is not None:
            self.debug[key] = t.clone()

    def _check_input(self, x):
        if not x.is_cuda:
            raise ops._lib.KMError("keymorph_b200 backbones run on CUDA only (no CPU fallback)")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected (N,1,D,H,W) input, got {tuple(x.shape)}")
        p = next(self.m.parameters())
        if p.device != x.device:
            raise ValueError(f"input on {x.device} but backbone weights on {p.device}")
        if torch.is_grad_enabled() and any(q.requires_grad for q in self.m.parameters()) \
                and self.m.training:
            raise NotImplementedError("keymorph_b200 kernels are inference-only: call .eval() / "
                                      "torch.no_grad() (training is out of scope, SURVEY.md 3.4)")
        return x.float().contiguous()

    def keypoints(self, x, want_mass=False, want_feat=False):
        raise NotImplementedError

    def heatmap(self, x):
        return self.keypoints(x, want_feat=True)[2]


class UNetEngine(_EngineBase):
    def _single_conv(self, key, sc_mod, x, scale=None, shift=None):
        """GN -> conv -> ReLU of one SingleConv.  With scale / shift, x is the RAW activation: the norm is
        folded into the z-folded pair kernel where that kernel takes the layer, otherwise applied in place
        first.  Without them x is already normalised."""
        w = sc_mod.conv.weight
        _, D, H, W, Cin = x.shape
        use_zf2 = ops.USE_ZFOLD_PAIR and D * H * W >= 96 ** 3 \
            and (w.shape[0] == 32 or Cin % 64 == 0 or (scale is not None and ops.USE_GN_FOLD and ops.zfold_pair_cin32_enabled() and Cin % 32 == 0)) \
            and ops.zfold_pair_supported(Cin, w.shape[0], D, H, W) \
            and not ops.zfold_supported(Cin, w.shape[0], D, H, W)
        use_pair = not use_zf2 and not ops.zfold_supported(Cin, w.shape[0], D, H, W) and ops.USE_PAIR_CONV \
            and D * H * W >= 64 ** 3 and ops.pair_supported(Cin, w.shape[0], D, H, W)
        if scale is not None:
            if ops.USE_GN_FOLD and (use_zf2 or (use_pair and ops.USE_GN_FOLD_TC_PAIR)):
                fn = ops.conv3d_zfold_pair_gn if use_zf2 else ops.conv3d_tc_pair_gn
                out, stats = fn(x, w.detach(), scale, shift, relu=True, want_stats=True)
                self._dbg(key, out)
                return out, stats
            x = ops.norm_apply(x, scale, shift, out=x)
        x_norm = x
        if ops.zfold_supported(Cin, w.shape[0], D, H, W):
            # first tensor-core layer (16 -> 32 at full resolution): dz taps folded into MMA N
            out, stats = ops.conv3d_zfold(x_norm, self.weights.get(key + ".zf", w, zfold=True),
                                          relu=True, want_stats=True)
        elif use_zf2:
            # 32 -> 32, 64 -> 64 and 192 -> 64 at 128^3: dz folded into N = 3 Cout AND the weight rows split
            # over a CTA pair (32 -> 64 measures the same as the plain pair kernel and stays there)
            out, stats = ops.conv3d_zfold_pair(x_norm, self.weights.get(key + ".zf2", w, zfold="pair"),
                                               relu=True, want_stats=True)
        elif use_pair:
            # Cout in {64, 128}: two SMs per M = 256 MMA, half of the weight rows per SM (measured faster
            # from 64^3 up; below that there are too few brick groups per CTA pair)
            out, stats = ops.conv3d_tc_pair(x_norm, self.weights.get(key, w), relu=True, want_stats=True)
        else:
            out, stats, _ = ops.conv3d_tc(x_norm, self.weights.get(key, w), relu=True, want_stats=True)
        self._dbg(key, out)
        return out, stats

    @torch.no_grad()
    def keypoints(self, x, want_mass=False, want_feat=False):
        m = self.m
        x = self._check_input(x)
        N, _, D, H, W = x.shape
        enc = m.encoders

        def nvox(t):
            return t.shape[1] * t.shape[2] * t.shape[3]

        # ---- encoder 0: GN(1 group) -> conv 1->16 (stem) -> GN -> conv
        sc0 = enc[0].basic_module.SingleConv1
        st = ops.volume_stats(x)
        scale, shift = ops.norm_finalize(st, D * H * W, sc0.groupnorm.weight, sc0.groupnorm.bias,
                                         sc0.groupnorm.num_groups, sc0.groupnorm.eps)
        w0 = sc0.conv.weight.detach()
        in_sc, in_sh = scale.reshape(-1), shift.reshape(-1)
        sc1 = enc[0].basic_module.SingleConv2
        g1 = sc1.groupnorm
        w1 = sc1.conv.weight
        zf = ops.zfold_supported(w0.shape[0], w1.shape[0], D, H, W)
        fold = ops.USE_GN_FOLD and ops.gn_fold_stem_enabled() and zf and self.debug is None
        if fold:
            # ONE stem pass stores the raw ReLU'd map and its statistics; the next GroupNorm is folded into
            # the z-folded conv (per-sample scaled weights + border-class bias), so the map is never normalised
            # in memory and the stem is not run a second time
            a, st = ops.conv3d_stem(x, w0, None, in_sc, in_sh, relu_pre=True)
            scale, shift = ops.norm_finalize(st, D * H * W, g1.weight, g1.bias, g1.num_groups, g1.eps)
        else:
            # the stem runs twice: a statistics pass, then a store pass with the next GroupNorm folded in
            _, st = ops.conv3d_stem(x, w0, None, in_sc, in_sh, relu_pre=True, store=False)
            scale, shift = ops.norm_finalize(st, D * H * W, g1.weight, g1.bias, g1.num_groups, g1.eps)
            if self.debug is not None:
                self._dbg("enc0.c1", ops.conv3d_stem(x, w0, None, in_sc, in_sh, relu_pre=True,
                                                     want_stats=False)[0])
            a, _ = ops.conv3d_stem(x, w0, None, in_sc, in_sh, scale, shift, relu_pre=True,
                                   want_stats=False)
        # the truncated net never consumes the first encoder's full-resolution output as a skip:
        # then the z-folded kernel pools in its epilogue and the 32-channel full-resolution map
        # (2.1 GB for a 256^3 pair) is never written
        need_full0 = len(m.decoders) >= len(enc) - 1 or self.debug is not None or len(enc) < 2
        pooled0 = None
        pool0 = not need_full0 and zf and min(D, H, W) >= 2
        if fold:
            r = ops.conv3d_zfold_gn(a, w1.detach(), scale, shift, relu=True, want_stats=True, pool=pool0,
                                    store=not pool0)
            if pool0:
                pooled0 = (r[1], r[2])
                cur = cur_st = None
            else:
                cur, cur_st = r
        elif pool0:
            _, p0, st0 = ops.conv3d_zfold(a, self.weights.get("enc0.c2.zf", w1, zfold=True), relu=True,
                                          want_stats=True, pool=True, store=False)
            pooled0 = (p0, st0)
            cur = cur_st = None
        else:
            cur, cur_st = self._single_conv("enc0.c2", sc1, a)
        del a
        feats = [(cur, cur_st)]
        # ---- encoders 1..L-1: pool -> (GN, conv) x 2
        for i in range(1, len(enc)):
            dc = enc[i].basic_module
            if i == 1 and pooled0 is not None:
                p, st = pooled0
            else:
                p, st = ops.maxpool2_stats(cur)
            g = dc.SingleConv1.groupnorm
            scale, shift = ops.norm_finalize(st, nvox(p), g.weight, g.bias, g.num_groups, g.eps)
            c1, st = self._single_conv(f"enc{i}.c1", dc.SingleConv1, p, scale, shift)
            del p
            g = dc.SingleConv2.groupnorm
            scale, shift = ops.norm_finalize(st, nvox(c1), g.weight, g.bias, g.num_groups, g.eps)
            cur, cur_st = self._single_conv(f"enc{i}.c2", dc.SingleConv2, c1, scale, shift)
            del c1
            feats.append((cur, cur_st))
        # the truncated net never consumes the first encoder's full-resolution output as a skip
        skips = feats[:-1][::-1][:len(m.decoders)]
        del feats
        # ---- decoders: GN(upsample(x) ++ skip) fused into one pass, then (conv, GN, conv)
        for j, dec in enumerate(m.decoders):
            dc = dec.basic_module
            skip, skip_st = skips[j]
            g = dc.SingleConv1.groupnorm
            exact = all(skip.shape[d] == 2 * cur.shape[d] for d in (1, 2, 3))
            if exact:
                scale, shift = ops.norm_finalize(skip_st, nvox(skip), g.weight, g.bias,
                                                 g.num_groups, g.eps, stats1=cur_st,
                                                 count1=nvox(cur), rep1=8.0)
                w1 = dc.SingleConv1.conv.weight
                Cs, Cu = skip.shape[-1], cur.shape[-1]
                sd = skip.shape[1:4]
                Cout, vox = w1.shape[0], sd[0] * sd[1] * sd[2]
                # (1) the upsampled half on the COARSE lattice: its 27 fine taps are 8 pre-summed taps per output
                # parity class (8/27 of the MMAs, conv_up2.cu) and the upsampled tensor is never written; the skip
                # half's kernel reads the skip raw (joint GroupNorm folded into both) and adds those partial sums
                # before bias / ReLU / statistics
                split = None
                if ops.USE_GN_FOLD and ops.USE_COARSE_UPCONV and Cs % 64 == 0 \
                        and ops.up2_supported(Cu, Cout, *cur.shape[1:4]):
                    if ops.USE_ZFOLD_PAIR and vox >= 96 ** 3 and Cout == 64 and ops.zfold_pair_supported(Cs, Cout, *sd):
                        split = "zfold_pair"
                    elif ops.USE_PAIR_CONV and vox >= 64 ** 3 and ops.pair_supported(Cs, Cout, *sd):
                        split = "tc_pair"
                # (2) the concat read in place through two tensor maps: only the upsampled half is written (raw)
                cat_in_place = ops.USE_GN_FOLD and ops.USE_ZFOLD_PAIR and vox >= 96 ** 3 \
                    and (Cout == 32 or (Cs % 64 == 0 and Cu % 64 == 0)) and Cs % 32 == 0 and Cu % 32 == 0 \
                    and ops.zfold_pair_supported(Cs + Cu, Cout, *sd)
                if split or cat_in_place:
                    if split:
                        part = ops.conv3d_up2_gn(cur, w1.detach(), scale, Cs)
                        c1, st = ops.conv3d_zfold_pair_gn_add(skip, w1.detach(), scale, shift, part, relu=True,
                                                              want_stats=True, kernel=split)
                        del part
                    else:
                        up = ops.upsample2(cur)
                        c1, st = ops.conv3d_zfold_pair_gn(skip, w1.detach(), scale, shift, relu=True,
                                                          want_stats=True, x1=up)
                        del up
                    self._dbg(f"dec{j}.c1", c1)
                    skips[j] = None
                    g = dc.SingleConv2.groupnorm
                    scale, shift = ops.norm_finalize(st, nvox(c1), g.weight, g.bias, g.num_groups, g.eps)
                    cur, cur_st = self._single_conv(f"dec{j}.c2", dc.SingleConv2, c1, scale, shift)
                    del c1
                    continue
                cat = ops.norm_apply(skip, scale, shift, src1=cur)
            else:
                C = skip.shape[-1] + cur.shape[-1]
                ones = torch.ones((N, C), device=x.device)
                cat = ops.norm_apply(skip, ones, torch.zeros_like(ones), src1=cur)
                st = ops.channel_stats(cat)
                scale, shift = ops.norm_finalize(st, nvox(cat), g.weight, g.bias, g.num_groups,
                                                 g.eps)
                cat = ops.norm_apply(cat, scale, shift, out=cat)
            skips[j] = None
            c1, st = self._single_conv(f"dec{j}.c1", dc.SingleConv1, cat)
            del cat
            g = dc.SingleConv2.groupnorm
            scale, shift = ops.norm_finalize(st, nvox(c1), g.weight, g.bias, g.num_groups, g.eps)
            cur, cur_st = self._single_conv(f"dec{j}.c2", dc.SingleConv2, c1, scale, shift)
            del c1
        # ---- final 1x1x1 conv (+bias) fused with ReLU + centre of mass
        K = m.final_conv.out_channels
        if not want_feat and K <= 512 and cur.shape[-1] <= 256:
            # transposed formulation: one epilogue thread per keypoint channel, no heat map at all
            wp = self.weights.get("final.t", m.final_conv.weight, pad_out_to=128)
            heat, com = None, ops.conv1x1_com(cur, wp, _padded(m.final_conv.bias.detach(), 128))
        else:
            wp = self.weights.get("final", m.final_conv.weight, pad_out_to=32)
            bias = _padded(m.final_conv.bias.detach(), 32)
            heat, _, com = ops.conv3d_tc(cur, wp, bias=bias, want_com=True, store=want_feat)
        pts, mass = ops.com_finalize(com, return_mass=True)
        pts, mass = pts[:, :K], mass[:, :K]
        feat = ops.ndhwc_to_ncdhw(heat)[:, :K] if want_feat else None
        return pts, (mass if want_mass else None), feat


class ConvNetEngine(_EngineBase):
    @torch.no_grad()
    def keypoints(self, x, want_mass=False, want_feat=False):
        m = self.m
        x = self._check_input(x)
        N = x.shape[0]
        blocks = m.blocks()
        K = blocks[-1].conv.out_channels
        b0 = blocks[0]
        w0, bias0 = b0.conv.weight.detach(), b0.conv.bias.detach()
        # block 1 (stem): statistics pass, then a store pass with InstanceNorm + ReLU folded in
        _, st = ops.conv3d_stem(x, w0, bias0, store=False)
        raw = None
        for b, blk in enumerate(blocks):
            if b == 0:
                C, nv = w0.shape[0], x.shape[2] * x.shape[3] * x.shape[4]
            else:
                C = raw.shape[-1]
                nv = raw.shape[1] * raw.shape[2] * raw.shape[3]
            if blk.norm is not None:
                scale, shift = ops.norm_finalize(st, nv, None, None, C, blk.norm.eps)
            else:
                scale = torch.ones((N, C), device=x.device)
                shift = torch.zeros_like(scale)
            if b == 0 and len(blocks) > 1 and not blk.down_sample:
                y, _ = ops.conv3d_stem(x, w0, bias0, None, None, scale, shift, relu_post=True,
                                       want_stats=False)
            else:
                if b == 0:
                    raw, _ = ops.conv3d_stem(x, w0, bias0, want_stats=False)
                if b == len(blocks) - 1:
                    heat = ops.norm_apply(raw, scale, shift, relu=True)
                    break
                y = ops.norm_apply(raw, scale, shift, relu=True, pool=blk.down_sample,
                                   out=None if blk.down_sample else raw)
            nxt = blocks[b + 1]
            last = b + 1 == len(blocks) - 1
            w = nxt.conv.weight
            Cin, Cout = w.shape[1], w.shape[0]
            Dn, Hn, Wn = y.shape[1:4]
            # An InstanceNorm follows the conv: a per-channel constant cancels exactly in (v - mean) / std, so
            # the conv bias is dropped there and the bias-free CTA-pair kernels of the UNet path take the layer
            # (z-folded pair for Cout 64 with Cin % 64 == 0 at >= 96^3, plain pair for Cout 64 / 128)
            no_bias = nxt.norm is not None and not last
            if no_bias and ops.USE_ZFOLD_PAIR and Dn * Hn * Wn >= 96 ** 3 and Cout == 64 and Cin % 64 == 0 \
                    and ops.zfold_pair_supported(Cin, Cout, Dn, Hn, Wn):
                raw, st = ops.conv3d_zfold_pair(y, self.weights.get(f"block{b + 2}.zf2", w, zfold="pair"), relu=False,
                                                want_stats=True)
            elif no_bias and ops.USE_PAIR_CONV and Dn * Hn * Wn >= 64 ** 3 and ops.pair_supported(Cin, Cout, Dn, Hn, Wn):
                raw, st = ops.conv3d_tc_pair(y, self.weights.get(f"block{b + 2}", w), relu=False, want_stats=True)
            else:
                wp = self.weights.get(f"block{b + 2}", w, pad_out_to=16 if last else None)
                bias = nxt.conv.bias.detach()
                raw, st, _ = ops.conv3d_tc(y, wp, bias=_padded(bias, 16) if last else bias, relu=False,
                                           want_stats=True)
            del y
        feat = ops.ndhwc_to_ncdhw(heat)[:, :K].contiguous()
        pts, mass = ops.com3d(feat, ij=True, return_mass=True)
        return pts, (mass if want_mass else None), (feat if want_feat else None)
