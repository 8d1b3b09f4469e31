/*
 * km_b200.h -- C-ABI of libkm_b200.so, the B200 (sm_100a) registration engine that sits
 * underneath the Python surface of alanqrwang/keymorph.
 *
 * The reference has no FFI layer of its own: its "plugin boundary" is the
 * set of torch call sites listed below.  Every entry point here replaces one of those call sites
 * and is what a ctypes / pybind stub on the reference side binds.
 *
 * Conventions
 *   - plain pointers + explicit sizes, no torch types; all pointers are DEVICE pointers unless
 *     the parameter name ends in _host;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     and never allocates: the caller owns outputs and workspaces (sizes from *_workspace_bytes);
 *   - return value 0 = success, negative = KM_E*; km_last_error() returns a thread-local message;
 *   - volumes are (N, C, D, H, W) fp32 contiguous ("NCDHW") at the API surface, activations
 *     inside the backbone are (N, D, H, W, C) bf16 ("NDHWC");
 *   - points are (N, K, 3) fp32 in the reference's 'ij' order (z, y, x), normalised to [-1, 1];
 *     flow fields are (N, D, H, W, 3) fp32 in grid_sample's (x, y, z) order.
 */
#ifndef KM_B200_H
#define KM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KM_OK 0
#define KM_EINVAL (-1)   /* bad argument (shape, alignment, unsupported size) */
#define KM_ECUDA (-2)    /* CUDA runtime / driver error, text in km_last_error() */
#define KM_ENOSUPPORT (-3)

typedef void* km_stream_t; /* cudaStream_t */

int km_version(void);
const char* km_last_error(void);
/* number of SMs of the current device (grid sizing for the persistent kernels) */
int km_sm_count(void);
/* runtime options: key KM_OPT_TPS_FAST (default 1): evaluate the TPS radial basis of the dense
 * flow field with lg2.approx / rsqrt.approx instead of logf / sqrtf. */
#define KM_OPT_TPS_FAST 1
/* key KM_OPT_CONV_FORCE_GENERIC (default 0): make km_conv3d_tc use the one-TMA-box-per-tap data
 * path even where the x-halo-reuse path applies (A/B testing of the two paths). */
#define KM_OPT_CONV_FORCE_GENERIC 2
/* key KM_OPT_CONV_NO_RESIDENT_WEIGHTS (default 0): stream the weights of small layers per tile
 * instead of keeping them in shared memory (A/B testing). */
#define KM_OPT_CONV_NO_RESIDENT_WEIGHTS 3
/* key KM_OPT_CONV_MAX_BRICKS (default 4): how many x-adjacent 16x8 bricks may share one weight
 * fetch in km_conv3d_tc (1, 2 or 4). */
#define KM_OPT_CONV_MAX_BRICKS 4
/* key KM_OPT_CONV_NO_EPILOGUE_BATCH (default 0): stage one brick per epilogue round (A/B). */
#define KM_OPT_CONV_NO_EPILOGUE_BATCH 5
/* key KM_OPT_CONV_HALO_AXIS (default 2): axis along which the three taps share one TMA box in
 * km_conv3d_tc: 2 = y (shared-memory atoms are x-runs, contiguous in global memory), 1 = x. */
#define KM_OPT_CONV_HALO_AXIS 6
/* key KM_OPT_TPS_SINGLE_CTA (default 0): how km_tps_fit solves the (K+4)^2 system.  0 = ONE cooperative
 * launch (assembly + blocked Gauss-Jordan with group barriers + final division); 1 = the un-blocked one-CTA
 * LU; 2 = the same blocked elimination as ~100 dependent launches (A/B; bit-identical to 0). */
#define KM_OPT_TPS_SINGLE_CTA 7
/* key KM_OPT_CONV_INTERLEAVE_BRICKS (default 0): km_conv3d_tc issues the MMAs of the bricks that
 * share a weight slice round-robin (consecutive tcgen05.mma accumulate into different TMEM tiles)
 * instead of brick by brick (A/B testing; measured slower: the issue loop needs more instructions). */
#define KM_OPT_CONV_INTERLEAVE_BRICKS 8
/* key KM_OPT_CONV_TWO_ISSUERS (default 128): layers whose output-channel block is at most this wide
 * split the bricks of a group between two MMA issuer warps (0 = always one issuer; A/B testing). */
#define KM_OPT_CONV_TWO_ISSUERS 9
/* key KM_OPT_ZF2_TWO_BRICKS (default 0): km_conv3d_zfold_pair processes two y-adjacent bricks per unit
 * for the 64 -> 64 shapes (half the weight traffic, but only one TMEM set: measured slower) instead of
 * one (A/B testing). */
#define KM_OPT_ZF2_TWO_BRICKS 10
/* key KM_OPT_TPS_PACKED (default 1): the dense TPS field issues its per-term FP32 arithmetic as packed
 * f32x2 instructions on voxel pairs (A/B switch; results are bit-identical to the scalar path). */
#define KM_OPT_TPS_PACKED 11
/* key KM_OPT_TPS_VPT (default 0 = chosen from W): voxels per thread of the dense TPS field (2, 4 or 8). */
#define KM_OPT_TPS_VPT 12
/* key KM_OPT_OPERAND_FP16 (default 1): element type of the 16-bit activation / weight tensors of the
 * backbone ("bf16" in the names below means "16-bit operand"): 1 = IEEE fp16, the AMP dtype of the
 * reference (keymorph/model.py:175-177; 11-bit significand), 0 = bf16 (exponent range of fp32, 8-bit
 * significand).  tcgen05 kind::f16 runs both at the same rate.  Tensors written under one setting must be
 * consumed under the same setting (packed weights included). */
#define KM_OPT_OPERAND_FP16 13
/* key KM_OPT_WARP_TILE (default 1): km_warp_loss (affine / grid coordinates) and km_grid_sample3d in bilinear
 * mode gather from TMA-staged shared-memory tiles (warp_tile.cu) instead of issuing 8 global loads per voxel
 * (0 = the direct-gather kernels, A/B; identical results). */
#define KM_OPT_WARP_TILE 14
int km_set_option(int key, int value);
/* 1 when the 16-bit tensors of the backbone are fp16, 0 when they are bf16 (KM_OPT_OPERAND_FP16) */
int km_operand_is_fp16(void);

/* ------------------------------------------------------------------------------------------ *
 * Warp: keymorph/utils.py:14-21  align_img -> F.grid_sample(mode, padding_mode="border",
 * align_corners=False)
 * ------------------------------------------------------------------------------------------ */
#define KM_INTERP_BILINEAR 0
#define KM_INTERP_NEAREST 1

/* out[n,c,d,h,w] = sample(x[n,c], grid[n,d,h,w,:]);  x: (N,C,Di,Hi,Wi), grid: (N,Do,Ho,Wo,3) */
int km_grid_sample3d(const float* x, const float* grid, float* out, int N, int C, int Di, int Hi,
                     int Wi, int Do, int Ho, int Wo, int mode, km_stream_t stream);

/* keymorph/transformations.py:37-79,98-114 (AffineTransform.get_flow_field) with
 * keymorph/utils.py:387-398 (uniform_norm_grid) generated in-register.
 * mat: (N,3,4) row-major, rows/cols in (z,y,x) order = inverse_transform_matrix[:, :3, :].
 * grid: (N,D,H,W,3) in (x,y,z) order (the reference's .flip(-1)). */
int km_flow_field_affine(const float* mat, float* grid, int N, int D, int H, int W,
                         km_stream_t stream);

/* keymorph/keypoint_aligners.py:365-449 (TPS.get_flow_field / transform_points).
 * ctrl: (N,K,3) control points (points_f), theta: (N,K+4,3) from km_tps_fit. */
int km_flow_field_tps(const float* ctrl, const float* theta, float* grid, int N, int K, int D,
                      int H, int W, km_stream_t stream);

/* keymorph/transformations.py:81-114: out[n,p,:] = mat[n] (3x4) * [pts[n,p,:]; 1] */
int km_points_transform_affine(const float* mat, const float* pts, float* out, int N, int P,
                               km_stream_t stream);
/* keymorph/keypoint_aligners.py:399-433 (TPS.transform_points) on arbitrary points */
int km_points_transform_tps(const float* ctrl, const float* theta, const float* pts, float* out,
                            int N, int K, int P, km_stream_t stream);

/* Fused warp + loss: coordinates are generated from the affine matrix (or the
 * TPS parameters, or read from `grid`), the moving volume is gathered (trilinear, border), the
 * warped volume is optionally stored (out may be NULL) and the loss partial sums are reduced
 * deterministically in two stages.
 *   sums layout per (n, c): [sum (a-f)^2, sum a*f, sum a*a, sum f*f]  (fp64, 4 doubles)
 * `fixed` may be NULL (then only the warp is performed and sums are untouched).
 * `grid_out` (N,D,H,W,3), affine / TPS coordinate modes only, may be NULL: the flow field the
 * coordinates were generated from, in grid_sample's (x,y,z) order -- identical to
 * km_flow_field_affine / km_flow_field_tps -- written by the same pass, so that a registration that
 * must return the grid (keymorph/model.py:252-262) never reads it back.
 * workspace: km_warp_loss_workspace_bytes(N, C). */
#define KM_COORD_AFFINE 0
#define KM_COORD_TPS 1
#define KM_COORD_GRID 2
size_t km_warp_loss_workspace_bytes(int N, int C);
int km_warp_loss(int coord_mode, const float* mat_or_ctrl, const float* theta, int K,
                 const float* grid, const float* moving, const float* fixed, float* out,
                 float* grid_out, double* sums, void* workspace, int N, int C, int D, int H, int W,
                 int mode, km_stream_t stream);

/* Label-map fast path for segmentation warp + Dice.  Replaces one_hot -> align_img
 * -> DiceLoss(soft) + DiceLoss(hard) of scripts/pairwise_register_eval.py:99-108,156-157,303-321 on
 * uint8 label volumes (N,D,H,W) with values < C <= 255: the one-hot volumes are never built.
 *   soft_sums (N,C,4) fp64 = [sum (a-t)^2, sum a*t, sum a*a, sum t*t] with a = trilinear warp of
 *   one_hot(labels_m)[c] (same per-voxel arithmetic as km_warp_loss), t = one_hot(labels_f)[c];
 *   hard_sums (N,C,4) fp64: the same with a replaced by one_hot(argmax_c a) (first maximum wins);
 *   labels_out (N,D,H,W) uint8, may be NULL: argmax_c a, i.e. the hard warped segmentation.
 * coord_mode KM_COORD_AFFINE (mat (N,3,4)) or KM_COORD_GRID (grid (N,D,H,W,3)).
 * workspace: km_warp_labels_workspace_bytes(N, C). */
size_t km_warp_labels_workspace_bytes(int N, int C);
int km_warp_labels_dice(int coord_mode, const float* mat, const float* grid, const uint8_t* labels_m,
                        const uint8_t* labels_f, uint8_t* labels_out, double* soft_sums,
                        double* hard_sums, void* workspace, int N, int C, int D, int H, int W,
                        km_stream_t stream);

/* keymorph/loss_ops.py:161-247 (_jacobian_determinant, jdstd, jdlessthan0): statistics of
 * the Jacobian determinant of a 3-component fp32 field over the interior cropped by 2 voxels.  The
 * field is addressed by element strides (n, component, z, y, x), so the (N,3,D,H,W) tensor the
 * reference passes and the (N,D,H,W,3) grid it is a permuted view of are both read in place.
 *   out (N,4) fp64 = [std (ddof 0), count(det <= 0), mean, number of interior voxels]
 * workspace: km_jacobian_stats_workspace_bytes(N).  Requires D, H, W > 4. */
size_t km_jacobian_stats_workspace_bytes(int N);
int km_jacobian_stats(const float* field, long long stride_n, long long stride_c, long long stride_z,
                      long long stride_y, long long stride_x, double* out, void* workspace, int N,
                      int D, int H, int W, km_stream_t stream);

/* keymorph/loss_ops.py:120-157 (_surfd, hausdorff_distance): symmetric Hausdorff
 * distance between the surfaces of two binary volumes (voxel != 0), surface = volume minus its
 * 6-connected binary erosion (array border counts as background), exact anisotropic Euclidean
 * distance transform with voxel sizes (sz, sy, sx) along (D, H, W).  The reference calls it on channel 0
 * of two (N,C,D,H,W) one-hot tensors with sampling (1.25, 1.25, 10): pass the channel-0 pointers and the
 * batch strides in elements.
 *   out (N,2) fp64 = [distance, empty] -- empty = 1 (distance = -1) when either surface has no voxel,
 *   a case in which scipy's transform is undefined and the reference returns garbage.
 * workspace: km_hausdorff_workspace_bytes(D, H, W) = 10 bytes per voxel. */
size_t km_hausdorff_workspace_bytes(int D, int H, int W);
int km_hausdorff(const float* a, const float* b, long long stride_a, long long stride_b, int N, int D,
                 int H, int W, float sz, float sy, float sx, double* out, void* workspace,
                 km_stream_t stream);

/* keymorph/loss_ops.py:9-13 and :16-63.  Elementwise-pair statistics of two (N,C,M) fp32 tensors:
 * sums[n,c,:] = [sum (p-t)^2, sum p*t, sum p*p, sum t*t] (fp64).  With hard != 0 pred is replaced
 * by one_hot(argmax_c pred) (first maximum wins, like torch.argmax) before the products.
 * workspace: km_pair_stats_workspace_bytes(N, C, M, hard) (hard Dice keeps the int32 label map
 * behind the partial sums). */
size_t km_pair_stats_workspace_bytes(int N, int C, long long M, int hard);
int km_pair_stats(const float* pred, const float* target, double* sums, void* workspace, int N,
                  int C, long long M, int hard, km_stream_t stream);
/* hard-Dice labels: labels[n,m] = argmax_c pred[n,c,m] (int32, first max wins) */
int km_argmax_channels(const float* pred, int32_t* labels, int N, int C, long long M,
                       km_stream_t stream);

/* ------------------------------------------------------------------------------------------ *
 * Keypoint layer: keymorph/layers.py:92-134 CenterOfMass3d (+ keymorph/model.py:95-109 power)
 * heat: (N,K,D,H,W) fp32.  points: (N,K,3) in 'ij' (z,y,x) order when ij != 0 else (x,y,z).
 * mass (optional, may be NULL): (N,K) = sum relu(heat).
 * ------------------------------------------------------------------------------------------ */
size_t km_com3d_workspace_bytes(int N, int K);
int km_com3d(const float* heat, float* points, float* mass, void* workspace, int N, int K, int D,
             int H, int W, int ij, km_stream_t stream);

/* ------------------------------------------------------------------------------------------ *
 * Closed-form aligners: keymorph/keypoint_aligners.py:76-114 (affine), :151-213 (rigid),
 * keymorph/transformations.py:10-35 (square + inverse).
 * x, y: (N,K,3); w: (N,K) or NULL.  Fits y ~ A [x;1].
 * A44: (N,4,4) = square(A); A44_inv: (N,4,4) = inverse(square(A)); status: (N) int32, nonzero
 * when the normal matrix (affine) or A (inverse) is singular (the reference raises LinAlgError).
 * ------------------------------------------------------------------------------------------ */
int km_fit_affine(const float* x, const float* y, const float* w, float* A44, float* A44_inv,
                  int32_t* status, int N, int K, km_stream_t stream);
int km_fit_rigid(const float* x, const float* y, const float* w, float* A44, float* A44_inv,
                 int32_t* status, int N, int K, km_stream_t stream);
/* inverse of N 4x4 matrices (keymorph/transformations.py:25,28) */
int km_inverse44(const float* m, float* inv, int32_t* status, int N, km_stream_t stream);

/* ------------------------------------------------------------------------------------------ *
 * TPS fit: keymorph/keypoint_aligners.py:276-363 (fit / fit_dim / d / u).
 * c_src, c_dst: (N,K,3); lmbda: (N); w: (N,K) or NULL; theta: (N,K+4,3) fp32.
 * Assembles [[U + lmbda*I, P],[P^T, 0]] and solves for the three right-hand sides at once with a
 * partially pivoted LU held in `workspace` (km_tps_fit_workspace_bytes(N, K)).
 * ------------------------------------------------------------------------------------------ */
size_t km_tps_fit_workspace_bytes(int N, int K);
int km_tps_fit(const float* c_src, const float* c_dst, const float* lmbda, const float* w,
               float* theta, int32_t* status, void* workspace, int N, int K, km_stream_t stream);

/* ------------------------------------------------------------------------------------------ *
 * Backbone (keymorph/unet3d/buildingblocks.py:39-132, keymorph/layers.py:137-187,
 * keymorph/net.py:7-36).  bf16 NDHWC activations, fp32 accumulation.
 * ------------------------------------------------------------------------------------------ */
/* fp32 (Cout,Cin,3,3,3) [or (Cout,Cin,1,1,1) with taps=1] -> bf16 [tap][Cout][Cin];
 * in_scale (Cin) optional per-input-channel multiplier (NULL = 1). */
int km_pack_weights(const float* w, void* packed_bf16, int Cout, int Cin, int taps,
                    km_stream_t stream);

/* per-(n,c) sum / sum-of-squares partials -> GroupNorm / InstanceNorm scale+shift.
 * stats layouts: (nparts, N, C, 2) fp32 partials written by the producing kernel. */
/* y = a*x + b with a = gamma*rstd, b = beta - mean*rstd*gamma per (n, c).
 * Two sources (for the decoder concat): channels [0,C0) come from stats0 with count0 elements per
 * channel, channels [C0,C0+C1) from stats1 with count1 elements, each weighted by rep (x8 for the
 * nearest-upsampled source: rep1 = 8).  groups: number of GroupNorm groups over C0+C1 channels
 * (groups == C0+C1 -> InstanceNorm).  gamma/beta may be NULL (no affine). */
int km_norm_finalize(const float* stats0, int nparts0, int C0, double count0,
                     const float* stats1, int nparts1, int C1, double count1, double rep1,
                     const float* gamma, const float* beta, int groups, float eps, float* scale,
                     float* shift, int N, km_stream_t stream);

/* per-(n,c) sum / sumsq partials of a bf16 NDHWC tensor: stats (nparts,N,C,2),
 * nparts = km_pool_nparts() (used when the decoder's upsample is not an exact x2). */
int km_channel_stats(const void* x, float* stats, int N, int C, long long nvox,
                     km_stream_t stream);

/* out[n,v,c] = act(scale[n,c]*src[n,v',c] + shift[n,c]) in bf16 NDHWC.
 * Source 0 (C0 channels) is sampled at the output resolution; source 1 (C1 channels, may be NULL /
 * C1 = 0) is nearest-upsampled from (D1,H1,W1) (keymorph/unet3d/buildingblocks.py:464-475,580-582)
 * and concatenated after source 0.  relu != 0 applies max(.,0). pool != 0 applies MaxPool3d(2)
 * on source 0 after the activation (output dims are then D/2,H/2,W/2; keymorph/layers.py:176-187). */
int km_norm_apply(const void* src0, int C0, const void* src1, int C1, int D1, int H1, int W1,
                  const float* scale, const float* shift, void* out, int N, int D, int H, int W,
                  int relu, int pool, km_stream_t stream);

/* MaxPool3d(2) (keymorph/unet3d/buildingblocks.py:363,387) on bf16 NDHWC + per-(n,c) stats of the
 * pooled tensor.  stats: (nparts,N,C,2) with nparts = km_pool_nparts(). */
int km_pool_nparts(void);
int km_maxpool2_stats(const void* src, void* out, float* stats, int N, int C, int D, int H, int W,
                      km_stream_t stream);

/* sum / sumsq of an fp32 volume per sample (GroupNorm of the 1-channel input image).
 * stats: (nparts,N,1,2), nparts = km_pool_nparts(). */
int km_volume_stats(const float* x, float* stats, int N, long long M, km_stream_t stream);

/* stem: 3x3x3 conv (pad 1) of the 1-channel fp32 volume on the warp-level tensor-core path (TF32
 * operands, fp32 accumulate) -- keymorph/unet3d/buildingblocks.py:39-132 (encoder 0, first SingleConv)
 * and keymorph/layers.py:137-187 (ConvNet block 1).  w: fp32 (Cout,1,3,3,3), Cout in {16,32}.
 *   v = conv(in_scale[n]*x + in_shift[n]) + bias      (zero padding of the NORMALISED volume;
 *                                                      in_scale/in_shift/bias may be NULL)
 *   v = relu(v)                                       if relu_pre
 *   stats (nparts,N,Cout,2) += [sum v, sum v^2]       if stats != NULL, nparts = km_stem_nparts()
 *   out = bf16(act(out_scale[n,c]*v + out_shift[n,c]))  if out != NULL (NDHWC; act = relu if relu_post;
 *                                                      out_scale/out_shift (N,Cout) may be NULL)
 * The layer is meant to run twice (statistics pass with out == NULL, then the store pass with the
 * following layer's normalisation folded in) instead of normalising its 2*Cout B/voxel output in
 * a separate pass. */
int km_stem_nparts(int N, int D, int H, int W);
int km_conv3d_stem(const float* x, const float* w, const float* bias, const float* in_scale,
                   const float* in_shift, const float* out_scale, const float* out_shift, void* out,
                   float* stats, int N, int Cout, int D, int H, int W, int relu_pre, int relu_post,
                   km_stream_t stream);

/* tcgen05 implicit-GEMM convolution, kernel 3x3x3 pad 1 (taps = 27) or 1x1x1 (taps = 1).
 *   x: bf16 NDHWC (N,D,H,W,Cin), Cin % 16 == 0;  wp: bf16 [tap][Cout][Cin] from km_pack_weights;
 *   out: bf16 NDHWC (N,D,H,W,Cout), Cout % 16 == 0 (may be NULL with KM_CONV_COM);
 *   bias: fp32 (Cout) or NULL;  flags: KM_CONV_RELU | KM_CONV_STATS | KM_CONV_COM.
 *   stats: (nparts,N,Cout,2) fp32 partial sum/sumsq of the stored values (KM_CONV_STATS);
 *   com:   (nparts,N,Cout,4) fp32 partial [sum h, sum h*lz, sum h*ly, sum h*lx] of h = relu(out)
 *          (KM_CONV_COM; the heat map is not written when out == NULL);
 *   nparts = km_conv_nparts() (one partial per persistent CTA). */
#define KM_CONV_RELU 1
#define KM_CONV_STATS 2
#define KM_CONV_COM 4
int km_conv_nparts(void);
int km_conv3d_tc(const void* x, const void* wp, const float* bias, void* out, float* stats,
                 float* com, int N, int Cin, int Cout, int D, int H, int W, int taps, int flags,
                 km_stream_t stream);

/* "z-folded" tcgen05 convolution for the first tensor-core layer of the UNet backbones (Cin = 16,
 * Cout = 32, 3x3x3, pad 1, no bias; keymorph/unet3d/buildingblocks.py:50-52): the three dz taps are
 * folded into the MMA N dimension and the output planes live as a rotating ring in TMEM while the
 * CTA walks along z (csrc/conv_zf.cu).  Same tensors / flags / stats as km_conv3d_tc (KM_CONV_RELU,
 * KM_CONV_STATS); wz comes from km_pack_weights_zfold (fp32 (Cout,Cin,3,3,3) -> bf16
 * [rotation][dx][dy][3*Cout][Cin], km_pack_weights_zfold_bytes bytes).
 * pooled (N,D/2,H/2,W/2,Cout), may be NULL: MaxPool3d(2) of the activated output
 * (keymorph/unet3d/buildingblocks.py:363,387) taken in the epilogue's registers; `stats` then describe
 * the POOLED tensor and `out` may be NULL (the truncated UNet never reads the full-resolution map). */
int km_conv3d_zfold_supported(int Cin, int Cout, int D, int H, int W);
size_t km_pack_weights_zfold_bytes(int Cout, int Cin);
int km_pack_weights_zfold(const float* w, void* packed, int Cout, int Cin, km_stream_t stream);
int km_conv3d_zfold(const void* x, const void* wz, void* out, void* pooled, float* stats, int N, int Cin,
                    int Cout, int D, int H, int W, int flags, km_stream_t stream);

/* The same layer with the preceding GroupNorm folded in (buildingblocks.py:50-52, order "gcr": nothing
 * non-linear sits between the norm and the conv).  x is the RAW bf16 activation; w the fp32
 * (Cout,Cin,3,3,3) weights; scale / shift (N,Cin) the per-sample GroupNorm coefficients of
 * km_norm_finalize.  conv(scale x + shift) with zero padding of the normalised input is evaluated as
 * conv_{w scale}(x) + bias[class(voxel)][cout], class = which taps fall outside the volume: a fold
 * kernel writes per-sample bf16 weights and the 36 x Cout bias table into the workspace, then the
 * z-folded kernel runs with its CTAs split evenly between the samples, each keeping its sample's weights
 * resident.  Removes the normalisation pass over the activation (for the
 * stem: its second run).  workspace: km_conv3d_zfold_gn_workspace_bytes(N), 16-byte aligned. */
size_t km_conv3d_zfold_gn_workspace_bytes(int N);
int km_conv3d_zfold_gn(const void* x, const float* w, const float* scale, const float* shift, void* out,
                       void* pooled, float* stats, void* workspace, int N, int Cin, int Cout, int D, int H,
                       int W, int flags, km_stream_t stream);

/* 2-CTA (cta_group::2) variant of km_conv3d_tc for the 3x3x3 layers with Cout in {64, 128} and
 * Cin % 32 == 0 (csrc/conv_tc2.cu): two SMs execute one M = 256 MMA, each supplying its own 128
 * activation rows and half of the weight rows, which halves the weight bytes read from shared
 * memory per SM (the bound of these layers).  Same tensors, weights (km_pack_weights, taps = 27),
 * flags (KM_CONV_RELU | KM_CONV_STATS) and statistics layout as km_conv3d_tc, no bias. */
int km_conv3d_tc_pair_supported(int Cin, int Cout, int D, int H, int W);
int km_conv3d_tc_pair(const void* x, const void* wp, void* out, float* stats, int N, int Cin, int Cout,
                      int D, int H, int W, int flags, km_stream_t stream);

/* km_conv3d_tc_pair with the preceding GroupNorm folded in (see km_conv3d_zfold_gn): RAW bf16 input,
 * fp32 (Cout,Cin,3,3,3) weights, scale / shift (N,Cin).  workspace:
 * km_conv3d_tc_pair_gn_workspace_bytes(N, Cin, Cout), 256-byte aligned. */
size_t km_conv3d_tc_pair_gn_workspace_bytes(int N, int Cin, int Cout);
int km_conv3d_tc_pair_gn(const void* x, const float* w, const float* scale, const float* shift, void* out,
                         float* stats, void* workspace, int N, int Cin, int Cout, int D, int H, int W,
                         int flags, km_stream_t stream);

/* z-folded AND 2-CTA convolution for the Cout = 64, Cin % 64 == 0 layers (csrc/conv_zf2.cu): the dz
 * taps are folded into N = 192 (TMEM ring of output planes as in km_conv3d_zfold) and the 192 weight
 * rows are split between the two CTAs of a pair (as in km_conv3d_tc_pair), weights streamed by TMA.
 * wz: km_pack_weights_zfold_pair (fp32 (Cout,Cin,3,3,3) -> bf16 [rotation][dx][dy][3*Cout][Cin],
 * 81*Cout*Cin*2 bytes).  Same tensors / flags / statistics as km_conv3d_tc (KM_CONV_RELU | KM_CONV_STATS);
 * `pooled` (may be NULL) as in km_conv3d_zfold: fused MaxPool3d(2), statistics of the pooled tensor,
 * `out` may then be NULL.  Shapes: Cout 64 with Cin % 32 == 0, Cout 32 with Cin % 32 == 0 or Cin == 16. */
int km_conv3d_zfold_pair_supported(int Cin, int Cout, int D, int H, int W);
int km_pack_weights_zfold_pair(const float* w, void* packed, int Cout, int Cin, km_stream_t stream);
int km_conv3d_zfold_pair(const void* x, const void* wz, void* out, void* pooled, float* stats, int N, int Cin,
                         int Cout, int D, int H, int W, int flags, km_stream_t stream);

/* km_conv3d_zfold_pair with the preceding GroupNorm folded in, exactly as km_conv3d_zfold_gn: x is the
 * RAW bf16 activation, w the fp32 (Cout,Cin,3,3,3) weights, scale / shift (N,Cin).  The per-sample
 * weight sets are the last dimension of the weight tensor map (rotation + 3 n); the epilogue adds
 * bias[n][border class][cout] after it has handed the TMEM block back to the MMA issuer.
 * workspace: km_conv3d_zfold_pair_gn_workspace_bytes(N, Cin, Cout), 256-byte aligned. */
size_t km_conv3d_zfold_pair_gn_workspace_bytes(int N, int Cin, int Cout);
int km_conv3d_zfold_pair_gn(const void* x, const float* w, const float* scale, const float* shift, void* out,
                            void* pooled, float* stats, void* workspace, int N, int Cin, int Cout, int D,
                            int H, int W, int flags, km_stream_t stream);

/* The same for a channel-concatenated input that is never materialised (decoder: cat(skip, upsampled x),
 * keymorph/unet3d/buildingblocks.py:409-445): input channels [0, Cin0) are read from x0 (N,D,H,W,Cin0) and
 * [Cin0, Cin0 + Cin1) from x1 (N,D,H,W,Cin1) through two tensor maps; w / scale / shift cover all
 * Cin0 + Cin1 channels.  Cin0 and Cin1 must be multiples of the K chunk (64 when both are, else 32). */
int km_conv3d_zfold_pair_gn_cat(const void* x0, const void* x1, int Cin0, int Cin1, const float* w,
                                const float* scale, const float* shift, void* out, void* pooled, float* stats,
                                void* workspace, int N, int Cout, int D, int H, int W, int flags,
                                km_stream_t stream);

/* The same layer split so that the upsampled half is computed on the COARSE lattice and never materialised
 * (decoder: cat(skip, F.interpolate(x, 2, 'nearest')) -> GN -> conv, keymorph/unet3d/buildingblocks.py:409-445).
 * Nearest upsampling repeats each coarse voxel on 2x2x2 fine voxels, so per output parity class the 27 taps
 * collapse to 8 taps over the coarse tensor with pre-summed weights (8/27 of the MMA work):
 *   km_conv3d_up2_gn            xc (N,Dc,Hc,Wc,Cu) raw 16-bit; w (Cout, Cs+Cu, 3,3,3) fp32; scale (N, Cs+Cu) or NULL;
 *                               out (N,2Dc,2Hc,2Wc,Cout) 16-bit PARTIAL sums of input channels [Cs, Cs+Cu)
 *                               (no bias / activation).  Cout % 64 == 0, Cu % 64 == 0.
 *   km_conv3d_zfold_pair_gn_add the skip half (input channels [0, Cs) of x (N,D,H,W,Cs)) with `addend` = the
 *                               partial sums above added before the folded-norm bias (all Cs+Cu channels' shift
 *                               terms), ReLU and statistics.  workspace as km_conv3d_zfold_pair_gn(N, Cs, Cout). */
int km_conv3d_up2_supported(int Cu, int Cout, int Dc, int Hc, int Wc);
size_t km_conv3d_up2_gn_workspace_bytes(int N, int Cu, int Cout);
int km_conv3d_up2_gn(const void* xc, const float* w, const float* scale, int Cs, int Cu, void* out, void* workspace,
                     int N, int Cout, int Dc, int Hc, int Wc, km_stream_t stream);
int km_conv3d_zfold_pair_gn_add(const void* x, int Cs, int Cu, const float* w, const float* scale, const float* shift,
                                const void* addend, void* out, float* stats, void* workspace, int N, int Cout, int D,
                                int H, int W, int flags, km_stream_t stream);
/* The same on the plain CTA-pair kernel (Cout in {64, 128}, below 96^3 voxels; workspace as
 * km_conv3d_tc_pair_gn(N, Cs, Cout)). */
int km_conv3d_tc_pair_gn_add(const void* x, int Cs, int Cu, const float* w, const float* scale, const float* shift,
                             const void* addend, void* out, float* stats, void* workspace, int N, int Cout, int D,
                             int H, int W, int flags, km_stream_t stream);

/* Nearest-neighbour x2 upsampling of a bf16 NDHWC tensor (F.interpolate in the decoders,
 * buildingblocks.py:409-445): (N,Dc,Hc,Wc,C) -> (N,2Dc,2Hc,2Wc,C), C % 8 == 0. */
int km_upsample2_ndhwc(const void* src, void* dst, int N, int C, int Dc, int Hc, int Wc, km_stream_t stream);

/* Final 1x1x1 convolution fused with ReLU + centre of mass, transposed tcgen05 formulation
 * (keymorph/unet3d/model.py:99,389 final_conv + keymorph/layers.py:92-134 + keymorph/model.py:95-109):
 * heat^T[channel, voxel] = W . X^T, one epilogue thread per keypoint channel, sums in registers; the
 * heat map is never materialised.
 *   x: bf16 NDHWC (N,D,H,W,Cin), Cin % 16 == 0, Cin <= 256;  wp: bf16 [Cout][Cin] (km_pack_weights
 *   with taps = 1), Cout % 128 == 0 (zero-pad), Cout <= 512;  bias: fp32 (Cout) or NULL;
 *   com: (km_conv1x1_com_nparts(),N,Cout,4) fp32 partial [sum h, sum h*lz, sum h*ly, sum h*lx],
 *   h = relu(conv), l* = linspace(0,1,n) -- the same partials as KM_CONV_COM, finished by
 *   km_com_finalize. */
int km_conv1x1_com_nparts(void);
int km_conv1x1_com(const void* x, const void* wp, const float* bias, float* com, int N, int Cin,
                   int Cout, int D, int H, int W, km_stream_t stream);

/* com partials -> keypoints (N,K,3) 'ij' order + optional mass (N,K)
 * (keymorph/layers.py:121-134: c = sum(lin*m)/(M+1e-8), *2-1). */
int km_com_finalize(const float* com, int nparts, float* points, float* mass, int N, int K,
                    km_stream_t stream);

/* bf16 NDHWC (N,D,H,W,C) -> fp32 NCDHW (N,C,D,H,W) and back (feature export / tests) */
int km_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, int D, int H, int W,
                               km_stream_t stream);
int km_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, int D, int H, int W,
                               km_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KM_B200_H */
