// Dense 3-D warp (grid_sample), flow-field generation and the fused warp + loss kernels.
//
// Reference call sites:
//   keymorph/utils.py:14-21            align_img = F.grid_sample(x, grid, mode, "border", align_corners=False)
//   keymorph/utils.py:387-398          uniform_norm_grid (linspace(-1,1,S)^3, 'ij' order)
//   keymorph/transformations.py:37-114 AffineTransform.get_flow_field / transformed points
//   keymorph/keypoint_aligners.py:365-449 TPS.get_flow_field / transform_points
//   keymorph/loss_ops.py:9-63          MSELoss / DiceLoss
// All kernels are HBM-bound streaming kernels: 4 consecutive x-voxels per thread, 16-byte loads
// and stores on the contiguous streams (grid, fixed, out), gathers through the read-only path.
#include <climits>
#include "km_common.cuh"

namespace {

int g_tps_fast = 1;
}  // namespace
void km_conv_set_force_generic(int v);
void km_conv_set_no_resident(int v);
void km_conv_set_max_mt(int v);
void km_conv_set_no_epi_batch(int v);
void km_conv_set_halo_axis(int v);
void km_conv_set_interleave(int v);
void km_conv_set_two_issuers(int v);
void km_zf2_set_two_bricks(int v);
void km_tps_set_single_cta(int v);
void km_tps_set_packed(int v);
void km_tps_set_vpt(int v);
void km_set_operand_fp16(int v);
void km_warp_set_tile(int v);
bool km_warp_tile_eligible(const float* moving, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo);
int km_warp_tile_launch(int coord, bool exact, const float* mat, const float* grid, const float* moving,
                        const float* fixed, float* out, float* grid_out, float* partials, int N, int C, int Di,
                        int Hi, int Wi, int Do, int Ho, int Wo, cudaStream_t st);
int km_tps_fast_enabled() { return g_tps_fast; }
namespace {

// ---- ATen grid_sampler_3d source-index arithmetic (align_corners=False, padding "border"),
// written without FMA contraction so that it rounds like the scalar CPU reference.
__device__ __forceinline__ float src_index(float g, int size) {
  float v = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
  v = fminf((float)(size - 1), fmaxf(v, 0.f));
  return v;
}

struct Tri {
  int x0, y0, z0;
  float ix, iy, iz;
};

__device__ __forceinline__ Tri make_tri(float gx, float gy, float gz, int D, int H, int W) {
  Tri t;
  t.ix = src_index(gx, W);
  t.iy = src_index(gy, H);
  t.iz = src_index(gz, D);
  t.x0 = (int)floorf(t.ix);
  t.y0 = (int)floorf(t.iy);
  t.z0 = (int)floorf(t.iz);
  return t;
}

// trilinear gather; corners outside the volume are skipped exactly like ATen does
__device__ __forceinline__ float tri_sample(const float* __restrict__ vol, const Tri& t, int D,
                                            int H, int W) {
  const float x0f = (float)t.x0, y0f = (float)t.y0, z0f = (float)t.z0;
  const float wx1 = __fsub_rn(t.ix, x0f), wx0 = __fsub_rn(x0f + 1.f, t.ix);
  const float wy1 = __fsub_rn(t.iy, y0f), wy0 = __fsub_rn(y0f + 1.f, t.iy);
  const float wz1 = __fsub_rn(t.iz, z0f), wz0 = __fsub_rn(z0f + 1.f, t.iz);
  const bool x1ok = t.x0 + 1 < W, y1ok = t.y0 + 1 < H, z1ok = t.z0 + 1 < D;
  const size_t HW = (size_t)H * W;
  const float* p = vol + (size_t)t.z0 * HW + (size_t)t.y0 * W + t.x0;
  float acc = 0.f;
  // order: tnw, tne, tsw, tse, bnw, bne, bsw, bse (t = z0, n = y0, w = x0)
  acc = __fadd_rn(acc, __fmul_rn(__ldg(p), __fmul_rn(__fmul_rn(wx0, wy0), wz0)));
  if (x1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + 1), __fmul_rn(__fmul_rn(wx1, wy0), wz0)));
  if (y1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + W), __fmul_rn(__fmul_rn(wx0, wy1), wz0)));
  if (x1ok && y1ok)
    acc = __fadd_rn(acc, __fmul_rn(__ldg(p + W + 1), __fmul_rn(__fmul_rn(wx1, wy1), wz0)));
  if (z1ok) {
    const float* q = p + HW;
    acc = __fadd_rn(acc, __fmul_rn(__ldg(q), __fmul_rn(__fmul_rn(wx0, wy0), wz1)));
    if (x1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(q + 1), __fmul_rn(__fmul_rn(wx1, wy0), wz1)));
    if (y1ok) acc = __fadd_rn(acc, __fmul_rn(__ldg(q + W), __fmul_rn(__fmul_rn(wx0, wy1), wz1)));
    if (x1ok && y1ok)
      acc = __fadd_rn(acc, __fmul_rn(__ldg(q + W + 1), __fmul_rn(__fmul_rn(wx1, wy1), wz1)));
  }
  return acc;
}

// Same value up to FMA rounding (<= 1 ulp of the result) with half the instructions: corner indices
// are clamped instead of predicated (a clamped corner always carries weight exactly 0), the x-y
// weight products are shared between the two z planes, 32-bit offsets, FMA accumulation.  Used by
// the fused warp+loss kernels; km_grid_sample3d keeps the bit-exact formulation above.
__device__ __forceinline__ float tri_sample_fast(const float* __restrict__ vol, const Tri& t, int D,
                                                 int H, int W) {
  const float x0f = (float)t.x0, y0f = (float)t.y0, z0f = (float)t.z0;
  const float wx1 = t.ix - x0f, wx0 = (x0f + 1.f) - t.ix;
  const float wy1 = t.iy - y0f, wy0 = (y0f + 1.f) - t.iy;
  const float wz1 = t.iz - z0f, wz0 = (z0f + 1.f) - t.iz;
  const int x1 = min(t.x0 + 1, W - 1), y1 = min(t.y0 + 1, H - 1), z1 = min(t.z0 + 1, D - 1);
  const int r00 = (t.z0 * H + t.y0) * W, r01 = (t.z0 * H + y1) * W;
  const int r10 = (z1 * H + t.y0) * W, r11 = (z1 * H + y1) * W;
  const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
  float acc = __ldg(vol + r00 + t.x0) * (w00 * wz0);
  acc = fmaf(__ldg(vol + r00 + x1), w10 * wz0, acc);
  acc = fmaf(__ldg(vol + r01 + t.x0), w01 * wz0, acc);
  acc = fmaf(__ldg(vol + r01 + x1), w11 * wz0, acc);
  acc = fmaf(__ldg(vol + r10 + t.x0), w00 * wz1, acc);
  acc = fmaf(__ldg(vol + r10 + x1), w10 * wz1, acc);
  acc = fmaf(__ldg(vol + r11 + t.x0), w01 * wz1, acc);
  acc = fmaf(__ldg(vol + r11 + x1), w11 * wz1, acc);
  return acc;
}

__device__ __forceinline__ float nearest_sample(const float* __restrict__ vol, const Tri& t, int H,
                                                int W) {
  // std::nearbyint: round half to even
  const int x = (int)rintf(t.ix), y = (int)rintf(t.iy), z = (int)rintf(t.iz);
  return __ldg(vol + ((size_t)z * H + y) * W + x);
}

// ---- coordinate generators: (z,y,x) voxel -> grid_sample (gx,gy,gz) ----------------------
struct AffineCoord {
  float m[12];  // rows in (z,y,x) order
  __device__ __forceinline__ void operator()(float pz, float py, float px, float& gx, float& gy,
                                             float& gz) const {
    gz = fmaf(m[0], pz, fmaf(m[1], py, fmaf(m[2], px, m[3])));
    gy = fmaf(m[4], pz, fmaf(m[5], py, fmaf(m[6], px, m[7])));
    gx = fmaf(m[8], pz, fmaf(m[9], py, fmaf(m[10], px, m[11])));
  }
};

// TPS radial basis, keymorph/keypoint_aligners.py:322-339:
//   r = sqrt(|a-b|^2 + 1e-6);  U = r^2 * log(r + 1e-6)
template <bool FAST>
__device__ __forceinline__ float tps_u(float d2) {
  const float s = d2 + 1e-6f;
  if (FAST) {
    // r^2 log(r + e) = s * (0.5 log s + log1p(e/r)) ~= 0.5 ln2 * s * lg2(s) + e * r      (e/r <= 1e-3)
    // The second term is at most 3.5e-6, so r = sqrt(s) only needs ~1e-2 relative accuracy: the
    // exponent-halving bit trick (max error 3.5 %) replaces the second MUFU op of every term; the
    // resulting absolute error (< 1.3e-7 per unit weight) is a fifth of the fp32 rounding of the
    // first term.
    const float r = __int_as_float((__float_as_int(s) >> 1) + 0x1fbd1df5);
    return fmaf(0.34657359028f * s, __log2f(s), 1e-6f * r);
  } else {
    const float r = sqrtf(s);
    return (r * r) * logf(r + 1e-6f);
  }
}

// smem layout for TPS: c4[t] = (cz, cy, cx, 0), w4[t] = (wz, wy, wx, 0); aff[12] = rows 1,z,y,x
template <bool FAST>
__device__ __forceinline__ void tps_eval(const float4* __restrict__ c4, const float4* __restrict__ w4,
                                         const float* __restrict__ aff, int K, float pz, float py,
                                         float px, float& oz, float& oy, float& ox) {
  float az = 0.f, ay = 0.f, ax = 0.f;
#pragma unroll 4
  for (int t = 0; t < K; ++t) {
    const float4 c = c4[t];
    const float4 w = w4[t];
    const float dz = pz - c.x, dy = py - c.y, dx = px - c.z;
    const float u = tps_u<FAST>(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    az = fmaf(u, w.x, az);
    ay = fmaf(u, w.y, ay);
    ax = fmaf(u, w.z, ax);
  }
  // z = [1, p] . affine  (keymorph/keypoint_aligners.py:427-433), out = z + b
  oz = (aff[0] + aff[3] * pz + aff[6] * py + aff[9] * px) + az;
  oy = (aff[1] + aff[4] * pz + aff[7] * py + aff[10] * px) + ay;
  ox = (aff[2] + aff[5] * pz + aff[8] * py + aff[11] * px) + ax;
}

__device__ __forceinline__ void load_tps_smem(const float* __restrict__ ctrl,
                                              const float* __restrict__ theta, int K, float4* c4,
                                              float4* w4, float* aff) {
  for (int t = threadIdx.x; t < K; t += blockDim.x) {
    c4[t] = make_float4(ctrl[t * 3 + 0], ctrl[t * 3 + 1], ctrl[t * 3 + 2], 0.f);
    w4[t] = make_float4(theta[t * 3 + 0], theta[t * 3 + 1], theta[t * 3 + 2], 0.f);
  }
  for (int i = threadIdx.x; i < 12; i += blockDim.x) aff[i] = theta[K * 3 + i];
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// flow fields
__global__ void __launch_bounds__(256)
flow_affine_kernel(const float* __restrict__ mat, float* __restrict__ grid, int D, int H, int W) {
  const int n = blockIdx.y;
  AffineCoord ac;
#pragma unroll
  for (int i = 0; i < 12; ++i) ac.m[i] = __ldg(mat + n * 12 + i);
  const long long nvox = (long long)D * H * W;
  float* gn = grid + (size_t)n * nvox * 3;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox;
       v += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(v % W), y = (int)((v / W) % H), z = (int)(v / ((long long)W * H));
    float gx, gy, gz;
    ac(km_linspace(-1.f, 1.f, D, z), km_linspace(-1.f, 1.f, H, y), km_linspace(-1.f, 1.f, W, x), gx,
       gy, gz);
    gn[v * 3 + 0] = gx;
    gn[v * 3 + 1] = gy;
    gn[v * 3 + 2] = gz;
  }
}

__global__ void points_affine_kernel(const float* __restrict__ mat, const float* __restrict__ pts,
                                     float* __restrict__ out, int P) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* m = mat + n * 12;
  const float* q = pts + ((size_t)n * P + p) * 3;
  float* o = out + ((size_t)n * P + p) * 3;
  const float a = q[0], b = q[1], c = q[2];
  o[0] = fmaf(m[0], a, fmaf(m[1], b, fmaf(m[2], c, m[3])));
  o[1] = fmaf(m[4], a, fmaf(m[5], b, fmaf(m[6], c, m[7])));
  o[2] = fmaf(m[8], a, fmaf(m[9], b, fmaf(m[10], c, m[11])));
}

__global__ void __launch_bounds__(128)
points_tps_kernel(const float* __restrict__ ctrl, const float* __restrict__ theta,
                  const float* __restrict__ pts, float* __restrict__ out, int K, int P) {
  extern __shared__ float4 s4[];
  float4* c4 = s4;
  float4* w4 = s4 + K;
  float* aff = reinterpret_cast<float*>(s4 + 2 * K);
  const int n = blockIdx.y;
  load_tps_smem(ctrl + (size_t)n * K * 3, theta + (size_t)n * (K + 4) * 3, K, c4, w4, aff);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* q = pts + ((size_t)n * P + p) * 3;
  float oz, oy, ox;
  tps_eval<false>(c4, w4, aff, K, q[0], q[1], q[2], oz, oy, ox);
  float* o = out + ((size_t)n * P + p) * 3;
  o[0] = oz;
  o[1] = oy;
  o[2] = ox;
}

// ------------------------------------------------------------------------------------------
// stand-alone grid_sample.  A warp owns 128 consecutive output voxels; lane i handles voxels
// i, i+32, i+64, i+96 of the chunk, so every load / store instruction of the warp touches 32
// CONSECUTIVE voxels (the gathers of a warp then fall into one or two cache lines per corner for
// the near-identity transforms of registration) while each thread still has 4 independent voxels
// in flight.
__global__ void __launch_bounds__(256)
grid_sample_kernel(const float* __restrict__ x, const float* __restrict__ grid,
                   float* __restrict__ out, int C, int Di, int Hi, int Wi, long long nvo,
                   int mode) {
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long nchunks = (nvo + 127) / 128;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const size_t in_vol = (size_t)Di * Hi * Wi;
  const float* gn = grid + (size_t)n * nvo * 3;
  for (long long chunk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks;
       chunk += warps) {
    const long long v0 = chunk * 128 + lane;
    Tri t[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long v = v0 + 32 * k;
      ok[k] = v < nvo;
      const float* gp = gn + (ok[k] ? v : 0) * 3;
      t[k] = make_tri(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), Di, Hi, Wi);
    }
    for (int c = 0; c < C; ++c) {
      const float* vol = x + ((size_t)n * C + c) * in_vol;
      float* o = out + ((size_t)n * C + c) * nvo + v0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ok[k])
          o[32 * k] = (mode == KM_INTERP_NEAREST) ? nearest_sample(vol, t[k], Hi, Wi)
                                                  : tri_sample(vol, t[k], Di, Hi, Wi);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused warp (+ store) (+ flow-field store) (+ loss partials).  grid (KM_RED_BLOCKS, N), 256 threads.
// partials: [gridDim.x][N][C][4] floats = sum (a-f)^2, sum a*f, sum a*a, sum f*f
//
// The (x,y,z) triples of a warp's 128 consecutive voxels are 1536 contiguous bytes of the flow
// field: they are moved with three 16-byte accesses per lane and transposed through a per-warp
// shared-memory buffer (stride-3 word access is bank-conflict free), instead of 12-byte-strided
// scalar accesses whose L1 wavefronts bound the kernel.
// (Occupancy was A/B-tested: compiled for 5 / 6 resident CTAs per SM (48 / 40 registers) the affine kernel takes
// 211 / 222 us instead of 208 us at 256^3 -- the gather is bound by L1 wavefronts, not by latency hiding.)
template <int COORD, int CCH, bool FAST>
__global__ void __launch_bounds__(256)
warp_loss_kernel(const float* __restrict__ mat_or_ctrl, const float* __restrict__ theta, int K,
                 const float* __restrict__ grid, const float* __restrict__ moving,
                 const float* __restrict__ fixed, float* __restrict__ out,
                 float* __restrict__ grid_out, float* __restrict__ partials, int N, int C, int D,
                 int H, int W, int mode) {
  extern __shared__ float4 s4[];
  __shared__ float red[8][CCH * 4];
  __shared__ __align__(16) float s_g[8][384];   // per-warp staging of 128 flow-field triples
  const int n = blockIdx.y;
  const long long nvox = (long long)D * H * W;
  AffineCoord ac;
  float4* c4 = s4;
  float4* w4 = s4 + K;
  float* aff = reinterpret_cast<float*>(s4 + 2 * K);
  if (COORD == KM_COORD_AFFINE) {
#pragma unroll
    for (int i = 0; i < 12; ++i) ac.m[i] = __ldg(mat_or_ctrl + n * 12 + i);
  } else if (COORD == KM_COORD_TPS) {
    load_tps_smem(mat_or_ctrl + (size_t)n * K * 3, theta + (size_t)n * (K + 4) * 3, K, c4, w4, aff);
  }
  const float* gn = (COORD == KM_COORD_GRID) ? grid + (size_t)n * nvox * 3 : nullptr;
  float* gon = grid_out ? grid_out + (size_t)n * nvox * 3 : nullptr;
  float* sg = s_g[threadIdx.x >> 5];
  const bool g_al = (((uintptr_t)(COORD == KM_COORD_GRID ? (const void*)gn : (const void*)gon)) & 15) == 0;
  // a warp owns 128 consecutive voxels, lane i handles voxels i, i+32, i+64, i+96 (see
  // grid_sample_kernel): coalesced streams AND gathers that share cache lines across the warp
  const int lane = threadIdx.x & 31;
  const int nv = (int)nvox;   // checked on the host
  const int nchunks = (nv + 127) / 128;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int HW = H * W;

  for (int cbase = 0; cbase < C; cbase += CCH) {
    const int cn = min(CCH, C - cbase);
    // single channel: loss sums in registers for the whole kernel.  Multi-channel volumes (one-hot
    // segmentations): every channel's sums of a chunk are warp-reduced and added to the warp's
    // shared-memory slots by lane 0 (program order: deterministic), so the channel loop stays
    // rolled and the register count (= occupancy) matches the single-channel instantiation
    float acc1[4] = {0.f, 0.f, 0.f, 0.f};
    if (CCH > 1) {
      __syncthreads();
      for (int i = lane; i < CCH * 4; i += 32) red[threadIdx.x >> 5][i] = 0.f;
      __syncwarp();
    }

    for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
      const int v0 = chunk * 128 + lane;
      // (z, y, x) of the chunk's first voxel: one division pair per 128 voxels
      const int cz = (chunk * 128) / HW, crem = chunk * 128 - cz * HW;
      const int cy = crem / W, cx = crem - cy * W;
      Tri t[4];
      bool ok[4];
      const bool vec = g_al && chunk * 128 + 128 <= nv;   // whole chunk inside, 16-byte aligned
      if (COORD == KM_COORD_GRID && vec && cbase == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(gn + (size_t)chunk * 384);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; ++i) reinterpret_cast<float4*>(sg)[lane + 32 * i] = __ldg(g4 + lane + 32 * i);
        __syncwarp();
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int v = v0 + 32 * k;
        ok[k] = v < nv;
        const int vv = ok[k] ? v : 0;
        float gx, gy, gz;
        if (COORD == KM_COORD_GRID) {
          if (vec && cbase == 0) {
            const float* gp = sg + (lane + 32 * k) * 3;
            gx = gp[0];
            gy = gp[1];
            gz = gp[2];
          } else {
            const float* gp = gn + (size_t)vv * 3;
            gx = __ldg(gp);
            gy = __ldg(gp + 1);
            gz = __ldg(gp + 2);
          }
        } else {
          int xx = cx + lane + 32 * k, y = cy, z = cz;
          while (xx >= W) {   // at most once per row the chunk spans
            xx -= W;
            if (++y == H) {
              y = 0;
              ++z;
            }
          }
          if (!ok[k]) xx = y = z = 0;
          const float pz = km_linspace(-1.f, 1.f, D, z), py = km_linspace(-1.f, 1.f, H, y);
          const float px = km_linspace(-1.f, 1.f, W, xx);
          if (COORD == KM_COORD_AFFINE) {
            ac(pz, py, px, gx, gy, gz);
          } else {
            tps_eval<FAST>(c4, w4, aff, K, pz, py, px, gz, gy, gx);
          }
          if (gon && cbase == 0) {
            if (vec) {
              float* gp = sg + (lane + 32 * k) * 3;
              gp[0] = gx;
              gp[1] = gy;
              gp[2] = gz;
            } else if (ok[k]) {
              float* gp = gon + (size_t)v * 3;
              gp[0] = gx;
              gp[1] = gy;
              gp[2] = gz;
            }
          }
        }
        t[k] = make_tri(gx, gy, gz, D, H, W);
      }
      if (COORD != KM_COORD_GRID && gon && vec && cbase == 0) {
        float4* g4 = reinterpret_cast<float4*>(gon + (size_t)chunk * 384);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; ++i) g4[lane + 32 * i] = reinterpret_cast<const float4*>(sg)[lane + 32 * i];
        __syncwarp();
      }
#pragma unroll 1
      for (int c = 0; c < cn; ++c) {
        const size_t ch = (size_t)n * C + cbase + c;
        const float* vol = moving + ch * nvox;
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = CCH > 1 ? 0.f : acc1[j];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ok[k]) {
            const float r = (mode == KM_INTERP_NEAREST) ? nearest_sample(vol, t[k], H, W)
                                                        : tri_sample_fast(vol, t[k], D, H, W);
            const size_t idx = ch * nvox + v0 + 32 * k;
            if (out) out[idx] = r;
            if (fixed) {
              const float fv = __ldg(fixed + idx);
              const float d = r - fv;
              acc[0] = fmaf(d, d, acc[0]);
              acc[1] = fmaf(r, fv, acc[1]);
              acc[2] = fmaf(r, r, acc[2]);
              acc[3] = fmaf(fv, fv, acc[3]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (CCH > 1) {
            if (fixed) {
              const float v = km_warp_sum(acc[j]);
              if (lane == 0) red[threadIdx.x >> 5][c * 4 + j] += v;
            }
          } else {
            acc1[j] = acc[j];
          }
        }
      }
    }
    if (fixed) {
      if (CCH == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = km_warp_sum(acc1[j]);
          if (lane == 0) red[threadIdx.x >> 5][j] = v;
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < cn * 4; i += blockDim.x) {
        float a = 0.f;
        for (int wv = 0; wv < 8; ++wv) a += red[wv][i];
        partials[(((size_t)blockIdx.x * N + n) * C + cbase) * 4 + i] = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Label-map fast path for segmentation warp + Dice.  The reference warps a C-channel
// fp32 one-hot volume (scripts/pairwise_register_eval.py:99-108,156-157: one_hot -> align_img ->
// DiceLoss soft and hard) -- 4*C bytes per voxel read, written and read again.  Here the integer
// label maps are gathered directly: the trilinear value of channel c at a voxel is the sum of the
// corner weights whose label equals c, accumulated in exactly the order tri_sample_fast would add
// them (zero terms are exact no-ops), so the soft-Dice sums equal those of the one-hot path bit
// for bit per voxel, and the hard-Dice sums are exact integer counts (argmax over c, first maximum
// wins like torch.argmax, loss_ops.py:46-49).  1 + 8 label bytes per voxel instead of 12*C.
//   soft partials [gridDim.x][N][C][4] float = sum (a-t)^2, sum a*t, sum a*a, sum t*t
//   hard counts   [N][C][3] unsigned long long = |hard==c & t==c|, |hard==c|, |t==c| (atomics on ints)
template <int COORD>
__global__ void __launch_bounds__(256)
warp_labels_kernel(const float* __restrict__ mat, const float* __restrict__ grid,
                   const uint8_t* __restrict__ lab_m, const uint8_t* __restrict__ lab_f,
                   uint8_t* __restrict__ lab_out, float* __restrict__ partials,
                   unsigned long long* __restrict__ hard, int N, int C, int D, int H, int W) {
  extern __shared__ float s_red[];            // [8 warps][C*4] soft sums, then [C*3] uint counts
  const int n = blockIdx.y;
  const int nv = D * H * W, HW = H * W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* red = s_red + (size_t)wid * C * 4;
  unsigned int* cnt = reinterpret_cast<unsigned int*>(s_red + (size_t)8 * C * 4);
  for (int i = threadIdx.x; i < 8 * C * 4; i += blockDim.x) s_red[i] = 0.f;
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) cnt[i] = 0u;
  __syncthreads();
  AffineCoord ac;
  if (COORD == KM_COORD_AFFINE) {
#pragma unroll
    for (int i = 0; i < 12; ++i) ac.m[i] = __ldg(mat + n * 12 + i);
  }
  const float* gn = (COORD == KM_COORD_GRID) ? grid + (size_t)n * nv * 3 : nullptr;
  const uint8_t* lm = lab_m + (size_t)n * nv;
  const uint8_t* lf = lab_f + (size_t)n * nv;
  const int nchunks = (nv + 127) / 128;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
    const int v0 = chunk * 128 + lane;
    const int cz = (chunk * 128) / HW, crem = chunk * 128 - cz * HW;
    const int cy = crem / W, cx = crem - cy * W;
    float cw[4][8];        // corner weights in tri_sample_fast's accumulation order
    uint32_t cl[4][2];     // corner labels, 4 per word
    int tl[4];             // fixed label, -1 for voxels past the end
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = v0 + 32 * k;
      const bool ok = v < nv;
      float gx, gy, gz;
      if (COORD == KM_COORD_GRID) {
        const float* gp = gn + (size_t)(ok ? v : 0) * 3;
        gx = __ldg(gp);
        gy = __ldg(gp + 1);
        gz = __ldg(gp + 2);
      } else {
        int xx = cx + lane + 32 * k, y = cy, z = cz;
        while (xx >= W) {
          xx -= W;
          if (++y == H) {
            y = 0;
            ++z;
          }
        }
        if (!ok) xx = y = z = 0;
        ac(km_linspace(-1.f, 1.f, D, z), km_linspace(-1.f, 1.f, H, y), km_linspace(-1.f, 1.f, W, xx), gx, gy,
           gz);
      }
      const Tri t = make_tri(gx, gy, gz, D, H, W);
      const float x0f = (float)t.x0, y0f = (float)t.y0, z0f = (float)t.z0;
      const float wx1 = t.ix - x0f, wx0 = (x0f + 1.f) - t.ix;
      const float wy1 = t.iy - y0f, wy0 = (y0f + 1.f) - t.iy;
      const float wz1 = t.iz - z0f, wz0 = (z0f + 1.f) - t.iz;
      const int x1 = min(t.x0 + 1, W - 1), y1 = min(t.y0 + 1, H - 1), z1 = min(t.z0 + 1, D - 1);
      const int r00 = (t.z0 * H + t.y0) * W, r01 = (t.z0 * H + y1) * W;
      const int r10 = (z1 * H + t.y0) * W, r11 = (z1 * H + y1) * W;
      const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
      cw[k][0] = w00 * wz0; cw[k][1] = w10 * wz0; cw[k][2] = w01 * wz0; cw[k][3] = w11 * wz0;
      cw[k][4] = w00 * wz1; cw[k][5] = w10 * wz1; cw[k][6] = w01 * wz1; cw[k][7] = w11 * wz1;
      cl[k][0] = (uint32_t)__ldg(lm + r00 + t.x0) | ((uint32_t)__ldg(lm + r00 + x1) << 8) |
                 ((uint32_t)__ldg(lm + r01 + t.x0) << 16) | ((uint32_t)__ldg(lm + r01 + x1) << 24);
      cl[k][1] = (uint32_t)__ldg(lm + r10 + t.x0) | ((uint32_t)__ldg(lm + r10 + x1) << 8) |
                 ((uint32_t)__ldg(lm + r11 + t.x0) << 16) | ((uint32_t)__ldg(lm + r11 + x1) << 24);
      tl[k] = ok ? (int)__ldg(lf + v) : -1;
    }
    float best[4] = {-1.f, -1.f, -1.f, -1.f};
    int bi[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // a = sum of the corner weights whose label is c, same order / FMA chain as the one-hot gather
        float a = (((cl[k][0]) & 0xffu) == (uint32_t)c) ? cw[k][0] : 0.f;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
          const uint32_t l = (cl[k][j >> 2] >> (8 * (j & 3))) & 0xffu;
          if (l == (uint32_t)c) a = a + cw[k][j];     // fmaf(1, w, a)
        }
        if (tl[k] >= 0) {
          const float tv = tl[k] == c ? 1.f : 0.f;
          const float d = a - tv;
          a0 = fmaf(d, d, a0);
          a1 = fmaf(a, tv, a1);
          a2 = fmaf(a, a, a2);
          a3 = fmaf(tv, tv, a3);
          if (a > best[k]) {      // first maximum wins
            best[k] = a;
            bi[k] = c;
          }
        }
      }
      a0 = km_warp_sum(a0);
      a1 = km_warp_sum(a1);
      a2 = km_warp_sum(a2);
      a3 = km_warp_sum(a3);
      if (lane == 0) {
        red[c * 4 + 0] += a0;
        red[c * 4 + 1] += a1;
        red[c * 4 + 2] += a2;
        red[c * 4 + 3] += a3;
      }
    }
    // hard Dice: exact counts via warp ballots
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      unsigned int tp = 0, pc = 0, tc = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = tl[k] >= 0;
        const unsigned bp = __ballot_sync(0xffffffffu, ok && bi[k] == c);
        const unsigned bt = __ballot_sync(0xffffffffu, ok && tl[k] == c);
        tp += __popc(bp & bt);
        pc += __popc(bp);
        tc += __popc(bt);
      }
      if (lane == 0 && (pc | tc)) {
        atomicAdd(&cnt[c * 3 + 0], tp);
        atomicAdd(&cnt[c * 3 + 1], pc);
        atomicAdd(&cnt[c * 3 + 2], tc);
      }
    }
    if (lab_out) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (tl[k] >= 0) lab_out[(size_t)n * nv + v0 + 32 * k] = (uint8_t)bi[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 4; i += blockDim.x) {
    float a = 0.f;
    for (int wv = 0; wv < 8; ++wv) a += s_red[(size_t)wv * C * 4 + i];
    partials[((size_t)blockIdx.x * N + n) * C * 4 + i] = a;
  }
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x)
    if (cnt[i]) atomicAdd(&hard[(size_t)n * C * 3 + i], (unsigned long long)cnt[i]);
}

// hard counts -> the same [sum (p-t)^2, sum p*t, sum p*p, sum t*t] layout as the soft sums
__global__ void hard_counts_to_sums_kernel(const unsigned long long* __restrict__ hard, double* __restrict__ sums,
                                           int NC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const double tp = (double)hard[i * 3], p = (double)hard[i * 3 + 1], t = (double)hard[i * 3 + 2];
  sums[i * 4 + 0] = p + t - 2.0 * tp;
  sums[i * 4 + 1] = tp;
  sums[i * 4 + 2] = p;
  sums[i * 4 + 3] = t;
}

// ------------------------------------------------------------------------------------------
// Jacobian-determinant statistics of a 3-component field (keymorph/loss_ops.py:161-247,
// _jacobian_determinant / jdstd / jdlessthan0): central differences with weights
// (-0.5, 0, 0.5) along z, y, x of every component (rounded to fp32 like scipy.ndimage.correlate on
// a float32 array), J = grad + I in fp64, determinant by the reference's cofactor expansion, over
// the interior cropped by 2 voxels on every side.  The field is addressed through element strides,
// so both the (N,3,D,H,W) tensor the reference passes and the (N,D,H,W,3) grid it was permuted
// from are read in place.  partials: [gridDim.x][N][3] double = sum det, sum det^2, count(det <= 0)
__global__ void __launch_bounds__(256)
jacobian_stats_kernel(const float* __restrict__ f, long long sn, long long sc, long long sz, long long sy,
                      long long sx, double* __restrict__ partials, int N, int D, int H, int W) {
  const int n = blockIdx.y;
  const int Dz = D - 4, Hy = H - 4, Wx = W - 4;
  const long long nint = (long long)Dz * Hy * Wx;
  const float* fn = f + n * sn;
  double s = 0.0, ss = 0.0, neg = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nint;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wx) + 2, y = (int)((i / Wx) % Hy) + 2, z = (int)(i / ((long long)Wx * Hy)) + 2;
    const float* p = fn + z * sz + y * sy + x * sx;
    double J[3][3];   // J[a][b] = d(component b) / d(direction a) + delta_ab, a in (z, y, x)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const float* q = p + b * sc;
      J[0][b] = (double)(float)(0.5 * (double)__ldg(q + sz) - 0.5 * (double)__ldg(q - sz));
      J[1][b] = (double)(float)(0.5 * (double)__ldg(q + sy) - 0.5 * (double)__ldg(q - sy));
      J[2][b] = (double)(float)(0.5 * (double)__ldg(q + sx) - 0.5 * (double)__ldg(q - sx));
    }
    J[0][0] += 1.0;
    J[1][1] += 1.0;
    J[2][2] += 1.0;
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
                       J[1][0] * (J[0][1] * J[2][2] - J[0][2] * J[2][1]) +
                       J[2][0] * (J[0][1] * J[1][2] - J[0][2] * J[1][1]);
    s += det;
    ss += det * det;
    neg += det <= 0.0 ? 1.0 : 0.0;
  }
  __shared__ double red[8][3];
  s = km_warp_sum(s);
  ss = km_warp_sum(ss);
  neg = km_warp_sum(neg);
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = s;
    red[threadIdx.x >> 5][1] = ss;
    red[threadIdx.x >> 5][2] = neg;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
    partials[((size_t)blockIdx.x * N + n) * 3 + threadIdx.x] = a;
  }
}

// partials -> out[n] = (std over the interior (ddof = 0), count(det <= 0), mean, number of voxels)
__global__ void jacobian_finalize_kernel(const double* __restrict__ partials, int nparts, double count,
                                         double* __restrict__ out, int N) {
  const int n = blockIdx.x;
  const int lane = threadIdx.x;
  double a[3] = {0.0, 0.0, 0.0};
  for (int p = lane; p < nparts; p += 32)
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] += partials[((size_t)p * N + n) * 3 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) a[k] = km_warp_sum(a[k]);
  if (lane == 0) {
    const double mean = a[0] / count;
    double var = a[1] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    out[n * 4 + 0] = sqrt(var);
    out[n * 4 + 1] = a[2];
    out[n * 4 + 2] = mean;
    out[n * 4 + 3] = count;
  }
}

// partials [nparts][NC][4] -> sums[NC][4] (fp64)
__global__ void sum_partials4_kernel(const float* __restrict__ partials, int nparts, int NC4,
                                     double* __restrict__ sums) {
  // one warp per output value: lanes stride over the partial slots, fixed-order shuffle reduce
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= NC4) return;
  double s = 0.0;
  for (int p = lane; p < nparts; p += 32) s += (double)partials[(size_t)p * NC4 + i];
  s = km_warp_sum(s);
  if (lane == 0) sums[i] = s;
}

// ------------------------------------------------------------------------------------------
// pair statistics (MSE / Dice), grid (KM_RED_BLOCKS, C, N): one channel per blockIdx.y
__global__ void __launch_bounds__(KM_RED_THREADS)
pair_stats_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                  const int32_t* __restrict__ labels, float* __restrict__ partials, int N, int C,
                  long long M) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float* p = pred + ((size_t)n * C + c) * M;
  const float* t = target + ((size_t)n * C + c) * M;
  const int32_t* lab = labels ? labels + (size_t)n * M : nullptr;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    const float tv = __ldg(t + i);
    const float pv = lab ? ((__ldg(lab + i) == c) ? 1.f : 0.f) : __ldg(p + i);
    const float d = pv - tv;
    a0 = fmaf(d, d, a0);
    a1 = fmaf(pv, tv, a1);
    a2 = fmaf(pv, pv, a2);
    a3 = fmaf(tv, tv, a3);
  }
  __shared__ float red[KM_RED_THREADS / 32][4];
  a0 = km_warp_sum(a0);
  a1 = km_warp_sum(a1);
  a2 = km_warp_sum(a2);
  a3 = km_warp_sum(a3);
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = a0;
    red[threadIdx.x >> 5][1] = a1;
    red[threadIdx.x >> 5][2] = a2;
    red[threadIdx.x >> 5][3] = a3;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float a = 0.f;
    for (int wv = 0; wv < KM_RED_THREADS / 32; ++wv) a += red[wv][threadIdx.x];
    partials[(((size_t)blockIdx.x * N + n) * C + c) * 4 + threadIdx.x] = a;
  }
}

// torch.argmax over the channel dim: first maximal value wins, NaN counts as maximal
__global__ void __launch_bounds__(256)
argmax_channels_kernel(const float* __restrict__ pred, int32_t* __restrict__ labels, int C,
                       long long M) {
  const int n = blockIdx.y;
  const float* p = pred + (size_t)n * C * M;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    float best = __ldg(p + i);
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __ldg(p + (size_t)c * M + i);
      if (!(best != best) && (v > best || v != v)) {
        best = v;
        bi = c;
      }
    }
    labels[(size_t)n * M + i] = bi;
  }
}

inline int blocks_for(long long items, int threads, int cap) {
  long long b = (items + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int km_operand_is_fp16(void) { return km_operand_fp16(); }

extern "C" int km_set_option(int key, int value) {
  if (key == KM_OPT_TPS_FAST) {
    g_tps_fast = value ? 1 : 0;
    return KM_OK;
  }
  if (key == KM_OPT_CONV_FORCE_GENERIC) {
    km_conv_set_force_generic(value);
    return KM_OK;
  }
  if (key == KM_OPT_CONV_NO_RESIDENT_WEIGHTS) {
    km_conv_set_no_resident(value);
    return KM_OK;
  }
  if (key == KM_OPT_CONV_MAX_BRICKS) {
    km_conv_set_max_mt(value);
    return KM_OK;
  }
  if (key == KM_OPT_CONV_NO_EPILOGUE_BATCH) {
    km_conv_set_no_epi_batch(value);
    return KM_OK;
  }
  if (key == KM_OPT_CONV_HALO_AXIS) {
    km_conv_set_halo_axis(value);
    return KM_OK;
  }
  if (key == KM_OPT_CONV_INTERLEAVE_BRICKS) {
    km_conv_set_interleave(value);
    return KM_OK;
  }
  if (key == KM_OPT_CONV_TWO_ISSUERS) {
    km_conv_set_two_issuers(value);
    return KM_OK;
  }
  if (key == KM_OPT_ZF2_TWO_BRICKS) {
    km_zf2_set_two_bricks(value);
    return KM_OK;
  }
  if (key == KM_OPT_TPS_SINGLE_CTA) {
    km_tps_set_single_cta(value);
    return KM_OK;
  }
  if (key == KM_OPT_TPS_PACKED) {
    km_tps_set_packed(value);
    return KM_OK;
  }
  if (key == KM_OPT_TPS_VPT) {
    km_tps_set_vpt(value);
    return KM_OK;
  }
  if (key == KM_OPT_OPERAND_FP16) {
    km_set_operand_fp16(value);
    return KM_OK;
  }
  if (key == KM_OPT_WARP_TILE) {
    km_warp_set_tile(value);
    return KM_OK;
  }

  km_set_error("km_set_option: unknown key %d", key);
  return KM_EINVAL;
}

extern "C" int km_grid_sample3d(const float* x, const float* grid, float* out, int N, int C, int Di,
                                int Hi, int Wi, int Do, int Ho, int Wo, int mode,
                                km_stream_t stream) {
  KM_CHECK_ARG(x && grid && out, "km_grid_sample3d: null pointer");
  KM_CHECK_ARG(N > 0 && C > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0,
               "km_grid_sample3d: bad shape");
  KM_CHECK_ARG(mode == KM_INTERP_BILINEAR || mode == KM_INTERP_NEAREST, "km_grid_sample3d: bad mode");
  const long long nvo = (long long)Do * Ho * Wo;
  if (mode == KM_INTERP_BILINEAR && km_warp_tile_eligible(x, C, Di, Hi, Wi, Do, Ho, Wo))
    // TMA-staged tiles, ATen's un-fused arithmetic: same bits as grid_sample_kernel (C = 14: 5.3 -> 2.0 ms;
    // the fused warp + loss kernels below stay on the direct gather: measured faster for one channel, and
    // their per-CTA summation order is what makes the label-map Dice path bit-identical to them)
    return km_warp_tile_launch(KM_COORD_GRID, true, nullptr, grid, x, nullptr, out, nullptr, nullptr, N, C, Di, Hi, Wi,
                               Do, Ho, Wo, km_cs(stream));
  const dim3 g(blocks_for((nvo + 3) / 4, 256, 148 * 8), N);
  grid_sample_kernel<<<g, 256, 0, km_cs(stream)>>>(x, grid, out, C, Di, Hi, Wi, nvo, mode);
  KM_LAUNCH_OK("grid_sample_kernel");
  return KM_OK;
}

extern "C" int km_flow_field_affine(const float* mat, float* grid, int N, int D, int H, int W,
                                    km_stream_t stream) {
  KM_CHECK_ARG(mat && grid && N > 0 && D > 0 && H > 0 && W > 0, "km_flow_field_affine: bad arguments");
  const dim3 g(blocks_for((long long)D * H * W, 256, 148 * 8), N);
  flow_affine_kernel<<<g, 256, 0, km_cs(stream)>>>(mat, grid, D, H, W);
  KM_LAUNCH_OK("flow_affine_kernel");
  return KM_OK;
}

extern "C" int km_points_transform_affine(const float* mat, const float* pts, float* out, int N,
                                          int P, km_stream_t stream) {
  KM_CHECK_ARG(mat && pts && out && N > 0 && P > 0, "km_points_transform_affine: bad arguments");
  points_affine_kernel<<<dim3((P + 127) / 128, N), 128, 0, km_cs(stream)>>>(mat, pts, out, P);
  KM_LAUNCH_OK("points_affine_kernel");
  return KM_OK;
}

extern "C" int km_points_transform_tps(const float* ctrl, const float* theta, const float* pts,
                                       float* out, int N, int K, int P, km_stream_t stream) {
  KM_CHECK_ARG(ctrl && theta && pts && out && N > 0 && K > 0 && P > 0,
               "km_points_transform_tps: bad arguments");
  const size_t smem = (size_t)(2 * K + 3) * sizeof(float4);
  KM_CHECK_ARG(smem <= 48 * 1024, "km_points_transform_tps: K=%d too large", K);
  points_tps_kernel<<<dim3((P + 127) / 128, N), 128, smem, km_cs(stream)>>>(ctrl, theta, pts, out, K,
                                                                           P);
  KM_LAUNCH_OK("points_tps_kernel");
  return KM_OK;
}

extern "C" size_t km_warp_loss_workspace_bytes(int N, int C) {
  return (size_t)KM_RED_BLOCKS * N * C * 4 * sizeof(float);
}
static size_t pair_partials_bytes(int N, int C) {
  return (size_t)KM_RED_BLOCKS * N * C * 4 * sizeof(float);
}
extern "C" size_t km_pair_stats_workspace_bytes(int N, int C, long long M, int hard) {
  return pair_partials_bytes(N, C) + (hard ? (size_t)N * (size_t)M * sizeof(int32_t) : 0);
}

template <int COORD, int CCH>
static void launch_warp_loss(bool fast, dim3 grid, size_t smem, cudaStream_t st, const float* a,
                             const float* theta, int K, const float* g, const float* mov,
                             const float* fix, float* out, float* gout, float* part, int N, int C,
                             int D, int H, int W, int mode) {
  if (fast)
    warp_loss_kernel<COORD, CCH, true><<<grid, 256, smem, st>>>(a, theta, K, g, mov, fix, out, gout,
                                                                part, N, C, D, H, W, mode);
  else
    warp_loss_kernel<COORD, CCH, false><<<grid, 256, smem, st>>>(a, theta, K, g, mov, fix, out, gout,
                                                                 part, N, C, D, H, W, mode);
}

extern "C" int km_warp_loss(int coord_mode, const float* mat_or_ctrl, const float* theta, int K,
                            const float* grid, const float* moving, const float* fixed, float* out,
                            float* grid_out, double* sums, void* workspace, int N, int C, int D,
                            int H, int W, int mode, km_stream_t stream) {
  KM_CHECK_ARG(moving && N > 0 && C > 0 && D > 0 && H > 0 && W > 0, "km_warp_loss: bad arguments");
  KM_CHECK_ARG((long long)D * H * W < (1ll << 31), "km_warp_loss: volume too large");
  KM_CHECK_ARG(mode == KM_INTERP_BILINEAR || mode == KM_INTERP_NEAREST, "km_warp_loss: bad mode");
  KM_CHECK_ARG(!fixed || (sums && workspace), "km_warp_loss: sums/workspace required with fixed");
  size_t smem = 0;
  if (coord_mode == KM_COORD_AFFINE) {
    KM_CHECK_ARG(mat_or_ctrl, "km_warp_loss: affine matrix missing");
  } else if (coord_mode == KM_COORD_TPS) {
    KM_CHECK_ARG(mat_or_ctrl && theta && K > 0, "km_warp_loss: TPS parameters missing");
    smem = (size_t)(2 * K + 3) * sizeof(float4);
    KM_CHECK_ARG(smem <= 32 * 1024, "km_warp_loss: K=%d too large", K);
  } else if (coord_mode == KM_COORD_GRID) {
    KM_CHECK_ARG(grid, "km_warp_loss: grid missing");
    KM_CHECK_ARG(!grid_out, "km_warp_loss: grid_out is for the affine / TPS coordinate modes");
  } else {
    km_set_error("km_warp_loss: bad coord_mode %d", coord_mode);
    return KM_EINVAL;
  }
  const dim3 g(KM_RED_BLOCKS, N);
  float* part = reinterpret_cast<float*>(workspace);
  cudaStream_t st = km_cs(stream);
  const bool fast = g_tps_fast != 0;
#define KM_WL(COORD)                                                                              \
  do {                                                                                            \
    if (C == 1)                                                                                   \
      launch_warp_loss<COORD, 1>(fast, g, smem, st, mat_or_ctrl, theta, K, grid, moving, fixed,   \
                                 out, grid_out, part, N, C, D, H, W, mode);                       \
    else                                                                                          \
      launch_warp_loss<COORD, 16>(fast, g, smem, st, mat_or_ctrl, theta, K, grid, moving, fixed,  \
                                  out, grid_out, part, N, C, D, H, W, mode);                      \
  } while (0)
  if (coord_mode == KM_COORD_AFFINE) KM_WL(KM_COORD_AFFINE);
  else if (coord_mode == KM_COORD_TPS) KM_WL(KM_COORD_TPS);
  else KM_WL(KM_COORD_GRID);
#undef KM_WL
  KM_LAUNCH_OK("warp_loss_kernel");
  if (fixed) {
    const int NC4 = N * C * 4;
    sum_partials4_kernel<<<(NC4 + 7) / 8, 256, 0, st>>>(part, KM_RED_BLOCKS, NC4, sums);
    KM_LAUNCH_OK("sum_partials4_kernel");
  }
  return KM_OK;
}

extern "C" size_t km_warp_labels_workspace_bytes(int N, int C) {
  return (size_t)KM_RED_BLOCKS * N * C * 4 * sizeof(float) + (size_t)N * C * 3 * sizeof(unsigned long long);
}

extern "C" int km_warp_labels_dice(int coord_mode, const float* mat, const float* grid,
                                   const uint8_t* labels_m, const uint8_t* labels_f, uint8_t* labels_out,
                                   double* soft_sums, double* hard_sums, void* workspace, int N, int C,
                                   int D, int H, int W, km_stream_t stream) {
  KM_CHECK_ARG(labels_m && labels_f && soft_sums && hard_sums && workspace, "km_warp_labels_dice: null argument");
  KM_CHECK_ARG(N > 0 && C > 0 && C <= 255 && D > 0 && H > 0 && W > 0, "km_warp_labels_dice: bad shape (C <= 255)");
  KM_CHECK_ARG((long long)D * H * W < (1ll << 31), "km_warp_labels_dice: volume too large");
  KM_CHECK_ARG((coord_mode == KM_COORD_AFFINE && mat) || (coord_mode == KM_COORD_GRID && grid),
               "km_warp_labels_dice: coord_mode must be KM_COORD_AFFINE (mat) or KM_COORD_GRID (grid)");
  float* part = reinterpret_cast<float*>(workspace);
  unsigned long long* hard = reinterpret_cast<unsigned long long*>(
      reinterpret_cast<uint8_t*>(workspace) + (size_t)KM_RED_BLOCKS * N * C * 4 * sizeof(float));
  cudaStream_t st = km_cs(stream);
  KM_CUDA_OK(cudaMemsetAsync(hard, 0, (size_t)N * C * 3 * sizeof(unsigned long long), st));
  const size_t smem = ((size_t)8 * C * 4 + (size_t)C * 3) * sizeof(float);
  KM_CHECK_ARG(smem <= 48 * 1024, "km_warp_labels_dice: too many classes for shared memory");
  const dim3 g(KM_RED_BLOCKS, N);
  if (coord_mode == KM_COORD_AFFINE)
    warp_labels_kernel<KM_COORD_AFFINE><<<g, 256, smem, st>>>(mat, grid, labels_m, labels_f, labels_out, part,
                                                              hard, N, C, D, H, W);
  else
    warp_labels_kernel<KM_COORD_GRID><<<g, 256, smem, st>>>(mat, grid, labels_m, labels_f, labels_out, part,
                                                            hard, N, C, D, H, W);
  KM_LAUNCH_OK("warp_labels_kernel");
  const int NC4 = N * C * 4;
  sum_partials4_kernel<<<(NC4 + 7) / 8, 256, 0, st>>>(part, KM_RED_BLOCKS, NC4, soft_sums);
  KM_LAUNCH_OK("sum_partials4_kernel");
  hard_counts_to_sums_kernel<<<(N * C + 127) / 128, 128, 0, st>>>(hard, hard_sums, N * C);
  KM_LAUNCH_OK("hard_counts_to_sums_kernel");
  return KM_OK;
}

extern "C" size_t km_jacobian_stats_workspace_bytes(int N) {
  return (size_t)KM_RED_BLOCKS * N * 3 * sizeof(double);
}

extern "C" int km_jacobian_stats(const float* field, long long stride_n, long long stride_c,
                                 long long stride_z, long long stride_y, long long stride_x, double* out,
                                 void* workspace, int N, int D, int H, int W, km_stream_t stream) {
  KM_CHECK_ARG(field && out && workspace && N > 0, "km_jacobian_stats: bad arguments");
  KM_CHECK_ARG(D > 4 && H > 4 && W > 4, "km_jacobian_stats: the field must be larger than the 2-voxel crop (got %dx%dx%d)", D, H, W);
  double* part = reinterpret_cast<double*>(workspace);
  jacobian_stats_kernel<<<dim3(KM_RED_BLOCKS, N), 256, 0, km_cs(stream)>>>(field, stride_n, stride_c, stride_z,
                                                                           stride_y, stride_x, part, N, D, H, W);
  KM_LAUNCH_OK("jacobian_stats_kernel");
  const double count = (double)(D - 4) * (H - 4) * (W - 4);
  jacobian_finalize_kernel<<<N, 32, 0, km_cs(stream)>>>(part, KM_RED_BLOCKS, count, out, N);
  KM_LAUNCH_OK("jacobian_finalize_kernel");
  return KM_OK;
}

extern "C" int km_argmax_channels(const float* pred, int32_t* labels, int N, int C, long long M,
                                  km_stream_t stream) {
  KM_CHECK_ARG(pred && labels && N > 0 && C > 0 && M > 0, "km_argmax_channels: bad arguments");
  argmax_channels_kernel<<<dim3(blocks_for(M, 256, 148 * 16), N), 256, 0, km_cs(stream)>>>(
      pred, labels, C, M);
  KM_LAUNCH_OK("argmax_channels_kernel");
  return KM_OK;
}

extern "C" int km_pair_stats(const float* pred, const float* target, double* sums, void* workspace,
                             int N, int C, long long M, int hard, km_stream_t stream) {
  KM_CHECK_ARG(pred && target && sums && workspace && N > 0 && C > 0 && M > 0,
               "km_pair_stats: bad arguments");
  KM_CHECK_ARG(C <= 65535 && N <= 65535, "km_pair_stats: too many channels");
  float* part = reinterpret_cast<float*>(workspace);
  // hard Dice: the label map lives behind the partials in the workspace
  int32_t* labels = nullptr;
  if (hard) {
    labels = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) +
                                        pair_partials_bytes(N, C));
    int rc = km_argmax_channels(pred, labels, N, C, M, stream);
    if (rc != KM_OK) return rc;
  }
  // fewer blocks per channel when there are many channels: keep ~148*8 blocks in flight
  int bx = KM_RED_BLOCKS;
  pair_stats_kernel<<<dim3(bx, C, N), KM_RED_THREADS, 0, km_cs(stream)>>>(pred, target, labels, part,
                                                                          N, C, M);
  KM_LAUNCH_OK("pair_stats_kernel");
  const int NC4 = N * C * 4;
  sum_partials4_kernel<<<(NC4 + 7) / 8, 256, 0, km_cs(stream)>>>(part, bx, NC4, sums);
  KM_LAUNCH_OK("sum_partials4_kernel");
  return KM_OK;
}
