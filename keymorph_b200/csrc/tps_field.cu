// Dense thin-plate-spline flow field, row-structured (keymorph/keypoint_aligners.py:365-449:
// TPS.get_flow_field -> transform_points; radial basis :322-339).
//
// Work = K radial-basis terms per voxel (8.6e9 at 256^3, K = 512): the kernel is bound by instruction
// issue and the XU (MUFU) pipe, not by memory (12 B written per voxel).  So the design minimises the
// instructions per term:
//   * a warp owns one x-row segment, a thread VPT voxels of it (x = x0 + lane + 32 j): dz, dy and
//     dz^2 + dy^2 + 1e-6 depend only on (row, control point) and are computed once per VPT terms;
//   * U = r^2 log(r + 1e-6) ~= 0.5 ln2 (s lg2 s + (1e-6 / (0.5 ln2)) r), s = d^2 + 1e-6: ONE MUFU (lg2) per
//     term, r from the exponent-halving bit trick (its 3.5 % error sits in a 1e-6-sized term), and the
//     constant 0.5 ln2 is folded into the spline weights when they are staged in shared memory;
//   * the per-term FP32 work (dx, s, s*lg2 s, + r term, 3 accumulations) is issued as packed
//     FADD2 / FMUL2 / FFMA2 (Blackwell f32x2) on voxel PAIRS: 7 packed instructions per two terms.
// Issue slots per term: 19.7 (r01 kernel, ncu) -> ~7.5; the XU pipe (one MUFU per term, 16 lanes/clk/SM)
// is then the bound: K*N / (16 * 148 * f_SM) = 1.86 ms at 256^3, K = 512, 1.95 GHz.
// The exact variant (KM_OPT_TPS_FAST = 0: sqrtf + logf per term) keeps the same structure.
#include "km_common.cuh"

int km_tps_fast_enabled();   // warp.cu (km_set_option KM_OPT_TPS_FAST)

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

constexpr float kHalfLn2 = 0.34657359028f;           // 0.5 * ln 2
constexpr float kRTerm = 1e-6f / 0.34657359028f;      // coefficient of r once 0.5 ln2 is factored out

// MUFU.LG2 without the denormal pre-scaling __log2f wraps around it (4 extra instructions per term):
// the argument is d^2 + 1e-6 >= 1e-6, never denormal
__device__ __forceinline__ float lg2_fast(float s) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
  return r;
}

// sqrt(s) to 3.5 %: halve the exponent, linear in the mantissa
__device__ __forceinline__ float rough_sqrt(float s) {
  return __int_as_float((__float_as_int(s) >> 1) + 0x1fbd1df5);
}

// smem: c4[t] = (cz, cy, cx, -), w4[t] = weights (wz, wy, wx) [* 0.5 ln2 when FAST], aff[12]
template <bool FAST>
__device__ __forceinline__ void stage_tps(const float* __restrict__ ctrl, const float* __restrict__ theta, int K,
                                          float4* c4, float4* w4, float* aff) {
  const float sc = FAST ? kHalfLn2 : 1.f;
  for (int t = threadIdx.x; t < K; t += blockDim.x) {
    c4[t] = make_float4(ctrl[t * 3 + 0], ctrl[t * 3 + 1], ctrl[t * 3 + 2], 0.f);
    w4[t] = make_float4(theta[t * 3 + 0] * sc, theta[t * 3 + 1] * sc, theta[t * 3 + 2] * sc, 0.f);
  }
  for (int i = threadIdx.x; i < 12; i += blockDim.x) aff[i] = theta[K * 3 + i];
  __syncthreads();
}

// VPT voxels per thread (even), PACKED: f32x2 arithmetic on voxel pairs
template <bool FAST, int VPT, bool PACKED>
__global__ void __launch_bounds__(256)
flow_tps_rows_kernel(const float* __restrict__ ctrl, const float* __restrict__ theta, float* __restrict__ grid,
                     int K, int D, int H, int W) {
  extern __shared__ float4 s4[];
  float4* c4 = s4;
  float4* w4 = s4 + K;
  float* aff = reinterpret_cast<float*>(s4 + 2 * K);
  const int n = blockIdx.y;
  stage_tps<FAST>(ctrl + (size_t)n * K * 3, theta + (size_t)n * (K + 4) * 3, K, c4, w4, aff);
  const int lane = threadIdx.x & 31;
  constexpr int SEG = 32 * VPT;
  const int nseg = (W + SEG - 1) / SEG;
  const long long nchunks = (long long)D * H * nseg;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float* gn = grid + (size_t)n * D * H * W * 3;
  for (long long chunk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
    const int seg = (int)(chunk % nseg);
    const int row = (int)(chunk / nseg);
    const int y = row % H, z = row / H;
    const float pz = km_linspace(-1.f, 1.f, D, z), py = km_linspace(-1.f, 1.f, H, y);
    float px[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int x = seg * SEG + lane + 32 * j;
      px[j] = km_linspace(-1.f, 1.f, W, x < W ? x : W - 1);
    }
    float az[VPT], ay[VPT], ax[VPT];
    if (PACKED) {
      u64 px2[VPT / 2], az2[VPT / 2], ay2[VPT / 2], ax2[VPT / 2];
#pragma unroll
      for (int j = 0; j < VPT / 2; ++j) {
        px2[j] = pk(px[2 * j], px[2 * j + 1]);
        az2[j] = ay2[j] = ax2[j] = pk(0.f, 0.f);
      }
      const u64 kr2 = pk(kRTerm, kRTerm);
#pragma unroll 2
      for (int t = 0; t < K; ++t) {
        const float4 c = c4[t];
        const float4 w = w4[t];
        const float dz = pz - c.x, dy = py - c.y;
        const float base = fmaf(dz, dz, fmaf(dy, dy, 1e-6f));
        const u64 base2 = pk(base, base), ncx2 = pk(-c.z, -c.z);
        const u64 wz2 = pk(w.x, w.x), wy2 = pk(w.y, w.y), wx2 = pk(w.z, w.z);
#pragma unroll
        for (int j = 0; j < VPT / 2; ++j) {
          const u64 dx2 = add2(px2[j], ncx2);
          const u64 s2 = fma2(dx2, dx2, base2);
          float s0, s1;
          unpk(s2, s0, s1);
          u64 u2;
          if (FAST) {
            const u64 l2 = pk(lg2_fast(s0), lg2_fast(s1));
            const u64 r2 = pk(rough_sqrt(s0), rough_sqrt(s1));
            u2 = fma2(r2, kr2, mul2(s2, l2));
          } else {
            const float r0 = sqrtf(s0), r1 = sqrtf(s1);
            u2 = pk((r0 * r0) * logf(r0 + 1e-6f), (r1 * r1) * logf(r1 + 1e-6f));
          }
          az2[j] = fma2(u2, wz2, az2[j]);
          ay2[j] = fma2(u2, wy2, ay2[j]);
          ax2[j] = fma2(u2, wx2, ax2[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < VPT / 2; ++j) {
        unpk(az2[j], az[2 * j], az[2 * j + 1]);
        unpk(ay2[j], ay[2 * j], ay[2 * j + 1]);
        unpk(ax2[j], ax[2 * j], ax[2 * j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < VPT; ++j) az[j] = ay[j] = ax[j] = 0.f;
#pragma unroll 2
      for (int t = 0; t < K; ++t) {
        const float4 c = c4[t];
        const float4 w = w4[t];
        const float dz = pz - c.x, dy = py - c.y;
        const float base = fmaf(dz, dz, fmaf(dy, dy, 1e-6f));
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const float dx = px[j] - c.z;
          const float s = fmaf(dx, dx, base);
          float u;
          if (FAST) {
            u = fmaf(rough_sqrt(s), kRTerm, s * lg2_fast(s));
          } else {
            const float r = sqrtf(s);
            u = (r * r) * logf(r + 1e-6f);
          }
          az[j] = fmaf(u, w.x, az[j]);
          ay[j] = fmaf(u, w.y, ay[j]);
          ax[j] = fmaf(u, w.z, ax[j]);
        }
      }
    }
    // z = [1, p] . affine  (keymorph/keypoint_aligners.py:427-433), out = z + b, stored in (x,y,z) order
    const float bz = aff[0] + aff[3] * pz + aff[6] * py;
    const float by = aff[1] + aff[4] * pz + aff[7] * py;
    const float bx = aff[2] + aff[5] * pz + aff[8] * py;
    float* gr = gn + ((size_t)row * W + (size_t)seg * SEG) * 3;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int xo = lane + 32 * j;
      if (seg * SEG + xo < W) {
        gr[xo * 3 + 0] = (bx + aff[11] * px[j]) + ax[j];
        gr[xo * 3 + 1] = (by + aff[10] * px[j]) + ay[j];
        gr[xo * 3 + 2] = (bz + aff[9] * px[j]) + az[j];
      }
    }
  }
}

int g_tps_packed = 1;   // km_set_option(KM_OPT_TPS_PACKED): f32x2 arithmetic (A/B)
int g_tps_vpt = 0;      // km_set_option(KM_OPT_TPS_VPT): voxels per thread, 0 = from W

template <bool FAST, int VPT>
void launch_rows(bool packed, dim3 g, size_t smem, cudaStream_t st, const float* ctrl, const float* theta,
                 float* grid, int K, int D, int H, int W) {
  if (packed)
    flow_tps_rows_kernel<FAST, VPT, true><<<g, 256, smem, st>>>(ctrl, theta, grid, K, D, H, W);
  else
    flow_tps_rows_kernel<FAST, VPT, false><<<g, 256, smem, st>>>(ctrl, theta, grid, K, D, H, W);
}

}  // namespace

void km_tps_set_packed(int v) { g_tps_packed = v ? 1 : 0; }
void km_tps_set_vpt(int v) { g_tps_vpt = v; }

extern "C" int km_flow_field_tps(const float* ctrl, const float* theta, float* grid, int N, int K, int D, int H,
                                 int W, km_stream_t stream) {
  KM_CHECK_ARG(ctrl && theta && grid && N > 0 && K > 0 && D > 0 && H > 0 && W > 0,
               "km_flow_field_tps: bad arguments");
  const size_t smem = (size_t)(2 * K + 3) * sizeof(float4);
  KM_CHECK_ARG(smem <= 48 * 1024, "km_flow_field_tps: K=%d too large", K);
  int vpt = g_tps_vpt;
  if (vpt != 2 && vpt != 4 && vpt != 8) vpt = W > 128 ? 8 : (W > 64 ? 4 : 2);
  const long long nchunks = (long long)D * H * ((W + 32 * vpt - 1) / (32 * vpt));
  long long blocks = (nchunks + 7) / 8;
  const long long cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  const dim3 g((unsigned)blocks, N);
  cudaStream_t st = km_cs(stream);
  const bool fast = km_tps_fast_enabled() != 0, packed = g_tps_packed != 0;
#define KM_TPS_ROWS(V)                                                                     \
  do {                                                                                     \
    if (fast) launch_rows<true, V>(packed, g, smem, st, ctrl, theta, grid, K, D, H, W);    \
    else launch_rows<false, V>(packed, g, smem, st, ctrl, theta, grid, K, D, H, W);        \
  } while (0)
  if (vpt == 8) KM_TPS_ROWS(8);
  else if (vpt == 4) KM_TPS_ROWS(4);
  else KM_TPS_ROWS(2);
#undef KM_TPS_ROWS
  KM_LAUNCH_OK("flow_tps_rows_kernel");
  return KM_OK;
}
