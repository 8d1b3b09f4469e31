// HBM-bound companions of the tcgen05 convolution (the 1-channel stem lives in stem.cu): GroupNorm /
// InstanceNorm statistics + application (with the decoder's nearest-upsample + concat folded in),
// MaxPool3d(2) and layout converters.  All reductions are two-stage and deterministic: every block
// writes one partial slot, the finalize kernel adds the slots in a fixed order in fp64.
//
// Reference call sites: keymorph/unet3d/buildingblocks.py:39-132 (GroupNorm -> Conv3d -> ReLU),
// :360-389 (MaxPool3d before the encoder's DoubleConv), :464-475,580-582 (nearest upsample + cat),
// keymorph/layers.py:137-187 (Conv3d -> InstanceNorm3d -> ReLU -> MaxPool3d).
#include "km_common.cuh"

namespace {

typedef uint16_t act16;   // one 16-bit activation / weight element: fp16 or bf16 bits (template parameter F16)

int g_operand_fp16 = 1;   // km_set_option(KM_OPT_OPERAND_FP16): fp16 (default) or bf16 operands

template <bool F16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = km_unpack2<F16>(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <bool F16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  uint32_t* p = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = km_pack2<F16>(f[2 * i], f[2 * i + 1]);
  return v;
}

// launch KERNEL<true> or KERNEL<false> according to the operand option
#define KM_LAUNCH_16(KERNEL, GRID, BLOCK, SMEM, STREAM, ...)                    \
  do {                                                                          \
    if (g_operand_fp16) KERNEL<true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__); \
    else KERNEL<false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);               \
  } while (0)

// ------------------------------------------------------------------------------------------
// sum / sumsq of an fp32 volume, grid (KM_RED_BLOCKS, N)
__global__ void __launch_bounds__(KM_RED_THREADS)
volume_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, long long M, int N) {
  const int n = blockIdx.y;
  const float* xn = x + (size_t)n * M;
  float s = 0.f, ss = 0.f;
  const long long M4 = (((uintptr_t)xn & 15) == 0) ? (M / 4) : 0;
  const float4* x4 = reinterpret_cast<const float4*>(xn);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  for (long long i = M4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = xn[i];
    s += v;
    ss += v * v;
  }
  __shared__ float red[2][KM_RED_THREADS / 32];
  s = km_warp_sum(s);
  ss = km_warp_sum(ss);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < KM_RED_THREADS / 32; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    float* d = stats + ((size_t)blockIdx.x * N + n) * 2;
    d[0] = a;
    d[1] = b;
  }
}

// ------------------------------------------------------------------------------------------
// generic per-channel stats of a bf16 NDHWC tensor, grid (KM_RED_BLOCKS, N)
template <bool F16>
__global__ void __launch_bounds__(KM_RED_THREADS)
channel_stats_kernel(const act16* __restrict__ x, float* __restrict__ stats, long long nvox, int C,
                     int N) {
  extern __shared__ float sacc[];  // [C][2]
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = C / 8;
  const long long total = nvox * cg;
  const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)n * nvox * C);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    float f[8];
    unpack8<F16>(__ldg(src + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&sacc[(g * 8 + k) * 2 + 0], f[k]);
      atomicAdd(&sacc[(g * 8 + k) * 2 + 1], f[k] * f[k]);
    }
  }
  __syncthreads();
  float* d = stats + ((size_t)blockIdx.x * N + n) * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) d[i] = sacc[i];
}

// ------------------------------------------------------------------------------------------
// GroupNorm / InstanceNorm finalize: one block per (sample, group), 512 threads (one warp per channel of the
// group at a time: with 4 warps the 12 - 48 channels of a group were walked serially, 10 us per launch x 12 launches)
constexpr int kFinalizeThreads = 512;
__global__ void __launch_bounds__(kFinalizeThreads)
norm_finalize_kernel(const float* __restrict__ stats0, int nparts0, int C0, double count0,
                     const float* __restrict__ stats1, int nparts1, int C1, double count1, double rep1,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int groups,
                     float eps, float* __restrict__ scale, float* __restrict__ shift, int N) {
  __shared__ double wsum[kFinalizeThreads / 32][3];
  const int n = blockIdx.x, g = blockIdx.y;
  const int C = C0 + C1;
  const int cpg = C / groups;
  // one warp per channel of the group: lanes stride over the partial slots, fixed-order shuffle
  // reduction; the warp then carries its channels' totals
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double gs = 0.0, gss = 0.0, gcnt = 0.0;
  for (int c = g * cpg + wid; c < (g + 1) * cpg; c += nwarps) {
    double s = 0.0, ss = 0.0;
    if (c < C0) {
      for (int p = lane; p < nparts0; p += 32) {
        const float2 q = *reinterpret_cast<const float2*>(stats0 + (((size_t)p * N + n) * C0 + c) * 2);
        s += (double)q.x;
        ss += (double)q.y;
      }
    } else {
      const int c1 = c - C0;
      for (int p = lane; p < nparts1; p += 32) {
        const float2 q = *reinterpret_cast<const float2*>(stats1 + (((size_t)p * N + n) * C1 + c1) * 2);
        s += (double)q.x;
        ss += (double)q.y;
      }
      s *= rep1;
      ss *= rep1;
    }
    gs += km_warp_sum(s);
    gss += km_warp_sum(ss);
    gcnt += (c < C0) ? count0 : count1 * rep1;
  }
  if (lane == 0) {
    wsum[wid][0] = gs;
    wsum[wid][1] = gss;
    wsum[wid][2] = gcnt;
  }
  __syncthreads();
  double s = 0.0, ss = 0.0, cnt = 0.0;
  for (int w = 0; w < nwarps; ++w) {
    s += wsum[w][0];
    ss += wsum[w][1];
    cnt += wsum[w][2];
  }
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;  // biased variance (torch group_norm / instance_norm)
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += blockDim.x) {
    const double ga = gamma ? (double)gamma[c] : 1.0;
    const double be = beta ? (double)beta[c] : 0.0;
    scale[(size_t)n * C + c] = (float)(ga * rstd);
    shift[(size_t)n * C + c] = (float)(be - mean * rstd * ga);
  }
}

// ------------------------------------------------------------------------------------------
// out = act(scale*src + shift) with optional second (nearest-upsampled) source.
// grid (ceil(W*cg/256), ceil(H/kNormRows), N*D): (n, z) is uniform per block, a thread owns one
// 16-byte channel group of one x position and walks kNormRows rows -> one 32-bit division per
// thread, no 64-bit index arithmetic, scale/shift loaded once.
constexpr int kNormRows = 8;
template <bool F16>
__global__ void __launch_bounds__(256)
norm_apply_kernel(const act16* __restrict__ src0, int C0, const act16* __restrict__ src1, int C1,
                  int D1, int H1, int W1, const float* __restrict__ scale,
                  const float* __restrict__ shift, act16* __restrict__ out, int N, int D, int H,
                  int W, int relu) {
  const int C = C0 + C1;
  const int cg = C / 8, cg0 = C0 / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // chunk index inside a row
  if (i >= W * cg) return;
  const int x = i / cg, g = i - x * cg;
  const int n = blockIdx.z / D, z = blockIdx.z - n * D;
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)n * C + g * 8);
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)n * C + g * 8);
  const float4 a0 = __ldg(sc), a1 = __ldg(sc + 1), b0 = __ldg(sh), b1 = __ldg(sh + 1);
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  // ATen upsample_nearest3d: src = min(floor(dst * (in / out)), in - 1) with a float scale
  const bool from1 = g >= cg0;
  const int x1 = from1 ? min((int)floorf(x * ((float)W1 / (float)W)), W1 - 1) : 0;
  const int z1 = from1 ? min((int)floorf(z * ((float)D1 / (float)D)), D1 - 1) : 0;
  const float sy = from1 ? (float)H1 / (float)H : 0.f;
  const int y_lo = blockIdx.y * kNormRows, y_hi = min(H, y_lo + kNormRows);
  uint4 raw[kNormRows];
#pragma unroll
  for (int r = 0; r < kNormRows; ++r) {
    const int y = y_lo + r;
    if (y < y_hi) {
      if (!from1) {
        raw[r] = __ldg(reinterpret_cast<const uint4*>(src0 + (((size_t)blockIdx.z * H + y) * W + x) * C0) + g);
      } else {
        const int y1 = min((int)floorf(y * sy), H1 - 1);
        const size_t v1 = (((size_t)n * D1 + z1) * H1 + y1) * W1 + x1;
        raw[r] = __ldg(reinterpret_cast<const uint4*>(src1 + v1 * C1) + (g - cg0));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kNormRows; ++r) {
    const int y = y_lo + r;
    if (y < y_hi) {
      float f[8];
      unpack8<F16>(raw[r], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = fmaf(a[k], f[k], b[k]);
        if (relu) f[k] = fmaxf(f[k], 0.f);
      }
      reinterpret_cast<uint4*>(out + (((size_t)blockIdx.z * H + y) * W + x) * C)[g] = pack8<F16>(f);
    }
  }
}

// normalise + activation + MaxPool3d(2) in one pass (ConvNet blocks 2/4/6/8)
template <bool F16>
__global__ void __launch_bounds__(256)
norm_apply_pool_kernel(const act16* __restrict__ src, int C, const float* __restrict__ scale,
                       const float* __restrict__ shift, act16* __restrict__ out, int N, int D, int H,
                       int W, int relu) {
  const int cg = C / 8;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long nvo = (long long)Do * Ho * Wo;
  const long long total = (long long)N * nvo * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long v = i / cg;
    const int n = (int)(v / nvo);
    const long long r = v % nvo;
    const int x = (int)(r % Wo), y = (int)((r / Wo) % Ho), z = (int)(r / ((long long)Wo * Ho));
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)n * C + g * 8);
    const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)n * C + g * 8);
    const float4 a0 = __ldg(sc), a1 = __ldg(sc + 1), b0 = __ldg(sh), b1 = __ldg(sh + 1);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t vi = (((size_t)n * D + (2 * z + dz)) * H + (2 * y + dy)) * W + (2 * x + dx);
          float f[8];
          unpack8<F16>(__ldg(reinterpret_cast<const uint4*>(src + vi * C) + g), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], fmaf(a[k], f[k], b[k]));
        }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], 0.f);
    }
    reinterpret_cast<uint4*>(out)[i] = pack8<F16>(m);
  }
}

// ------------------------------------------------------------------------------------------
// MaxPool3d(2) + per-channel stats of the pooled tensor, grid (KM_RED_BLOCKS, N), 256 threads.
// Requires (gridDim.x * 256) % (C/8) == 0 and 256 % (C/8) == 0 so that a thread keeps one channel
// group for its whole grid-stride loop (checked on the host).
template <bool F16>
__global__ void __launch_bounds__(256)
maxpool2_stats_kernel(const act16* __restrict__ src, act16* __restrict__ out,
                      float* __restrict__ stats, int N, int C, int D, int H, int W) {
  __shared__ float red[256][17];
  const int n = blockIdx.y;
  const int cg = C / 8;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long nvo = (long long)Do * Ho * Wo;
  const long long total = nvo * cg;
  const act16* sn = src + (size_t)n * D * H * W * C;
  uint4* on = reinterpret_cast<uint4*>(out + (size_t)n * nvo * C);
  float s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = ss[k] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long r = i / cg;
    const int x = (int)(r % Wo), y = (int)((r / Wo) % Ho), z = (int)(r / ((long long)Wo * Ho));
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t vi = ((size_t)(2 * z + dz) * H + (2 * y + dy)) * W + (2 * x + dx);
          float f[8];
          unpack8<F16>(__ldg(reinterpret_cast<const uint4*>(sn + vi * C) + g), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
        }
    on[i] = pack8<F16>(m);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += m[k];
      ss[k] = fmaf(m[k], m[k], ss[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[threadIdx.x][k] = s[k];
    red[threadIdx.x][8 + k] = ss[k];
  }
  __syncthreads();
  // channel c = g*8 + k is owned by the threads t with t % cg == g
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / 8, k = c % 8;
    float a = 0.f, b = 0.f;
    for (int t = g; t < 256; t += cg) {
      a += red[t][k];
      b += red[t][8 + k];
    }
    float* d = stats + (((size_t)blockIdx.x * N + n) * C + c) * 2;
    d[0] = a;
    d[1] = b;
  }
}

// ------------------------------------------------------------------------------------------
template <bool F16>
__global__ void ndhwc_to_ncdhw_kernel(const act16* __restrict__ src, float* __restrict__ dst, int N,
                                      int C, long long nvox) {
  const long long total = (long long)N * C * nvox;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % nvox;
    const int c = (int)((i / nvox) % C);
    const int n = (int)(i / (nvox * C));
    dst[i] = km_to_float<F16>(src[((size_t)n * nvox + v) * C + c]);
  }
}
template <bool F16>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, act16* __restrict__ dst, int N,
                                      int C, long long nvox) {
  const long long total = (long long)N * C * nvox;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long v = (i / C) % nvox;
    const int n = (int)(i / (nvox * C));
    dst[i] = km_from_float<F16>(src[((size_t)n * C + c) * nvox + v]);
  }
}

inline int blocks_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = 148ll * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int km_operand_fp16() { return g_operand_fp16; }
void km_set_operand_fp16(int v) { g_operand_fp16 = v ? 1 : 0; }

extern "C" int km_pool_nparts(void) { return KM_RED_BLOCKS; }

extern "C" int km_volume_stats(const float* x, float* stats, int N, long long M,
                               km_stream_t stream) {
  KM_CHECK_ARG(x && stats && N > 0 && M > 0, "km_volume_stats: bad arguments");
  volume_stats_kernel<<<dim3(KM_RED_BLOCKS, N), KM_RED_THREADS, 0, km_cs(stream)>>>(x, stats, M, N);
  KM_LAUNCH_OK("volume_stats_kernel");
  return KM_OK;
}

extern "C" int km_channel_stats(const void* x, float* stats, int N, int C, long long nvox,
                                km_stream_t stream) {
  KM_CHECK_ARG(x && stats && N > 0 && C % 8 == 0 && nvox > 0, "km_channel_stats: bad arguments");
  KM_LAUNCH_16(channel_stats_kernel, dim3(KM_RED_BLOCKS, N), KM_RED_THREADS, 2 * C * sizeof(float), km_cs(stream), reinterpret_cast<const act16*>(x), stats, nvox, C, N);
  KM_LAUNCH_OK("channel_stats_kernel");
  return KM_OK;
}

extern "C" int km_norm_finalize(const float* stats0, int nparts0, int C0, double count0,
                                const float* stats1, int nparts1, int C1, double count1,
                                double rep1, const float* gamma, const float* beta, int groups,
                                float eps, float* scale, float* shift, int N, km_stream_t stream) {
  KM_CHECK_ARG(stats0 && C0 > 0 && nparts0 > 0 && scale && shift && N > 0,
               "km_norm_finalize: bad arguments");
  KM_CHECK_ARG(C1 == 0 || (stats1 && nparts1 > 0), "km_norm_finalize: second source missing");
  const int C = C0 + C1;
  KM_CHECK_ARG(groups > 0 && C % groups == 0, "km_norm_finalize: %d channels not divisible by %d groups",
               C, groups);
  KM_CHECK_ARG(groups <= 65535, "km_norm_finalize: too many groups");
  norm_finalize_kernel<<<dim3(N, groups), kFinalizeThreads, 0, km_cs(stream)>>>(stats0, nparts0, C0, count0, stats1, nparts1,
                                                                  C1, count1, rep1, gamma, beta, groups, eps,
                                                                  scale, shift, N);
  KM_LAUNCH_OK("norm_finalize_kernel");
  return KM_OK;
}

extern "C" int km_norm_apply(const void* src0, int C0, const void* src1, int C1, int D1, int H1,
                             int W1, const float* scale, const float* shift, void* out, int N,
                             int D, int H, int W, int relu, int pool, km_stream_t stream) {
  KM_CHECK_ARG(src0 && scale && shift && out && N > 0 && D > 0 && H > 0 && W > 0,
               "km_norm_apply: bad arguments");
  KM_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C0 > 0, "km_norm_apply: channels must be multiples of 8");
  KM_CHECK_ARG(C1 == 0 || src1, "km_norm_apply: second source missing");
  if (pool) {
    KM_CHECK_ARG(C1 == 0, "km_norm_apply: pool is only supported for a single source");
    KM_CHECK_ARG(D >= 2 && H >= 2 && W >= 2, "km_norm_apply: volume too small to pool");
    const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (C0 / 8);
    KM_LAUNCH_16(norm_apply_pool_kernel, blocks_for(total, 256), 256, 0, km_cs(stream), 
        reinterpret_cast<const act16*>(src0), C0, scale, shift, reinterpret_cast<act16*>(out), N, D, H,
        W, relu);
    KM_LAUNCH_OK("norm_apply_pool_kernel");
    return KM_OK;
  }
  KM_CHECK_ARG(H <= 65535 && (long long)N * D <= 65535, "km_norm_apply: volume too large for the grid");
  const dim3 grid((W * ((C0 + C1) / 8) + 255) / 256, (H + kNormRows - 1) / kNormRows, N * D);
  KM_LAUNCH_16(norm_apply_kernel, grid, 256, 0, km_cs(stream), 
      reinterpret_cast<const act16*>(src0), C0, reinterpret_cast<const act16*>(src1), C1,
      C1 ? D1 : 1, C1 ? H1 : 1, C1 ? W1 : 1, scale, shift, reinterpret_cast<act16*>(out), N, D, H, W,
      relu);
  KM_LAUNCH_OK("norm_apply_kernel");
  return KM_OK;
}

extern "C" int km_maxpool2_stats(const void* src, void* out, float* stats, int N, int C, int D,
                                 int H, int W, km_stream_t stream) {
  KM_CHECK_ARG(src && out && stats && N > 0, "km_maxpool2_stats: bad arguments");
  KM_CHECK_ARG(C % 8 == 0 && 256 % (C / 8) == 0,
               "km_maxpool2_stats: C/8 must divide 256 (C=%d)", C);
  KM_CHECK_ARG(D >= 2 && H >= 2 && W >= 2, "km_maxpool2_stats: volume too small");
  KM_LAUNCH_16(maxpool2_stats_kernel, dim3(KM_RED_BLOCKS, N), 256, 0, km_cs(stream), 
      reinterpret_cast<const act16*>(src), reinterpret_cast<act16*>(out), stats, N, C, D, H, W);
  KM_LAUNCH_OK("maxpool2_stats_kernel");
  return KM_OK;
}

// GroupNorm folded into the consumer conv (keymorph/unet3d/buildingblocks.py:50-52, order "gcr": GN -> conv
// -> ReLU with nothing non-linear between GN and the conv): conv(scale x + shift) with zero padding of the
// normalised input = conv_{w scale}(x) + sum over the IN-BOUNDS taps of w . shift.  For every sample n:
//   packed[n] = bf16(w * scale[n]) in the layout the conv kernel streams
//               (layout 0: [tap][Cout][Cin]; layout 1, z-folded: [rot][dx][dy][j][Cout][Cin], dz = (rot+1-j) mod 3)
//   bias[n][class][cout], class = zcode * 9 + ycode * 3 + xcode; code bit 0 = voxel on the low border of that
//               axis (tap offset -1 outside), bit 1 = on the high border (tap offset +1 outside)
// Blocks [0, pack_blocks) pack (one thread per (cout, cin) pair); the others build the tables, one block per
// (cout, sample) with the threads over the input channels: the 36 class sums of a 3x3x3 kernel come from three
// separable reductions.
template <bool F16>
__global__ void __launch_bounds__(256)
fold_gn_kernel(const float* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ shift,
               act16* __restrict__ packed, float* __restrict__ bias, int N, int Cout, int Cin, int layout,
               int pack_blocks, int c0, int Cp) {
  // the packed weights cover input channels [c0, c0 + Cp) (Cp = Cin unless the layer is split over two kernels,
  // conv_up2.cu); the bias tables always sum over all Cin channels
  if ((int)blockIdx.x < pack_blocks) {
    // one thread per (cout, cin): its 27 taps are contiguous in w (a warp reads one contiguous span), and every
    // packed position is written by consecutive threads to consecutive elements
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= Cout * Cp) return;
    const int co = t / Cp, cl = t % Cp, ci = c0 + cl;
    float wk[27];
    const float* src = w + ((size_t)co * Cin + ci) * 27;
#pragma unroll
    for (int k = 0; k < 27; ++k) wk[k] = src[k];
    const size_t plane = (size_t)Cout * Cp;
    const size_t total = (size_t)27 * (layout ? 3 : 1) * plane;
    for (int n = 0; n < N; ++n) {
      const float sc = scale[n * Cin + ci];
      act16* dst = packed + (size_t)n * total + (size_t)co * Cp + cl;
      if (layout) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                dst[(size_t)(((r * 3 + dx) * 3 + dy) * 3 + j) * plane] =
                    km_from_float<F16>(wk[((r + 1 - j + 3) % 3) * 9 + dy * 3 + dx] * sc);
      } else {
#pragma unroll
        for (int k = 0; k < 27; ++k) dst[(size_t)k * plane] = km_from_float<F16>(wk[k] * sc);
      }
    }
    return;
  }
  // bias tables: one block per (cout, sample), the warps split the input channels
  __shared__ float s_part[8][36];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bb = (int)blockIdx.x - pack_blocks;
  const int co = bb % Cout, n = bb / Cout;
  {
    float acc[36];
#pragma unroll
    for (int c = 0; c < 36; ++c) acc[c] = 0.f;
    for (int ci = threadIdx.x; ci < Cin; ci += 256) {
      const float* wk = w + ((size_t)co * Cin + ci) * 27;
      const float sh = shift[n * Cin + ci];
      float xs[3][3][3];   // [dz][dy][x code]: mid = all three dx, low border = dx 1..2, high border = dx 0..1
#pragma unroll
      for (int d = 0; d < 9; ++d) {
        const float w0 = wk[d * 3], w1 = wk[d * 3 + 1], w2 = wk[d * 3 + 2];
        xs[d / 3][d % 3][0] = w0 + w1 + w2;
        xs[d / 3][d % 3][1] = w1 + w2;
        xs[d / 3][d % 3][2] = w0 + w1;
      }
#pragma unroll
      for (int xc = 0; xc < 3; ++xc) {
        float ys[3][3];    // [dz][y code]
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
          const float a0 = xs[dz][0][xc], a1 = xs[dz][1][xc], a2 = xs[dz][2][xc];
          ys[dz][0] = a0 + a1 + a2;
          ys[dz][1] = a1 + a2;
          ys[dz][2] = a0 + a1;
        }
#pragma unroll
        for (int yc = 0; yc < 3; ++yc) {
          const float b0 = ys[0][yc], b1 = ys[1][yc], b2 = ys[2][yc];
          acc[(0 * 3 + yc) * 3 + xc] = fmaf(b0 + b1 + b2, sh, acc[(0 * 3 + yc) * 3 + xc]);
          acc[(1 * 3 + yc) * 3 + xc] = fmaf(b1 + b2, sh, acc[(1 * 3 + yc) * 3 + xc]);
          acc[(2 * 3 + yc) * 3 + xc] = fmaf(b0 + b1, sh, acc[(2 * 3 + yc) * 3 + xc]);
          acc[(3 * 3 + yc) * 3 + xc] = fmaf(b1, sh, acc[(3 * 3 + yc) * 3 + xc]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 36; ++c) {
      const float v = km_warp_sum(acc[c]);
      if (lane == 0) s_part[wid][c] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s_part[k][threadIdx.x];
    bias[((size_t)n * 36 + threadIdx.x) * Cout + co] = v;
  }
}

int km_fold_gn_part(const float* w, const float* scale, const float* shift, void* packed, float* bias, int N,
                    int Cout, int Cin, int c0, int Cp, int layout, km_stream_t stream) {
  const int pack_blocks = (Cout * Cp + 255) / 256;
  const int bias_blocks = Cout * N;
  KM_LAUNCH_16(fold_gn_kernel, pack_blocks + bias_blocks, 256, 0, km_cs(stream), w, scale, shift, reinterpret_cast<act16*>(packed),
                                                                       bias, N, Cout, Cin, layout, pack_blocks, c0, Cp);
  KM_LAUNCH_OK("fold_gn_kernel");
  return KM_OK;
}

int km_fold_gn(const float* w, const float* scale, const float* shift, void* packed, float* bias, int N, int Cout,
               int Cin, int layout, km_stream_t stream) {
  return km_fold_gn_part(w, scale, shift, packed, bias, N, Cout, Cin, 0, Cin, layout, stream);
}

// nearest-neighbour x2 upsampling of a bf16 NDHWC tensor (F.interpolate(scale 2, 'nearest') in the
// decoders, keymorph/unet3d/buildingblocks.py:409-445): every 16-byte chunk of a coarse voxel is read
// once and written to its 2x2x2 fine voxels
__global__ void __launch_bounds__(256)
upsample2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long chunks, int cpv, int Dc,
                 int Hc, int Wc) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < chunks; i += 256ll * gridDim.x) {
    const int c = (int)(i % cpv);
    long long v = i / cpv;
    const int x = (int)(v % Wc);
    v /= Wc;
    const int y = (int)(v % Hc);
    v /= Hc;
    const int z = (int)(v % Dc);
    const long long n = v / Dc;
    const uint4 val = src[i];
    const long long W2 = 2ll * Wc, H2 = 2ll * Hc;
    const long long base = (((n * 2 * Dc + 2 * z) * H2 + 2 * y) * W2 + 2 * x) * cpv + c;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        uint4* o = dst + base + ((long long)dz * H2 + dy) * W2 * cpv;
        o[0] = val;
        o[cpv] = val;
      }
  }
}

extern "C" int km_upsample2_ndhwc(const void* src, void* dst, int N, int C, int Dc, int Hc, int Wc,
                                  km_stream_t stream) {
  KM_CHECK_ARG(src && dst && N > 0 && C > 0 && C % 8 == 0 && Dc > 0 && Hc > 0 && Wc > 0,
               "km_upsample2_ndhwc: bad arguments (C must be a multiple of 8)");
  KM_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "km_upsample2_ndhwc: pointers must be 16-byte aligned");
  const long long chunks = (long long)N * Dc * Hc * Wc * (C / 8);
  upsample2_kernel<<<blocks_for(chunks, 256), 256, 0, km_cs(stream)>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), chunks, C / 8, Dc, Hc, Wc);
  KM_LAUNCH_OK("upsample2_kernel");
  return KM_OK;
}

extern "C" int km_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, int D, int H,
                                          int W, km_stream_t stream) {
  KM_CHECK_ARG(src && dst && N > 0 && C > 0, "km_ndhwc_bf16_to_ncdhw_f32: bad arguments");
  const long long nvox = (long long)D * H * W;
  KM_LAUNCH_16(ndhwc_to_ncdhw_kernel, blocks_for((long long)N * C * nvox, 256), 256, 0, km_cs(stream), 
      reinterpret_cast<const act16*>(src), dst, N, C, nvox);
  KM_LAUNCH_OK("ndhwc_to_ncdhw_kernel");
  return KM_OK;
}
extern "C" int km_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, int D, int H,
                                          int W, km_stream_t stream) {
  KM_CHECK_ARG(src && dst && N > 0 && C > 0, "km_ncdhw_f32_to_ndhwc_bf16: bad arguments");
  const long long nvox = (long long)D * H * W;
  KM_LAUNCH_16(ncdhw_to_ndhwc_kernel, blocks_for((long long)N * C * nvox, 256), 256, 0, km_cs(stream), 
      src, reinterpret_cast<act16*>(dst), N, C, nvox);
  KM_LAUNCH_OK("ncdhw_to_ndhwc_kernel");
  return KM_OK;
}
