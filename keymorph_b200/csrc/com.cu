// Centre-of-mass keypoint layer (keymorph/layers.py:78-134) and power weights
// (keymorph/model.py:95-109) as ONE pass over the heat map: the three marginal sums of the
// reference collapse to four running sums per channel, [sum v, sum v*lz, sum v*ly, sum v*lx] with
// v = relu(heat) and l* = linspace(0,1,n), reduced in two deterministic stages.
#include "km_common.cuh"

namespace {

constexpr int kMaxParts = 64;

template <bool VEC>
__global__ void __launch_bounds__(256)
com3d_kernel(const float* __restrict__ heat, float* __restrict__ partials, int NK, int D, int H,
             int W) {
  const int ch = blockIdx.y;
  const long long nvox = (long long)D * H * W;
  const float* h = heat + (size_t)ch * nvox;
  float s0 = 0.f, sz = 0.f, sy = 0.f, sx = 0.f;
  if (VEC) {
    const int W4 = W / 4;
    const long long ng = nvox / 4;
    const float4* h4 = reinterpret_cast<const float4*>(h);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ng;
         g += (long long)gridDim.x * blockDim.x) {
      const int x = (int)(g % W4) * 4, y = (int)((g / W4) % H), z = (int)(g / ((long long)W4 * H));
      const float4 v4 = __ldg(h4 + g);
      const float v[4] = {fmaxf(v4.x, 0.f), fmaxf(v4.y, 0.f), fmaxf(v4.z, 0.f), fmaxf(v4.w, 0.f)};
      const float t = (v[0] + v[1]) + (v[2] + v[3]);
      s0 += t;
      sz = fmaf(t, km_linspace(0.f, 1.f, D, z), sz);
      sy = fmaf(t, km_linspace(0.f, 1.f, H, y), sy);
#pragma unroll
      for (int k = 0; k < 4; ++k) sx = fmaf(v[k], km_linspace(0.f, 1.f, W, x + k), sx);
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvox;
         i += (long long)gridDim.x * blockDim.x) {
      const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((long long)W * H));
      const float v = fmaxf(__ldg(h + i), 0.f);
      s0 += v;
      sz = fmaf(v, km_linspace(0.f, 1.f, D, z), sz);
      sy = fmaf(v, km_linspace(0.f, 1.f, H, y), sy);
      sx = fmaf(v, km_linspace(0.f, 1.f, W, x), sx);
    }
  }
  __shared__ float red[8][4];
  s0 = km_warp_sum(s0);
  sz = km_warp_sum(sz);
  sy = km_warp_sum(sy);
  sx = km_warp_sum(sx);
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = s0;
    red[threadIdx.x >> 5][1] = sz;
    red[threadIdx.x >> 5][2] = sy;
    red[threadIdx.x >> 5][3] = sx;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
    partials[((size_t)blockIdx.x * NK + ch) * 4 + threadIdx.x] = a;
  }
}

// partials (nparts, NK, 4) -> points (NK, 3), mass (NK)
__global__ void com_finalize_kernel(const float* __restrict__ com, int nparts, float* points,
                                    float* mass, int NK, int ij) {
  // one warp per keypoint channel: lanes stride over the partial slots
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= NK) return;
  double s[4] = {0, 0, 0, 0};
  for (int p = lane; p < nparts; p += 32) {
    const float4 c = *reinterpret_cast<const float4*>(com + ((size_t)p * NK + i) * 4);
    s[0] += c.x;
    s[1] += c.y;
    s[2] += c.z;
    s[3] += c.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = km_warp_sum(s[k]);
  if (lane != 0) return;
  // keymorph/layers.py:112-134: M = sum(m) + 1e-8 (fp32); c = sum(lin*m) / M; out = c*2 - 1
  const float M = (float)s[0];
  const float den = M + 1e-8f;
  const float cz = ((float)s[1] / den) * 2.f - 1.f;
  const float cy = ((float)s[2] / den) * 2.f - 1.f;
  const float cx = ((float)s[3] / den) * 2.f - 1.f;
  if (ij) {
    points[i * 3 + 0] = cz;
    points[i * 3 + 1] = cy;
    points[i * 3 + 2] = cx;
  } else {
    points[i * 3 + 0] = cx;
    points[i * 3 + 1] = cy;
    points[i * 3 + 2] = cz;
  }
  if (mass) mass[i] = M;
}

}  // namespace

extern "C" size_t km_com3d_workspace_bytes(int N, int K) {
  return (size_t)kMaxParts * N * K * 4 * sizeof(float);
}

extern "C" int km_com_finalize(const float* com, int nparts, float* points, float* mass, int N,
                               int K, km_stream_t stream) {
  KM_CHECK_ARG(com && points && nparts > 0 && N > 0 && K > 0, "km_com_finalize: bad arguments");
  const int NK = N * K;
  com_finalize_kernel<<<(NK + 7) / 8, 256, 0, km_cs(stream)>>>(com, nparts, points, mass, NK, 1);
  KM_LAUNCH_OK("com_finalize_kernel");
  return KM_OK;
}

extern "C" int km_com3d(const float* heat, float* points, float* mass, void* workspace, int N,
                        int K, int D, int H, int W, int ij, km_stream_t stream) {
  KM_CHECK_ARG(heat && points && workspace && N > 0 && K > 0 && D > 0 && H > 0 && W > 0,
               "km_com3d: bad arguments");
  const int NK = N * K;
  KM_CHECK_ARG(NK <= 65535, "km_com3d: N*K too large");
  const long long nvox = (long long)D * H * W;
  int parts = (int)((nvox + 256 * 16 - 1) / (256 * 16));
  if (parts > kMaxParts) parts = kMaxParts;
  if (parts < 1) parts = 1;
  float* partials = reinterpret_cast<float*>(workspace);
  const bool vec = (W % 4 == 0) && (((uintptr_t)heat & 15) == 0);
  if (vec)
    com3d_kernel<true><<<dim3(parts, NK), 256, 0, km_cs(stream)>>>(heat, partials, NK, D, H, W);
  else
    com3d_kernel<false><<<dim3(parts, NK), 256, 0, km_cs(stream)>>>(heat, partials, NK, D, H, W);
  KM_LAUNCH_OK("com3d_kernel");
  com_finalize_kernel<<<(NK + 7) / 8, 256, 0, km_cs(stream)>>>(partials, parts, points, mass, NK,
                                                                  ij);
  KM_LAUNCH_OK("com_finalize_kernel");
  return KM_OK;
}
