// Shared helpers for libkm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/km_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libkm_b200 is written for sm_100a (B200) only"
#endif

void km_set_error(const char* fmt, ...);

#define KM_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      km_set_error(__VA_ARGS__);       \
      return KM_EINVAL;                \
    }                                  \
  } while (0)

#define KM_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      km_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                   __LINE__);                                                         \
      return KM_ECUDA;                                                                \
    }                                                                                 \
  } while (0)

#define KM_LAUNCH_OK(name)                                                            \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      km_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));          \
      return KM_ECUDA;                                                                \
    }                                                                                 \
  } while (0)

// conv_misc.cu: per-sample bf16 weights w * scale[n] (layout 0: [tap][Cout][Cin], 1: z-folded) and the
// [N][36 border classes][Cout] bias tables of a GroupNorm folded into its consumer conv
int km_fold_gn(const float* w, const float* scale, const float* shift, void* packed, float* bias, int N, int Cout,
               int Cin, int layout, km_stream_t stream);

// cudaFuncSetAttribute (opt-in shared memory) is per device: one bit per device ordinal in a per-kernel mask,
// so that a process driving several GPUs sets it on each of them once
static inline bool km_first_use_on_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

static inline cudaStream_t km_cs(km_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of partial-sum slots written by the grid-stride reduction kernels (one per block).
// 148 SMs x 4 resident blocks: a multiple of the SM count so that the last wave is full.
#define KM_RED_BLOCKS 592
#define KM_RED_THREADS 256

// torch.linspace(start, end, n)[i] in fp32 (ATen RangeFactories: symmetric evaluation,
// step = (end - start) / (n - 1)); used by uniform_norm_grid (keymorph/utils.py:387-398)
// and CenterOfMass3d (keymorph/layers.py:99-107).
__host__ __device__ __forceinline__ float km_linspace(float start, float end, int n, int i) {
  if (n <= 1) return start;
  const float step = (end - start) / (float)(n - 1);
  const int half = n / 2;
  // torch's CPU (AVX2) and CUDA kernels both contract start + step*i into one FMA; verified
  // bit-for-bit against torch.linspace for n in 2..512 (tests/test_gpu_parity.py)
  return (i < half) ? fmaf(step, (float)i, start) : fmaf(-step, (float)(n - 1 - i), end);
}

__device__ __forceinline__ float km_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double km_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
