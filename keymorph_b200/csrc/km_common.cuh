// Shared helpers for libkm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
// the same with the packed weights restricted to input channels [c0, c0 + Cp) (the bias tables cover all Cin)
int km_fold_gn_part(const float* w, const float* scale, const float* shift, void* packed, float* bias, int N,
                    int Cout, int Cin, int c0, int Cp, int layout, km_stream_t stream);

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

// ---- 16-bit activation / weight element type of the backbone ---------------------------------------
// The tcgen05 kind::f16 MMA takes fp16 or bf16 operands at the same rate; which one the backbone stores is a
// runtime option (km_set_option(KM_OPT_OPERAND_FP16, ...)): fp16 is the reference's own AMP dtype
// (keymorph/model.py:175-177) and carries 3 more mantissa bits, bf16 has fp32's exponent range.  Kernels are
// templated on it (F16 = true: __half), so the choice costs nothing in the epilogues.
int km_operand_fp16();    // conv_misc.cu: current value of the option
// element type of the TMA tensor maps over activation / weight tensors (tile mode copies bits either way)
#define KM_TMAP_16 (km_operand_fp16() ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)

template <bool F16>
__device__ __forceinline__ uint32_t km_pack2(float lo, float hi) {
  if (F16) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  } else {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
}
template <bool F16>
__device__ __forceinline__ float2 km_unpack2(uint32_t u) {
  if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
template <bool F16>
__device__ __forceinline__ float km_to_float(uint16_t u) {
  if (F16) return __half2float(*reinterpret_cast<const __half*>(&u));
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&u));
}
template <bool F16>
__device__ __forceinline__ uint16_t km_from_float(float f) {
  if (F16) {
    const __half h = __float2half_rn(f);
    return *reinterpret_cast<const uint16_t*>(&h);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(f);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
}
// elementwise max of two packed pairs (max commutes with the rounding, so pooling stored values is exact)
template <bool F16>
__device__ __forceinline__ uint32_t km_max2(uint32_t a, uint32_t b) {
  if (F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                     *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
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
