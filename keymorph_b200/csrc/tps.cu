// Thin-plate-spline fit on the device (keymorph/keypoint_aligners.py:276-363).
//
// The reference assembles A = [[U + lambda*I, P], [P^T, 0]] on the HOST (P, v and A are created
// without device=), copies U over, calls LAPACK gesv three times (x, y, z targets) and copies the
// result back.  Here the system is assembled by a grid-wide kernel directly in its final place
// (augmented with the three right-hand sides) and solved once by a partially pivoted LU in fp64;
// rows are never moved, a permutation vector in shared memory names the pivot rows.
//   U_ij = r^2 log(r + 1e-6),  r = sqrt(|c_i - c_j|^2 + 1e-6)          (:322-339)
//   weighted: K = U + lambda / (diag_embed(w) + 1e-6)  applied to the DENSE matrix, i.e. the
//   off-diagonal entries receive lambda * 1e6 exactly like the reference (:298-302).
#include "km_common.cuh"

namespace {

__device__ __forceinline__ double tps_u64(double d2) {
  const double r = sqrt(d2 + 1e-6);
  return (r * r) * log(r + 1e-6);
}

// A: (N, n, ld) row-major fp64, n = K + 4, ld = n + 3 (three RHS columns appended)
__global__ void __launch_bounds__(256)
tps_assemble_kernel(const float* __restrict__ c_src, const float* __restrict__ c_dst,
                    const float* __restrict__ lmbda, const float* __restrict__ w,
                    double* __restrict__ A, int K) {
  const int b = blockIdx.y;
  const int n = K + 4, ld = n + 3;
  const float* cs = c_src + (size_t)b * K * 3;
  const float* cd = c_dst + (size_t)b * K * 3;
  const double lam = (double)lmbda[b];
  double* Ab = A + (size_t)b * n * ld;
  const long long total = (long long)n * ld;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / ld), j = (int)(idx % ld);
    double v = 0.0;
    if (i < K) {
      if (j < K) {
        const double dz = (double)cs[i * 3] - (double)cs[j * 3];
        const double dy = (double)cs[i * 3 + 1] - (double)cs[j * 3 + 1];
        const double dx = (double)cs[i * 3 + 2] - (double)cs[j * 3 + 2];
        v = tps_u64(dz * dz + dy * dy + dx * dx);
        if (w) {
          const double wij = (i == j) ? (double)w[(size_t)b * K + i] : 0.0;
          v += lam / (wij + 1e-6);
        } else if (i == j) {
          v += lam;
        }
      } else if (j < n) {
        v = (j == K) ? 1.0 : (double)cs[i * 3 + (j - K - 1)];   // P = [1, c]
      } else {
        v = (double)cd[i * 3 + (j - n)];                        // targets
      }
    } else {
      if (j < K) v = (i == K) ? 1.0 : (double)cs[j * 3 + (i - K - 1)];  // P^T
      // zero block and zero right-hand side otherwise
    }
    Ab[idx] = v;
  }
}

// one CTA per system
__global__ void __launch_bounds__(1024)
tps_solve_kernel(double* __restrict__ A, float* __restrict__ theta, int32_t* __restrict__ status,
                 int K) {
  extern __shared__ double sdyn[];
  const int n = K + 4, ld = n + 3;
  double* cand = sdyn;                                  // [n] |A[perm[i]][k]| for the next column
  double* xs = sdyn + n;                                // [3][n] solution
  int* perm = reinterpret_cast<int*>(sdyn + 4 * (size_t)n);  // [n]
  __shared__ int s_prow;
  __shared__ int s_sing;
  const int b = blockIdx.x;
  double* Ab = A + (size_t)b * n * ld;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;

  for (int i = tid; i < n; i += blockDim.x) {
    perm[i] = i;
    cand[i] = fabs(Ab[(size_t)i * ld]);
  }
  if (tid == 0) s_sing = 0;

  for (int k = 0; k < n; ++k) {
    __syncthreads();  // cand[] of column k complete
    if (warp == 0) {
      double best = -1.0;  // NaNs never compare greater: they are ignored here and caught below
      int bi = k;
      for (int i = k + lane; i < n; i += 32) {
        const double v = cand[i];
        if (v > best) {
          best = v;
          bi = i;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if ((ob > best) || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (lane == 0) {
        const int t = perm[k];
        perm[k] = perm[bi];
        perm[bi] = t;
        s_prow = perm[k];
        if (!(best > 0.0) || !isfinite(best)) s_sing = 1;
      }
    }
    __syncthreads();
    const int prow = s_prow;
    const double* prp = Ab + (size_t)prow * ld;
    const double pinv = 1.0 / prp[k];
    for (int i = k + 1 + warp; i < n; i += nwarps) {
      double* rp = Ab + (size_t)perm[i] * ld;
      const double l = rp[k] * pinv;
      for (int j = k + 1 + lane; j < ld; j += 32) {
        const double v = rp[j] - l * prp[j];
        rp[j] = v;
        if (j == k + 1) cand[i] = fabs(v);
      }
    }
  }
  __syncthreads();
  // back substitution, one warp per right-hand side
  if (warp < 3) {
    double* x = xs + (size_t)warp * n;
    for (int k = n - 1; k >= 0; --k) {
      const double* rp = Ab + (size_t)perm[k] * ld;
      double acc = 0.0;
      for (int j = k + 1 + lane; j < n; j += 32) acc += rp[j] * x[j];
      acc = km_warp_sum(acc);
      if (lane == 0) x[k] = (rp[n + warp] - acc) / rp[k];
      __syncwarp();
    }
  }
  __syncthreads();
  float* th = theta + (size_t)b * n * 3;
  for (int i = tid; i < n * 3; i += blockDim.x) th[i] = (float)xs[(size_t)(i % 3) * n + i / 3];
  if (tid == 0) status[b] = s_sing;
}

}  // namespace

extern "C" size_t km_tps_fit_workspace_bytes(int N, int K) {
  const size_t n = (size_t)K + 4;
  return (size_t)N * n * (n + 3) * sizeof(double);
}

extern "C" int km_tps_fit(const float* c_src, const float* c_dst, const float* lmbda,
                          const float* w, float* theta, int32_t* status, void* workspace, int N,
                          int K, km_stream_t stream) {
  KM_CHECK_ARG(c_src && c_dst && lmbda && theta && status && workspace && N > 0 && K > 0,
               "km_tps_fit: bad arguments");
  const int n = K + 4;
  const size_t smem = (size_t)4 * n * sizeof(double) + (size_t)n * sizeof(int);
  KM_CHECK_ARG(smem <= 200 * 1024, "km_tps_fit: K=%d too large", K);
  double* A = reinterpret_cast<double*>(workspace);
  const long long total = (long long)n * (n + 3);
  int bx = (int)((total + 255) / 256);
  if (bx > 1184) bx = 1184;
  tps_assemble_kernel<<<dim3(bx, N), 256, 0, km_cs(stream)>>>(c_src, c_dst, lmbda, w, A, K);
  KM_LAUNCH_OK("tps_assemble_kernel");
  if (smem > 48 * 1024) {
    KM_CUDA_OK(cudaFuncSetAttribute(tps_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  }
  tps_solve_kernel<<<N, 1024, smem, km_cs(stream)>>>(A, theta, status, K);
  KM_LAUNCH_OK("tps_solve_kernel");
  return KM_OK;
}
