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

int g_tps_single_cta = 0;   // km_set_option(KM_OPT_TPS_SINGLE_CTA): 0 = one cooperative launch (default),
                            // 1 = the un-blocked one-CTA LU, 2 = the multi-launch blocked elimination (A/B)

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


// ------------------------------------------------------------------------------------------
// Blocked Gauss-Jordan elimination with partial pivoting, spread over the whole GPU.
// Rows are never moved: pivrow[k] names the pivot row of column k, elig[r] tells whether row r may
// still become a pivot.  Per panel of kNB columns:
//   tps_panel_kernel  (1 CTA / system, one THREAD per row, the row's kNB panel entries live in
//                      registers): for each column pick the largest eligible entry, broadcast
//                      the pivot row segment through smem, eliminate the column from every other
//                      row (multipliers overwrite the eliminated entries);
//   tps_update_kernel (many CTAs): applies the panel's kNB row operations to all remaining
//                      columns (and the 3 right-hand sides) of ALL rows.
// After the last panel the matrix is diagonal in the pivot order: x_k = b[pivrow[k]] / a[pivrow[k]][k].
constexpr int kNB = 16;

__global__ void __launch_bounds__(1024)
tps_panel_kernel(double* __restrict__ A, int* __restrict__ pivrow, int* __restrict__ elig,
                 int32_t* __restrict__ status, int n, int ld, int c0) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ double s_prow[kNB];
  __shared__ int s_piv;
  const int b = blockIdx.x;
  double* Ab = A + (size_t)b * n * ld;
  int* piv = pivrow + (size_t)b * n;
  int* el = elig + (size_t)b * n;
  const int r = threadIdx.x, lane = r & 31, wid = r >> 5, nwarps = blockDim.x >> 5;
  const int nb = min(kNB, n - c0);
  const bool has_row = r < n;
  double a[kNB];
  bool eligible = false;
  if (has_row) {
    eligible = c0 == 0 ? true : (el[r] != 0);
#pragma unroll
    for (int j = 0; j < kNB; ++j) a[j] = (j < nb) ? Ab[(size_t)r * ld + c0 + j] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < kNB; ++k) {
    if (k < nb) {   // uniform
      // ---- pivot search among eligible rows
      double v = (has_row && eligible) ? fabs(a[k]) : -1.0;
      int vi = r;
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (lane == 0) {
        s_val[wid] = v;
        s_idx[wid] = vi;
      }
      __syncthreads();
      if (wid == 0) {
        double w = lane < nwarps ? s_val[lane] : -1.0;
        int wi = lane < nwarps ? s_idx[lane] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, w, o);
          const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
          if (ov > w || (ov == w && oi < wi)) {
            w = ov;
            wi = oi;
          }
        }
        if (lane == 0) {
          s_piv = wi;
          if (!(w > 0.0) || !isfinite(w)) status[b] = 1;
        }
      }
      __syncthreads();
      const int pr = s_piv;
      if (r == pr) {
#pragma unroll
        for (int j = 0; j < kNB; ++j) s_prow[j] = a[j];
        eligible = false;
        piv[c0 + k] = r;
      }
      __syncthreads();
      // ---- eliminate column k from every other row; the multiplier replaces the entry
      if (has_row && r != pr) {
        const double l = a[k] / s_prow[k];
        a[k] = l;
#pragma unroll
        for (int j = 0; j < kNB; ++j)
          if (j > k) a[j] -= l * s_prow[j];
      }
    }
  }
  if (has_row) {
#pragma unroll
    for (int j = 0; j < kNB; ++j)
      if (j < nb) Ab[(size_t)r * ld + c0 + j] = a[j];
    el[r] = eligible ? 1 : 0;
  }
}

// U~[t][j] for the columns right of the panel: pivot row t as it was when it became the pivot, i.e.
// corrected by the pivots t' < t of the same panel (sequential per column, columns independent).
// Written to a side buffer because the update kernel overwrites the pivot rows themselves.
__global__ void __launch_bounds__(64)
tps_urow_kernel(const double* __restrict__ A, const int* __restrict__ pivrow, double* __restrict__ Ut,
                int n, int ld, int c0) {
  const int b = blockIdx.y;
  const double* Ab = A + (size_t)b * n * ld;
  const int* piv = pivrow + (size_t)b * n;
  const int nb = min(kNB, n - c0);
  const int j = c0 + nb + blockIdx.x * 64 + threadIdx.x;
  if (j >= ld) return;
  double u[kNB];
#pragma unroll
  for (int t = 0; t < kNB; ++t) {
    if (t < nb) {
      const int pr = piv[c0 + t];
      double v = Ab[(size_t)pr * ld + j];
#pragma unroll
      for (int t2 = 0; t2 < kNB; ++t2)
        if (t2 < t) v -= Ab[(size_t)pr * ld + c0 + t2] * u[t2];
      u[t] = v;
      Ut[((size_t)b * kNB + t) * ld + j] = v;
    }
  }
}

// grid (column tiles of 64, row tiles of 32, systems); 256 threads = 64 columns x 4 row lanes
__global__ void __launch_bounds__(256)
tps_update_kernel(double* __restrict__ A, const int* __restrict__ pivrow,
                  const double* __restrict__ Ut, int n, int ld, int c0) {
  __shared__ double s_u[kNB][64];     // pivot rows as they were when they became pivots
  __shared__ double s_l[32][kNB + 1]; // multipliers of this row tile
  __shared__ int s_pr[kNB];
  const int b = blockIdx.z;
  double* Ab = A + (size_t)b * n * ld;
  const int* piv = pivrow + (size_t)b * n;
  const int nb = min(kNB, n - c0);
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = c0 + nb + blockIdx.x * 64 + tx;   // column handled by this thread
  const bool col_ok = j < ld;
  if (threadIdx.x < kNB) s_pr[threadIdx.x] = threadIdx.x < nb ? piv[c0 + threadIdx.x] : -1;
  __syncthreads();
  for (int i = threadIdx.x; i < kNB * 64; i += 256) {
    const int t = i >> 6, c = i & 63;
    const int jj = c0 + nb + blockIdx.x * 64 + c;
    s_u[t][c] = (t < nb && jj < ld) ? Ut[((size_t)b * kNB + t) * ld + jj] : 0.0;
  }
  // multipliers of the rows of this tile
  const int r0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < 32 * kNB; i += 256) {
    const int rr = i / kNB, t = i % kNB;
    const int r = r0 + rr;
    s_l[rr][t] = (r < n && t < nb) ? Ab[(size_t)r * ld + c0 + t] : 0.0;
  }
  __syncthreads();
  if (!col_ok) return;
  for (int rr = ty; rr < 32; rr += 4) {
    const int r = r0 + rr;
    if (r >= n) break;
    double v = Ab[(size_t)r * ld + j];
#pragma unroll
    for (int t = 0; t < kNB; ++t)
      if (t < nb && s_pr[t] != r) v -= s_l[rr][t] * s_u[t][tx];
    Ab[(size_t)r * ld + j] = v;
  }
}

__global__ void tps_finish_kernel(const double* __restrict__ A, const int* __restrict__ pivrow,
                                  float* __restrict__ theta, int n, int ld) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const int k = i / 3, d = i % 3;
  const double* row = A + ((size_t)b * n + pivrow[(size_t)b * n + k]) * ld;
  theta[((size_t)b * n + k) * 3 + d] = (float)(row[n + d] / row[k]);
}


// ------------------------------------------------------------------------------------------
// The same blocked Gauss-Jordan elimination as ONE cooperative launch (default).  The multi-launch version
// above spends its 1.35 ms (K = 512) in ~100 dependent launches of microsecond kernels; here a group of
// `cps` co-resident CTAs per system walks the panels with a group barrier (one global counter per system)
// between the two phases of a panel:
//   phase A (CTA 0 of the group): the panel factorisation of tps_panel_kernel -- one thread per row, the
//            row's kNB panel entries in registers -- with TWO block barriers per pivot step: every warp
//            reduces the per-warp candidates redundantly, so the pivot index needs no broadcast round;
//   phase B (all CTAs): tiles of 64 columns x kRowTile rows of the trailing matrix (and the right-hand
//            sides): the 16 pivot rows "as they were when they became pivots" are re-derived per column
//            in registers (the recurrence of tps_urow_kernel), then the rank-16 update of the tile.
// Assembly and the final division run in the same launch.  The arithmetic and its order are those of the
// multi-launch kernels, so the two paths produce identical bits (tests compare them).
constexpr int kRowTile = 128;
constexpr int kGjThreads = 1024;   // largest block: one thread per matrix row, n = K + 4 <= 1024

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all CTAs of a system group arrive; `target` = cps * (number of barriers passed so far + 1)
__device__ __forceinline__ void group_barrier(unsigned int* cnt, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(cnt, 1u);
    while (ld_acquire_u32(cnt) < target) {
    }
  }
  __syncthreads();
}

template <int THREADS>   // 640 (n <= 640: 102 registers per thread, no spills) or 1024
__global__ void __launch_bounds__(THREADS, 1)
tps_gj_coop_kernel(const float* __restrict__ c_src, const float* __restrict__ c_dst, const float* __restrict__ lmbda,
                   const float* __restrict__ w, double* __restrict__ A, int* __restrict__ pivrow,
                   unsigned int* __restrict__ bar, float* __restrict__ theta, int32_t* __restrict__ status, int K,
                   int cps) {
  __shared__ double s_val[2][32];
  __shared__ int s_idx[2][32];
  __shared__ double s_prow[2][kNB];
  __shared__ double s_u[kNB][64];
  __shared__ double s_l[kRowTile][kNB + 1];
  __shared__ double s_lp[kNB][kNB];   // multipliers of the pivot rows themselves (urow recurrence)
  __shared__ int s_pr[kNB];
  const int b = blockIdx.x / cps, cta = blockIdx.x - b * cps;
  const int n = K + 4, ld = n + 3;
  double* Ab = A + (size_t)b * n * ld;
  int* piv = pivrow + (size_t)b * n;
  unsigned int* cnt = bar + b;
  unsigned int nbar = 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = THREADS >> 5;

  // ---- assembly (tps_assemble_kernel), spread over the group
  {
    const float* cs = c_src + (size_t)b * K * 3;
    const float* cd = c_dst + (size_t)b * K * 3;
    const double lam = (double)lmbda[b];
    const int total = n * ld;
    for (int idx = cta * THREADS + tid; idx < total; idx += cps * THREADS) {
      const int i = idx / ld, j = idx - i * ld;
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
          v = (j == K) ? 1.0 : (double)cs[i * 3 + (j - K - 1)];
        } else {
          v = (double)cd[i * 3 + (j - n)];
        }
      } else if (j < K) {
        v = (i == K) ? 1.0 : (double)cs[j * 3 + (i - K - 1)];
      }
      Ab[idx] = v;
    }
  }
  group_barrier(cnt, (++nbar) * cps);

  bool eligible = true;          // CTA 0: row tid may still become a pivot
  int sing = 0;
  for (int c0 = 0; c0 < n; c0 += kNB) {
    const int nb = min(kNB, n - c0);
    // ---------------- phase A: panel factorisation by CTA 0 (one thread per row)
    if (cta == 0) {
      const int r = tid;
      const bool has_row = r < n;
      double a[kNB];
#pragma unroll
      for (int j = 0; j < kNB; ++j) a[j] = (has_row && j < nb) ? Ab[(size_t)r * ld + c0 + j] : 0.0;
#pragma unroll
      for (int k = 0; k < kNB; ++k) {
        if (k < nb) {   // uniform
          const int pb = k & 1;
          double v = (has_row && eligible) ? fabs(a[k]) : -1.0;
          int vi = r;
          for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
            if (ov > v || (ov == v && oi < vi)) {
              v = ov;
              vi = oi;
            }
          }
          if (lane == 0) {
            s_val[pb][wid] = v;
            s_idx[pb][wid] = vi;
          }
          __syncthreads();
          // every warp reduces the per-warp candidates itself: no broadcast round for the pivot index
          double wv = lane < nwarps ? s_val[pb][lane] : -1.0;
          int wi = lane < nwarps ? s_idx[pb][lane] : 0x7fffffff;
          for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, wv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (ov > wv || (ov == wv && oi < wi)) {
              wv = ov;
              wi = oi;
            }
          }
          const int pr = wi;
          if (tid == 0 && (!(wv > 0.0) || !isfinite(wv))) sing = 1;
          if (r == pr) {
#pragma unroll
            for (int j = 0; j < kNB; ++j) s_prow[pb][j] = a[j];
            eligible = false;
            piv[c0 + k] = r;
          }
          __syncthreads();
          if (has_row && r != pr) {
            const double l = a[k] / s_prow[pb][k];
            a[k] = l;
#pragma unroll
            for (int j = 0; j < kNB; ++j)
              if (j > k) a[j] -= l * s_prow[pb][j];
          }
        }
      }
      if (has_row) {
#pragma unroll
        for (int j = 0; j < kNB; ++j)
          if (j < nb) Ab[(size_t)r * ld + c0 + j] = a[j];
      }
    }
    group_barrier(cnt, (++nbar) * cps);

    // ---------------- phase B: trailing update, tiles of 64 columns x kRowTile rows
    const int rest = ld - c0 - nb;
    if (rest > 0) {
      const int nct = (rest + 63) / 64, nrt = (n + kRowTile - 1) / kRowTile;
      if (tid < kNB) s_pr[tid] = tid < nb ? piv[c0 + tid] : -1;
      __syncthreads();
      // multipliers of the pivot rows among themselves: L[t][t2] = A[piv t][c0 + t2], t2 < t
      if (tid < kNB * kNB) {
        const int t = tid / kNB, t2 = tid - t * kNB;
        s_lp[t][t2] = (t < nb && t2 < t) ? Ab[(size_t)s_pr[t] * ld + c0 + t2] : 0.0;
      }
      for (int item = cta; item < nct * nrt; item += cps) {
        const int ct = item % nct, rt = item / nct;
        const int j0 = c0 + nb + ct * 64, r0 = rt * kRowTile;
        __syncthreads();   // s_u / s_l of the previous item are free (and s_lp / s_pr are visible)
        if (tid < 64) {
          const int j = j0 + tid;
          double u[kNB];
#pragma unroll
          for (int t = 0; t < kNB; ++t) u[t] = (t < nb && j < ld) ? Ab[(size_t)s_pr[t] * ld + j] : 0.0;
#pragma unroll
          for (int t = 0; t < kNB; ++t) {
            double v = u[t];
#pragma unroll
            for (int t2 = 0; t2 < kNB; ++t2)
              if (t2 < t) v -= s_lp[t][t2] * u[t2];
            u[t] = v;
            s_u[t][tid] = v;
          }
        }
        for (int i = tid; i < kRowTile * kNB; i += THREADS) {
          const int rr = i / kNB, t = i - rr * kNB;
          const int r = r0 + rr;
          s_l[rr][t] = (r < n && t < nb) ? Ab[(size_t)r * ld + c0 + t] : 0.0;
        }
        __syncthreads();
        const int tx = tid & 63, ty = tid >> 6;     // 64 columns x 16 row lanes
        const int j = j0 + tx;
        if (j < ld) {
          for (int rr = ty; rr < kRowTile; rr += THREADS / 64) {
            const int r = r0 + rr;
            if (r >= n) break;
            double v = Ab[(size_t)r * ld + j];
#pragma unroll
            for (int t = 0; t < kNB; ++t)
              if (t < nb && s_pr[t] != r) v -= s_l[rr][t] * s_u[t][tx];
            Ab[(size_t)r * ld + j] = v;
          }
        }
      }
    }
    group_barrier(cnt, (++nbar) * cps);
  }
  // ---- x_k = b[pivrow[k]] / a[pivrow[k]][k]
  for (int i = cta * THREADS + tid; i < n * 3; i += cps * THREADS) {
    const int k = i / 3, d = i - k * 3;
    const double* row = Ab + (size_t)piv[k] * ld;
    theta[((size_t)b * n + k) * 3 + d] = (float)(row[n + d] / row[k]);
  }
  if (cta == 0 && tid == 0) status[b] = sing;
}

}  // namespace

extern "C" size_t km_tps_fit_workspace_bytes(int N, int K) {
  const size_t n = (size_t)K + 4;
  // augmented matrix (fp64) + pivot rows + eligibility flags + one barrier counter per system
  return (size_t)N * (n + kNB) * (n + 3) * sizeof(double) + 2 * (size_t)N * n * sizeof(int) +
         (size_t)N * sizeof(unsigned int) + 64;
}

extern "C" int km_tps_fit(const float* c_src, const float* c_dst, const float* lmbda,
                          const float* w, float* theta, int32_t* status, void* workspace, int N,
                          int K, km_stream_t stream) {
  KM_CHECK_ARG(c_src && c_dst && lmbda && theta && status && workspace && N > 0 && K > 0,
               "km_tps_fit: bad arguments");
  const int n = K + 4, ld = n + 3;
  cudaStream_t st = km_cs(stream);
  double* A = reinterpret_cast<double*>(workspace);
  if (n <= kGjThreads && g_tps_single_cta == 0) {
    // one cooperative launch: assembly, blocked Gauss-Jordan and the final division (see tps_gj_coop_kernel)
    double* Ut = A + (size_t)N * n * ld;
    int* pivrow = reinterpret_cast<int*>(Ut + (size_t)N * kNB * ld);
    unsigned int* bar = reinterpret_cast<unsigned int*>(pivrow + 2 * (size_t)N * n);
    KM_CUDA_OK(cudaMemsetAsync(bar, 0, (size_t)N * sizeof(unsigned int), st));
    int dev = 0, nsm = 148;
    KM_CUDA_OK(cudaGetDevice(&dev));
    KM_CUDA_OK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    // every CTA of a launch must be resident (the group barrier spins): one 1024-thread CTA per SM
    const int chunk = N < nsm ? N : nsm;
    for (int b0 = 0; b0 < N; b0 += chunk) {
      const int nb = N - b0 < chunk ? N - b0 : chunk;
      int cps = nsm / nb;
      if (cps > 24) cps = 24;
      const float* cs = c_src + (size_t)b0 * K * 3;
      const float* cd = c_dst + (size_t)b0 * K * 3;
      const float* lm = lmbda + b0;
      const float* wp = w ? w + (size_t)b0 * K : nullptr;
      double* Ab = A + (size_t)b0 * n * ld;
      int* pv = pivrow + (size_t)b0 * n;
      unsigned int* br = bar + b0;
      float* th = theta + (size_t)b0 * n * 3;
      int32_t* stt = status + b0;
      int Kk = K;
      void* args[] = {&cs, &cd, &lm, &wp, &Ab, &pv, &br, &th, &stt, &Kk, &cps};
      if (n <= 640)
        KM_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(tps_gj_coop_kernel<640>), dim3(nb * cps),
                                               dim3(640), args, 0, st));
      else
        KM_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(tps_gj_coop_kernel<1024>),
                                               dim3(nb * cps), dim3(1024), args, 0, st));
    }
    return KM_OK;
  }
  const long long total = (long long)n * ld;
  if (n <= 1024 && g_tps_single_cta != 1) {
    double* Ut = A + (size_t)N * n * ld;
    int* pivrow = reinterpret_cast<int*>(Ut + (size_t)N * kNB * ld);
    int* elig = pivrow + (size_t)N * n;
    KM_CUDA_OK(cudaMemsetAsync(status, 0, (size_t)N * sizeof(int32_t), st));
    const int threads = (n + 31) / 32 * 32;
    for (int c0 = 0; c0 < n; c0 += kNB) {
      tps_panel_kernel<<<N, threads, 0, st>>>(A, pivrow, elig, status, n, ld, c0);
      const int nb = n - c0 < kNB ? n - c0 : kNB;
      const int rest = ld - c0 - nb;   // columns right of the panel (incl. the 3 right-hand sides)
      tps_urow_kernel<<<dim3((rest + 63) / 64, N), 64, 0, st>>>(A, pivrow, Ut, n, ld, c0);
      tps_update_kernel<<<dim3((rest + 63) / 64, (n + 31) / 32, N), 256, 0, st>>>(A, pivrow, Ut, n, ld,
                                                                                 c0);
    }
    KM_LAUNCH_OK("tps_panel/update_kernel");
    tps_finish_kernel<<<dim3((n * 3 + 127) / 128, N), 128, 0, st>>>(A, pivrow, theta, n, ld);
    KM_LAUNCH_OK("tps_finish_kernel");
    return KM_OK;
  }
  // very large systems: one CTA per system, matrix streamed from L2
  const size_t smem = (size_t)4 * n * sizeof(double) + (size_t)n * sizeof(int);
  KM_CHECK_ARG(smem <= 200 * 1024, "km_tps_fit: K=%d too large", K);
  if (smem > 48 * 1024) {
    KM_CUDA_OK(cudaFuncSetAttribute(tps_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  }
  tps_solve_kernel<<<N, 1024, smem, st>>>(A, theta, status, K);
  KM_LAUNCH_OK("tps_solve_kernel");
  return KM_OK;
}

void km_tps_set_single_cta(int v) { g_tps_single_cta = (v == 1 || v == 2) ? v : 0; }
