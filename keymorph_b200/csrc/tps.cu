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
// The blocked Gauss-Jordan elimination as ONE cooperative launch (default).  The multi-launch version above
// spends its 1.35 ms (K = 512) in ~100 dependent launches of microsecond kernels whose critical path is the
// 516 sequential pivot steps, each a block-wide arg-max over 17 warps with three block barriers.  Here a
// group of `cps` co-resident 256-thread CTAs per system walks the panels with a group barrier (one global
// counter per system) between the two phases of a panel:
//   phase A (PT = 128 or 256 threads of CTA 0): panel factorisation with RPT rows per thread, the rows'
//            kNB panel entries in registers.  Pivot search = thread-local maximum, then ONE warp redux on
//            the high word of |a| (a pivot within 2^-20 of the largest candidate is as good as the largest;
//            ties go to the lowest row), then the 4..8 per-warp candidates through shared memory: two
//            named barriers per pivot step among PT threads instead of three among 544;
//   phase B (all CTAs): column tiles of the trailing matrix (and the right-hand sides), each owned by one
//            CTA for all rows: the 16 pivot rows "as they were when they became pivots" are re-derived per
//            column in registers (the recurrence of tps_urow_kernel) BEFORE the owner overwrites them,
//            then the rank-16 update of the tile, 64 rows at a time.
// Assembly and the final division run in the same launch.
constexpr int kGjThreads = 256;
constexpr int kGjColTile = 8;      // columns of the trailing matrix owned by one CTA per panel (65 tiles at K = 512)
constexpr int kGjMaxRows = 1024;   // n = K + 4 <= PT * RPT

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
template <int PT>
__device__ __forceinline__ void panel_barrier() {   // named barrier 1 among the PT panel threads
  asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");
}

template <int PT, int RPT>
__global__ void __launch_bounds__(kGjThreads, 1)
tps_gj_coop_kernel(const float* __restrict__ c_src, const float* __restrict__ c_dst, const float* __restrict__ lmbda,
                   const float* __restrict__ w, double* __restrict__ A, int* __restrict__ pivrow,
                   unsigned int* __restrict__ bar, float* __restrict__ theta, int32_t* __restrict__ status, int K,
                   int cps, int dbg) {
  __shared__ unsigned int s_key[2][PT / 32];
  __shared__ int s_idx[2][PT / 32];
  __shared__ double s_prow[2][kNB];
  __shared__ double s_pinv[2];
  __shared__ double s_u[kNB][kGjColTile];
  __shared__ double s_lp[kNB][kNB];   // multipliers of the pivot rows among themselves (urow recurrence)
  __shared__ int s_pr[kNB];
  const int b = blockIdx.x / cps, cta = blockIdx.x - b * cps;
  const int n = K + 4, ld = n + 3;
  double* Ab = A + (size_t)b * n * ld;
  int* piv = pivrow + (size_t)b * n;
  unsigned int* cnt = bar + b;
  unsigned int nbar = 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  extern __shared__ double s_mul[];   // [n][kNB + 1]: multipliers of every row for the current panel

  // ---- assembly (tps_assemble_kernel), spread over the group
  {
    const float* cs = c_src + (size_t)b * K * 3;
    const float* cd = c_dst + (size_t)b * K * 3;
    const double lam = (double)lmbda[b];
    const int total = n * ld;
    for (int idx = cta * kGjThreads + tid; idx < total; idx += cps * kGjThreads) {
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

  unsigned int elig = (1u << RPT) - 1u;   // panel threads of CTA 0: bit i = row tid + i * PT may still be a pivot
  int sing = 0;
  long long t_a = 0, t_b = 0, t_bar = 0, t_mark = clock64();   // dbg: cycles of CTA 0 per phase
  for (int c0 = 0; c0 < n; c0 += kNB) {
    const int nb = min(kNB, n - c0);
    // ---------------- phase A: panel factorisation by PT threads of CTA 0, RPT rows per thread
    if (cta == 0 && tid < PT) {
      double a[RPT][kNB];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = tid + i * PT;
#pragma unroll
        for (int j = 0; j < kNB; ++j) a[i][j] = (r < n && j < nb) ? Ab[(size_t)r * ld + c0 + j] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < kNB; ++k) {
        if (k < nb) {   // uniform
          const int pb = k & 1;
          // candidates are compared on the high word of |a| (sign cleared): FP64 issue is the scarce resource
          // of this chip (measured ~16 cycles per warp instruction), integer compares are free next to it
          unsigned int key = 0u;
          int bi = 0x7fffffff;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = tid + i * PT;
            unsigned int ki = (unsigned int)__double2hiint(a[i][k]) & 0x7fffffffu;
            if (ki == 0u && __double2loint(a[i][k]) != 0) ki = 1u;        // tiny but non-zero stays eligible
            if (ki >= 0x7ff00000u) ki = 0x7ff00000u;                       // Inf / NaN: surfaces as "singular"
            if (r < n && ((elig >> i) & 1u) && ki > key) {
              key = ki;
              bi = r;
            }
          }
          const unsigned int wkey = __reduce_max_sync(0xffffffffu, key);
          const int widx = __reduce_min_sync(0xffffffffu, (key == wkey && key != 0u) ? bi : 0x7fffffff);
          if (lane == 0) {
            s_key[pb][wid] = wkey;
            s_idx[pb][wid] = widx;
          }
          panel_barrier<PT>();
          unsigned int gk = 0u;
          int pr = 0x7fffffff;
#pragma unroll
          for (int q = 0; q < PT / 32; ++q) {
            const unsigned int kq = s_key[pb][q];
            const int iq = s_idx[pb][q];
            if (kq > gk || (kq == gk && iq < pr)) {
              gk = kq;
              pr = iq;
            }
          }
          // a vanishing or non-finite pivot marks the system singular
          if (tid == 0 && (gk == 0u || gk >= 0x7ff00000u || pr == 0x7fffffff)) sing = 1;
          if (pr == 0x7fffffff) pr = 0;   // nothing eligible (cannot happen for k < n): keep going, flagged
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            if (tid + i * PT == pr) {
#pragma unroll
              for (int j = 0; j < kNB; ++j) s_prow[pb][j] = a[i][j];
              // one division per step, by the owner of the pivot row (an fp32-seeded Newton reciprocal was
              // measured SLOWER: +0.4 ms over the 516 steps, profiles/r02_tps_fit_cooperative_phases.log)
              s_pinv[pb] = 1.0 / a[i][k];
              elig &= ~(1u << i);
              piv[c0 + k] = pr;
            }
          }
          panel_barrier<PT>();
          const double pinv = s_pinv[pb];
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = tid + i * PT;
            if (r < n && r != pr) {
              const double l = a[i][k] * pinv;
              a[i][k] = l;
#pragma unroll
              for (int j = 0; j < kNB; ++j)
                if (j > k) a[i][j] -= l * s_prow[pb][j];
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = tid + i * PT;
        if (r < n) {
#pragma unroll
          for (int j = 0; j < kNB; ++j)
            if (j < nb) Ab[(size_t)r * ld + c0 + j] = a[i][j];
        }
      }
    }
    if (dbg) { const long long t = clock64(); t_a += t - t_mark; t_mark = t; }
    group_barrier(cnt, (++nbar) * cps);
    if (dbg) { const long long t = clock64(); t_bar += t - t_mark; t_mark = t; }

    // ---------------- phase B: trailing update.  A column tile (kGjColTile columns, all rows) belongs to ONE
    // CTA: the pivot rows "as they were" (u) are derived from the tile's own columns before that CTA
    // overwrites them, and no other CTA touches these columns -- no side buffer, no extra barrier.
    const int rest = ld - c0 - nb;
    if (rest > 0) {
      const int nct = (rest + kGjColTile - 1) / kGjColTile;
      if (tid < kNB) s_pr[tid] = tid < nb ? piv[c0 + tid] : -1;
      __syncthreads();
      // multipliers of the pivot rows among themselves: L[t][t2] = A[piv t][c0 + t2], t2 < t
      {
        const int t = tid / kNB, t2 = tid - t * kNB;    // 256 threads = kNB x kNB
        s_lp[t][t2] = (t < nb && t2 < t) ? Ab[(size_t)s_pr[t] * ld + c0 + t2] : 0.0;
      }
      // multipliers of every row for this panel: n x kNB doubles, staged once (each thread's loads are
      // independent: one L2 round trip), pitch kNB + 1 against bank conflicts
      const bool mine = cta < nct;
      if (mine) {
        for (int i = tid; i < n * kNB; i += kGjThreads) {
          const int r = i / kNB, t = i - r * kNB;
          s_mul[r * (kNB + 1) + t] = t < nb ? Ab[(size_t)r * ld + c0 + t] : 0.0;
        }
      }
      for (int ct = cta; ct < nct; ct += cps) {
        const int j0 = c0 + nb + ct * kGjColTile;
        __syncthreads();   // s_u of the previous tile is free (and s_lp / s_pr / s_mul are visible)
        if (tid < kGjColTile) {
          const int j = j0 + tid;
          double u[kNB];
#pragma unroll
          for (int t = 0; t < kNB; ++t) u[t] = (t < nb && j < ld) ? Ab[(size_t)s_pr[t] * ld + j] : 0.0;
          // forward substitution in axpy form: the critical path is kNB dependent FMAs, not kNB^2 / 2
#pragma unroll
          for (int t = 0; t < kNB; ++t) {
#pragma unroll
            for (int t3 = 0; t3 < kNB; ++t3)
              if (t3 > t) u[t3] -= s_lp[t3][t] * u[t];
            s_u[t][tid] = u[t];
          }
        }
        __syncthreads();   // s_u is visible
        // rank-16 update of all rows of the tile: a thread owns column tx of the rows ty, ty + 32, ...
        const int tx = tid % kGjColTile, ty = tid / kGjColTile;
        const int j = j0 + tx;
        if (j < ld) {
          double uc[kNB];
#pragma unroll
          for (int t = 0; t < kNB; ++t) uc[t] = s_u[t][tx];
          constexpr int kLanes = kGjThreads / kGjColTile;
          constexpr int kBatch = 6;               // rows in flight per thread: independent L2 loads
          for (int rb = ty; rb < n; rb += kLanes * kBatch) {
            double v[kBatch];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
              const int r = rb + q * kLanes;
              v[q] = r < n ? Ab[(size_t)r * ld + j] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
              const int r = rb + q * kLanes;
              if (r < n) {
                const double* l = s_mul + r * (kNB + 1);
                double acc = v[q];
#pragma unroll
                for (int t = 0; t < kNB; ++t)
                  if (s_pr[t] != r) acc -= l[t] * uc[t];      // l[t] = 0 for t >= nb
                Ab[(size_t)r * ld + j] = acc;
              }
            }
          }
        }
      }
    }
    if (dbg) { const long long t = clock64(); t_b += t - t_mark; t_mark = t; }
    group_barrier(cnt, (++nbar) * cps);
    if (dbg) { const long long t = clock64(); t_bar += t - t_mark; t_mark = t; }
  }
  if (dbg && blockIdx.x == 0 && tid == 0)
    printf("tps_gj_coop: n=%d cps=%d cycles of CTA 0: panel %lld, update %lld, barriers %lld\n", n, cps, t_a, t_b, t_bar);
  // ---- x_k = b[pivrow[k]] / a[pivrow[k]][k]
  for (int i = cta * kGjThreads + tid; i < n * 3; i += cps * kGjThreads) {
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
  if (n <= kGjMaxRows && (g_tps_single_cta == 0 || g_tps_single_cta == 3)) {
    // one cooperative launch: assembly, blocked Gauss-Jordan and the final division (see tps_gj_coop_kernel)
    double* Ut = A + (size_t)N * n * ld;
    int* pivrow = reinterpret_cast<int*>(Ut + (size_t)N * kNB * ld);
    unsigned int* bar = reinterpret_cast<unsigned int*>(pivrow + 2 * (size_t)N * n);
    KM_CUDA_OK(cudaMemsetAsync(bar, 0, (size_t)N * sizeof(unsigned int), st));
    int dev = 0, nsm = 148;
    KM_CUDA_OK(cudaGetDevice(&dev));
    KM_CUDA_OK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    // every CTA of a launch must be resident (the group barrier spins): at most one CTA per SM
    const int chunk = N < nsm ? N : nsm;
    for (int b0 = 0; b0 < N; b0 += chunk) {
      const int nbs = N - b0 < chunk ? N - b0 : chunk;
      int cps = nsm / nbs;
      const int max_tiles = (ld - kNB + kGjColTile - 1) / kGjColTile;   // tiles of the first (widest) panel
      if (cps > max_tiles) cps = max_tiles;
      if (cps < 1) cps = 1;
      const float* cs = c_src + (size_t)b0 * K * 3;
      const float* cd = c_dst + (size_t)b0 * K * 3;
      const float* lm = lmbda + b0;
      const float* wp = w ? w + (size_t)b0 * K : nullptr;
      double* Ab = A + (size_t)b0 * n * ld;
      int* pv = pivrow + (size_t)b0 * n;
      unsigned int* br = bar + b0;
      float* th = theta + (size_t)b0 * n * 3;
      int32_t* stt = status + b0;
      int Kk = K, dbg = g_tps_single_cta == 3;
      void* args[] = {&cs, &cd, &lm, &wp, &Ab, &pv, &br, &th, &stt, &Kk, &cps, &dbg};
      const void* fn = n <= 640 ? reinterpret_cast<const void*>(tps_gj_coop_kernel<128, 5>)
                                : reinterpret_cast<const void*>(tps_gj_coop_kernel<256, 4>);
      const size_t smem = (size_t)n * (kNB + 1) * sizeof(double);     // <= 139 KB at n = 1024
      static unsigned long long attr_set = 0;
      if (km_first_use_on_device(&attr_set)) {
        KM_CUDA_OK(cudaFuncSetAttribute(tps_gj_coop_kernel<128, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        KM_CUDA_OK(cudaFuncSetAttribute(tps_gj_coop_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      }
      KM_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(nbs * cps), dim3(kGjThreads), args, smem, st));
    }
    return KM_OK;
  }
  const long long total = (long long)n * ld;
  int bx = (int)((total + 255) / 256);
  if (bx > 1184) bx = 1184;
  tps_assemble_kernel<<<dim3(bx, N), 256, 0, st>>>(c_src, c_dst, lmbda, w, A, K);
  KM_LAUNCH_OK("tps_assemble_kernel");
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

void km_tps_set_single_cta(int v) { g_tps_single_cta = (v >= 1 && v <= 3) ? v : 0; }   // 3: 0 + phase timing printed
