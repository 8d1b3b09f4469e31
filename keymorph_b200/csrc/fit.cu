// Closed-form keypoint aligners: one warp per pair.
//   affine: keymorph/keypoint_aligners.py:76-114   A = Y W X^T (X W X^T)^-1, X homogeneous
//   rigid : keymorph/keypoint_aligners.py:151-213  Arun et al. (centroids, H = q1 q2^T, SVD,
//           R = V U^T, reflection step as the reference applies it (last ROW of V negated), T = c2 - R c1)
//   square + inverse: keymorph/transformations.py:25-35
// The moments are shuffle-reduced in fp64 and the 4x4 / 3x3 algebra runs in registers of lane 0;
// results are written as fp32.  (The reference computes in fp32 through cuBLAS / cuSOLVER; fp64
// here only makes the answer closer to the exact one.)
#include "km_common.cuh"

namespace {

// Gauss-Jordan inverse with partial pivoting. Returns false when a pivot is exactly zero or the
// matrix is not finite (torch.inverse raises LinAlgError in that case).
template <int n>
__device__ bool invert_gj(const double (&a_in)[n][n], double (&inv)[n][n]) {
  double a[n][2 * n];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      a[i][j] = a_in[i][j];
      a[i][n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int col = 0; col < n; ++col) {
    int piv = col;
    double best = fabs(a[col][col]);
    for (int r = col + 1; r < n; ++r)
      if (fabs(a[r][col]) > best) {
        best = fabs(a[r][col]);
        piv = r;
      }
    if (!(best > 0.0) || !isfinite(best)) return false;
    if (piv != col)
      for (int j = 0; j < 2 * n; ++j) {
        const double t = a[col][j];
        a[col][j] = a[piv][j];
        a[piv][j] = t;
      }
    const double d = 1.0 / a[col][col];
    for (int j = 0; j < 2 * n; ++j) a[col][j] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = a[r][col];
      if (f != 0.0)
        for (int j = 0; j < 2 * n; ++j) a[r][j] -= f * a[col][j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i][j] = a[i][n + j];
  return true;
}

__device__ void write44(const double (&A)[3][4], float* A44, float* A44_inv, int32_t* status) {
  double m[4][4], inv[4][4];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) m[i][j] = A[i][j];
  m[3][0] = m[3][1] = m[3][2] = 0.0;
  m[3][3] = 1.0;
  // the reference builds the square matrix in fp32 and inverts THAT (transformations.py:28,32-35)
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) m[i][j] = (double)(float)m[i][j];
  const bool ok = invert_gj<4>(m, inv);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      A44[i * 4 + j] = (float)m[i][j];
      A44_inv[i * 4 + j] = ok ? (float)inv[i][j] : __int_as_float(0x7fc00000);
    }
  if (!ok) *status |= 2;
}

__global__ void __launch_bounds__(32)
fit_affine_kernel(const float* __restrict__ x, const float* __restrict__ y,
                  const float* __restrict__ w, float* __restrict__ A44, float* __restrict__ A44_inv,
                  int32_t* __restrict__ status, int K) {
  const int n = blockIdx.x, lane = threadIdx.x;
  const float* xn = x + (size_t)n * K * 3;
  const float* yn = y + (size_t)n * K * 3;
  const float* wn = w ? w + (size_t)n * K : nullptr;
  double xx[10], yx[12];
#pragma unroll
  for (int i = 0; i < 10; ++i) xx[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 12; ++i) yx[i] = 0.0;
  for (int k = lane; k < K; k += 32) {
    const double wk = wn ? (double)wn[k] : 1.0;
    const double h[4] = {(double)xn[k * 3], (double)xn[k * 3 + 1], (double)xn[k * 3 + 2], 1.0};
    const double t[3] = {(double)yn[k * 3], (double)yn[k * 3 + 1], (double)yn[k * 3 + 2]};
    int q = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j) xx[q++] += wk * h[i] * h[j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) yx[i * 4 + j] += wk * t[i] * h[j];
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) xx[i] = km_warp_sum(xx[i]);
#pragma unroll
  for (int i = 0; i < 12; ++i) yx[i] = km_warp_sum(yx[i]);
  if (lane != 0) return;
  int32_t st = 0;
  double S[4][4], Sinv[4][4], A[3][4];
  int q = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = i; j < 4; ++j) {
      S[i][j] = xx[q];
      S[j][i] = xx[q];
      ++q;
    }
  const bool ok = invert_gj<4>(S, Sinv);
  if (!ok) st |= 1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double a = 0.0;
      for (int k = 0; k < 4; ++k) a += yx[i * 4 + k] * Sinv[k][j];
      A[i][j] = ok ? a : (double)__int_as_float(0x7fc00000);
    }
  write44(A, A44 + n * 16, A44_inv + n * 16, &st);
  status[n] = st;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// R (3x3) from the cross-covariance H = sum q1 q2^T, with R = V U^T (H = U S V^T), det(R) = +1.
__device__ void rotation_from_H(const double (&Hm)[3][3], double (&R)[3][3]) {
  // one-sided Jacobi: G = H * V with orthogonal columns
  double G[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      G[i][j] = Hm[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; ++i) {
          alpha += G[i][p] * G[i][p];
          beta += G[i][q] * G[i][q];
          gamma += G[i][p] * G[i][q];
        }
        if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; ++i) {
          const double gp = G[i][p], gq = G[i][q];
          G[i][p] = c * gp - s * gq;
          G[i][q] = s * gp + c * gq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]);
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (sig[ord[b]] > sig[ord[a]]) {
        const int t = ord[a];
        ord[a] = ord[b];
        ord[b] = t;
      }
  double U[3][3], Vs[3][3], s[3];
  for (int j = 0; j < 3; ++j) {
    s[j] = sig[ord[j]];
    for (int i = 0; i < 3; ++i) {
      Vs[i][j] = V[i][ord[j]];
      U[i][j] = (s[j] > 0.0) ? G[i][ord[j]] / s[j] : 0.0;
    }
  }
  // the inputs are fp32: singular values below ~1e-6 * s_max are rounding noise of the points.
  // Treating them as zero makes collinear / coplanar keypoint sets (reference KATs
  // test/test.py:259-413) resolve to the minimal rotation instead of a noise-driven one.
  const double tol = 1e-6 * s[0];
  int rank = 0;
  for (int j = 0; j < 3; ++j)
    if (s[j] > tol && s[j] > 0.0) ++rank;

  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
  if (rank == 0) return;
  if (rank == 1) {
    // only one direction is constrained: the minimal rotation taking u1 to v1
    const double a[3] = {U[0][0], U[1][0], U[2][0]};
    const double b[3] = {Vs[0][0], Vs[1][0], Vs[2][0]};
    double v[3];
    cross3(a, b, v);
    const double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    if (c > -1.0 + 1e-12) {
      const double k = 1.0 / (1.0 + c);
      const double vx[3][3] = {{0, -v[2], v[1]}, {v[2], 0, -v[0]}, {-v[1], v[0], 0}};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double v2 = 0.0;
          for (int m = 0; m < 3; ++m) v2 += vx[i][m] * vx[m][j];
          R[i][j] += vx[i][j] + k * v2;
        }
    } else {
      // u1 = -v1: half turn about an axis perpendicular to u1
      double e[3] = {1, 0, 0};
      if (fabs(a[0]) > 0.9) { e[0] = 0; e[1] = 1; }
      double ax[3];
      cross3(a, e, ax);
      const double nrm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      for (int i = 0; i < 3; ++i) ax[i] /= nrm;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = 2.0 * ax[i] * ax[j] - ((i == j) ? 1.0 : 0.0);
    }
    return;
  }
  if (rank == 2) {
    const double u1[3] = {U[0][0], U[1][0], U[2][0]}, u2[3] = {U[0][1], U[1][1], U[2][1]};
    double u3[3];
    cross3(u1, u2, u3);
    for (int i = 0; i < 3; ++i) U[i][2] = u3[i];
  }
  // R = V U^T, then the reference's reflection step (keymorph/keypoint_aligners.py:199-206): `dets` is
  // stacked along axis 1, so V * sign(dets) negates the last ROW of V, i.e. R <- diag(1, 1, sign det) V U^T.
  // (Arun et al. negate the last COLUMN of V; for mirror-related point sets the two differ and the
  // reference's choice is reproduced here -- parity first.)  When H has rank 2 the third singular
  // vectors were completed above and their relative sign is arbitrary: there the column flip selects the
  // proper rotation, which is what LAPACK's sign choice gives the reference in its coplanar KATs.
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double a = 0.0;
        for (int k = 0; k < 3; ++k) a += Vs[i][k] * U[j][k];
        R[i][j] = a;
      }
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) -
                       R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    if (det >= 0.0) break;
    if (rank == 3) {
      for (int j = 0; j < 3; ++j) R[2][j] = -R[2][j];
      break;
    }
    for (int i = 0; i < 3; ++i) Vs[i][2] = -Vs[i][2];
  }
}

__global__ void __launch_bounds__(32)
fit_rigid_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                 const float* __restrict__ w, float* __restrict__ A44, float* __restrict__ A44_inv,
                 int32_t* __restrict__ status, int K) {
  const int n = blockIdx.x, lane = threadIdx.x;
  const float* a = p1 + (size_t)n * K * 3;
  const float* b = p2 + (size_t)n * K * 3;
  const float* wn = w ? w + (size_t)n * K : nullptr;
  // centroids: weighted SUM when weights are given (keypoint_aligners.py:168-175), mean otherwise
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  for (int k = lane; k < K; k += 32) {
    const double wk = wn ? (double)wn[k] : 1.0;
    for (int i = 0; i < 3; ++i) {
      c1[i] += wk * (double)a[k * 3 + i];
      c2[i] += wk * (double)b[k * 3 + i];
    }
  }
  for (int i = 0; i < 3; ++i) {
    c1[i] = km_warp_sum(c1[i]);
    c2[i] = km_warp_sum(c2[i]);
    if (!wn) {
      c1[i] /= (double)K;
      c2[i] /= (double)K;
    }
  }
  double Hs[9];
  for (int i = 0; i < 9; ++i) Hs[i] = 0.0;
  for (int k = lane; k < K; k += 32) {
    const double wk = wn ? (double)wn[k] : 1.0;
    double q1[3], q2[3];
    for (int i = 0; i < 3; ++i) {
      q1[i] = ((double)a[k * 3 + i] - c1[i]) * wk;  // both sides are scaled by w (:181-183)
      q2[i] = ((double)b[k * 3 + i] - c2[i]) * wk;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Hs[i * 3 + j] += q1[i] * q2[j];
  }
  for (int i = 0; i < 9; ++i) Hs[i] = km_warp_sum(Hs[i]);
  if (lane != 0) return;
  double Hm[3][3], R[3][3], A[3][4];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Hm[i][j] = Hs[i * 3 + j];
  rotation_from_H(Hm, R);
  for (int i = 0; i < 3; ++i) {
    double t = c2[i];
    for (int j = 0; j < 3; ++j) {
      A[i][j] = R[i][j];
      t -= R[i][j] * c1[j];
    }
    A[i][3] = t;
  }
  int32_t st = 0;
  write44(A, A44 + n * 16, A44_inv + n * 16, &st);
  status[n] = st;
}

__global__ void inverse44_kernel(const float* __restrict__ m, float* __restrict__ inv,
                                 int32_t* __restrict__ status, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a[4][4], r[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a[i][j] = (double)m[n * 16 + i * 4 + j];
  const bool ok = invert_gj<4>(a, r);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      inv[n * 16 + i * 4 + j] = ok ? (float)r[i][j] : __int_as_float(0x7fc00000);
  if (status) status[n] = ok ? 0 : 2;
}

}  // namespace

extern "C" int km_fit_affine(const float* x, const float* y, const float* w, float* A44,
                             float* A44_inv, int32_t* status, int N, int K, km_stream_t stream) {
  KM_CHECK_ARG(x && y && A44 && A44_inv && status && N > 0 && K > 0, "km_fit_affine: bad arguments");
  fit_affine_kernel<<<N, 32, 0, km_cs(stream)>>>(x, y, w, A44, A44_inv, status, K);
  KM_LAUNCH_OK("fit_affine_kernel");
  return KM_OK;
}

extern "C" int km_fit_rigid(const float* x, const float* y, const float* w, float* A44,
                            float* A44_inv, int32_t* status, int N, int K, km_stream_t stream) {
  KM_CHECK_ARG(x && y && A44 && A44_inv && status && N > 0 && K > 0, "km_fit_rigid: bad arguments");
  fit_rigid_kernel<<<N, 32, 0, km_cs(stream)>>>(x, y, w, A44, A44_inv, status, K);
  KM_LAUNCH_OK("fit_rigid_kernel");
  return KM_OK;
}

extern "C" int km_inverse44(const float* m, float* inv, int32_t* status, int N,
                            km_stream_t stream) {
  KM_CHECK_ARG(m && inv && N > 0, "km_inverse44: bad arguments");
  inverse44_kernel<<<(N + 63) / 64, 64, 0, km_cs(stream)>>>(m, inv, status, N);
  KM_LAUNCH_OK("inverse44_kernel");
  return KM_OK;
}
