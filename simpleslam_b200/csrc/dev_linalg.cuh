// Small dense FP64 linear algebra shared by the kernels and the host-side drivers (host+device).
// Each routine follows the Eigen algorithm the reference calls (cited per function) so that accept/reject
// decisions at thresholds match the reference's CPU path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

#define PCR_HD __host__ __device__ __forceinline__

namespace pcr {

// ----------------------------------------------------------------------------------------------------------------
// x = argmin |A x - b| via column-pivoted Householder QR, A is 5x3, basic solution using Eigen's
// nonzeroPivots() rule — Eigen::ColPivHouseholderQR::compute/solve as called at PCR/src/LoamRegister.cpp:34.
// a: 5x3 row-major (destroyed). b: 5 (destroyed).
// ----------------------------------------------------------------------------------------------------------------
PCR_HD void cpqr5x3_solve(double (&a)[5][3], double (&b)[5], double (&x)[3]) {
  int perm[3] = {0, 1, 2};
  double hc[3];
  double nu[3], nd[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; i++) s += a[i][k] * a[i][k];
    nu[k] = nd[k] = sqrt(s);
  }
  const double eps = DBL_EPSILON;
  double maxn = fmax(nu[0], fmax(nu[1], nu[2]));
  const double thr_helper = (maxn * eps) * (maxn * eps) / 5.0;
  const double downdate_thr = sqrt(eps);
  int nzp = 3;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int big = k;
    double bigv = nu[k];
#pragma unroll
    for (int j = k + 1; j < 3; j++)
      if (nu[j] > bigv) { bigv = nu[j]; big = j; }
    if (nzp == 3 && bigv * bigv < thr_helper * double(5 - k)) nzp = k;
    if (big != k) {
#pragma unroll
      for (int i = 0; i < 5; i++) { double t = a[i][k]; a[i][k] = a[i][big]; a[i][big] = t; }
      double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
      t = nd[k]; nd[k] = nd[big]; nd[big] = t;
      int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
    }
    double tail = 0.0;
#pragma unroll
    for (int i = k + 1; i < 5; i++) tail += a[i][k] * a[i][k];
    double c0 = a[k][k], tau, beta;
    if (tail <= DBL_MIN) {
      tau = 0.0; beta = c0;
#pragma unroll
      for (int i = k + 1; i < 5; i++) a[i][k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
      double den = c0 - beta;
#pragma unroll
      for (int i = k + 1; i < 5; i++) a[i][k] = a[i][k] / den;
      tau = (beta - c0) / beta;
    }
    a[k][k] = beta;
    hc[k] = tau;
    if (tau != 0.0) {
#pragma unroll
      for (int j = k + 1; j < 3; j++) {
        double tmp = 0.0;
#pragma unroll
        for (int i = k + 1; i < 5; i++) tmp += a[i][k] * a[i][j];
        tmp += a[k][j];
        a[k][j] -= tau * tmp;
#pragma unroll
        for (int i = k + 1; i < 5; i++) a[i][j] -= tau * a[i][k] * tmp;
      }
    }
#pragma unroll
    for (int j = k + 1; j < 3; j++) {
      if (nu[j] != 0.0) {
        double t = fabs(a[k][j]) / nu[j];
        t = (1.0 + t) * (1.0 - t);
        t = t < 0.0 ? 0.0 : t;
        double r = nu[j] / nd[j];
        double t2 = t * r * r;
        if (t2 <= downdate_thr) {
          double s = 0.0;
#pragma unroll
          for (int i = k + 1; i < 5; i++) s += a[i][j] * a[i][j];
          nd[j] = sqrt(s);
          nu[j] = nd[j];
        } else {
          nu[j] *= sqrt(t);
        }
      }
    }
  }
  if (nzp == 0) { x[0] = x[1] = x[2] = 0.0; return; }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (k < nzp && hc[k] != 0.0) {
      double tmp = 0.0;
#pragma unroll
      for (int i = k + 1; i < 5; i++) tmp += a[i][k] * b[i];
      tmp += b[k];
      b[k] -= hc[k] * tmp;
#pragma unroll
      for (int i = k + 1; i < 5; i++) b[i] -= hc[k] * a[i][k] * tmp;
    }
  }
#pragma unroll
  for (int i = 2; i >= 0; i--) {
    if (i < nzp) {
      b[i] /= a[i][i];
#pragma unroll
      for (int r = 0; r < i; r++) b[r] -= b[i] * a[r][i];
    }
  }
  double y[3];
#pragma unroll
  for (int i = 0; i < 3; i++) y[i] = (i < nzp) ? b[i] : 0.0;
  // x[perm[i]] = y[i] without dynamic register indexing
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (perm[i] == j) x[j] = y[i];
  }
}

// ----------------------------------------------------------------------------------------------------------------
// 6x6 symmetric solve via LDL^T with diagonal pivoting — Eigen::LDLT as called at PCR/src/LoamRegister.cpp:198 and
// third_parties/pclomp/src/lsq_registration_impl.hpp:111,136. A: full symmetric row-major.
// ----------------------------------------------------------------------------------------------------------------
PCR_HD void ldlt6_solve(const double* A, const double* rhs, double* x) {
  double m[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) m[i][j] = (j <= i) ? A[i * 6 + j] : A[j * 6 + i];
  int tr[6];
  bool zero = false;
  for (int k = 0; k < 6; k++) {
    int big = k;
    double bigv = fabs(m[k][k]);
    for (int i = k + 1; i < 6; i++)
      if (fabs(m[i][i]) > bigv) { bigv = fabs(m[i][i]); big = i; }
    tr[k] = big;
    if (big != k) {
      for (int j = 0; j < 6; j++) { double t = m[k][j]; m[k][j] = m[big][j]; m[big][j] = t; }
      for (int i = 0; i < 6; i++) { double t = m[i][k]; m[i][k] = m[i][big]; m[i][big] = t; }
    }
    double temp[6];
    for (int j = 0; j < k; j++) temp[j] = m[j][j] * m[k][j];
    double akk = m[k][k];
    for (int j = 0; j < k; j++) akk -= m[k][j] * temp[j];
    m[k][k] = akk;
    for (int i = k + 1; i < 6; i++) {
      double v = m[i][k];
      for (int j = 0; j < k; j++) v -= m[i][j] * temp[j];
      m[i][k] = v;
    }
    bool valid = fabs(akk) > 0.0;
    if (k == 0 && !valid) {
      for (int j = 1; j < 6; j++) tr[j] = j;
      zero = true;
      break;
    }
    if (valid)
      for (int i = k + 1; i < 6; i++) m[i][k] /= akk;
    for (int i = k + 1; i < 6; i++) m[k][i] = m[i][k];
  }
  double y[6];
  for (int i = 0; i < 6; i++) y[i] = rhs[i];
  if (zero) { for (int i = 0; i < 6; i++) x[i] = 0.0; return; }
  for (int k = 0; k < 6; k++)
    if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < i; j++) y[i] -= m[i][j] * y[j];
  for (int i = 0; i < 6; i++) {
    if (fabs(m[i][i]) > DBL_MIN) y[i] /= m[i][i];
    else y[i] = 0.0;
  }
  for (int i = 5; i >= 0; i--)
    for (int j = i + 1; j < 6; j++) y[i] -= m[j][i] * y[j];
  for (int k = 5; k >= 0; k--)
    if (tr[k] != k) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  for (int i = 0; i < 6; i++) x[i] = y[i];
}

#ifdef __CUDACC__
// Warp-cooperative version of ldlt6_solve for the on-device Gauss-Newton step: the same pivoted LDL^T (Eigen::LDLT)
// on a 6x6 matrix held in shared memory. m: 36 doubles row-major (destroyed), y: rhs in / solution out, tr: 6 ints.
// Must be called by all 32 lanes of one warp.
__device__ __forceinline__ void ldlt6_solve_warp(double* m, double* y, int* tr, int lane) {
  bool zero = false;
  for (int k = 0; k < 6; k++) {
    int big = k;
    double bigv = fabs(m[k * 6 + k]);
    for (int i = k + 1; i < 6; i++) {
      const double v = fabs(m[i * 6 + i]);
      if (v > bigv) { bigv = v; big = i; }
    }
    if (lane == 0) tr[k] = big;
    if (big != k) {  // symmetric row / column swap
      if (lane < 6) { const double t = m[k * 6 + lane]; m[k * 6 + lane] = m[big * 6 + lane]; m[big * 6 + lane] = t; }
      __syncwarp();
      if (lane < 6) { const double t = m[lane * 6 + k]; m[lane * 6 + k] = m[lane * 6 + big]; m[lane * 6 + big] = t; }
      __syncwarp();
    }
    double akk = m[k * 6 + k];
    for (int j = 0; j < k; j++) akk -= m[k * 6 + j] * (m[j * 6 + j] * m[k * 6 + j]);
    const int i = k + 1 + lane;
    const bool act = i < 6;
    double v = 0.0;
    if (act) {
      v = m[i * 6 + k];
      for (int j = 0; j < k; j++) v -= m[i * 6 + j] * (m[j * 6 + j] * m[k * 6 + j]);
    }
    __syncwarp();
    if (lane == 0) m[k * 6 + k] = akk;
    const bool valid = fabs(akk) > 0.0;
    if (k == 0 && !valid) { zero = true; break; }
    if (act) {
      if (valid) v /= akk;
      m[i * 6 + k] = v;
      m[k * 6 + i] = v;
    }
    __syncwarp();
  }
  if (zero) {
    if (lane < 6) y[lane] = 0.0;
    __syncwarp();
    return;
  }
  if (lane == 0) {
    for (int k = 0; k < 6; k++) {
      const int t = tr[k];
      if (t != k) { const double u = y[k]; y[k] = y[t]; y[t] = u; }
    }
    double Y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) Y[i] = y[i];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < i; j++) Y[i] -= m[i * 6 + j] * Y[j];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const double d = m[i * 6 + i];
      Y[i] = (fabs(d) > DBL_MIN) ? Y[i] / d : 0.0;
    }
#pragma unroll
    for (int i = 5; i >= 0; i--)
#pragma unroll
      for (int j = i + 1; j < 6; j++) Y[i] -= m[j * 6 + i] * Y[j];
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = Y[i];
    for (int k = 5; k >= 0; k--) {
      const int t = tr[k];
      if (t != k) { const double u = y[k]; y[k] = y[t]; y[t] = u; }
    }
  }
  __syncwarp();
}
#endif

// ----------------------------------------------------------------------------------------------------------------
// SE(3) exponential, ordering [rho; omega], left Jacobian V — geometry::manifolds::exp
// (common/geometry/manifolds.hpp:33-60). E: column-major 4x4.
// ----------------------------------------------------------------------------------------------------------------
PCR_HD void se3_exp(const double* k, double* E) {
  double t = sqrt(k[3] * k[3] + k[4] * k[4] + k[5] * k[5]);
  for (int i = 0; i < 16; i++) E[i] = (i % 5 == 0) ? 1.0 : 0.0;
  if (t < 1e-6) { E[12] = k[0]; E[13] = k[1]; E[14] = k[2]; return; }
  double a[3] = {k[3] / t, k[4] / t, k[5] / t};
  double ct = cos(t), st = sin(t);
  double ah[3][3] = {{0.0, -a[2], a[1]}, {a[2], 0.0, -a[0]}, {-a[1], a[0], 0.0}};
  double V[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double I = (i == j) ? 1.0 : 0.0, aa = a[i] * a[j];
      E[j * 4 + i] = ct * I + (1.0 - ct) * aa + st * ah[i][j];
      V[i][j] = st / t * I + (1.0 - st / t) * aa + ((1.0 - ct) / t) * ah[i][j];
    }
  for (int i = 0; i < 3; i++) E[12 + i] = V[i][0] * k[0] + V[i][1] * k[1] + V[i][2] * k[2];
}

// C = A * B, column-major 4x4
PCR_HD void mat4_mul(const double* A, const double* B, double* C) {
  double t[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      double v = 0.0;
      for (int q = 0; q < 4; q++) v += A[q * 4 + r] * B[c * 4 + q];
      t[c * 4 + r] = v;
    }
  for (int i = 0; i < 16; i++) C[i] = t[i];
}

// Rotation block -> unit quaternion -> rotation: geometry::trans::T2SE3 (common/geometry/trans.hpp:54-65), i.e.
// Eigen::Quaterniond(R).normalized().toRotationMatrix(). T column-major.
PCR_HD void t2se3(double* T) {
#define PCR_M(r, c) T[(c) * 4 + (r)]
  double q[4];
  double t = PCR_M(0, 0) + PCR_M(1, 1) + PCR_M(2, 2);
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (PCR_M(2, 1) - PCR_M(1, 2)) * t;
    q[1] = (PCR_M(0, 2) - PCR_M(2, 0)) * t;
    q[2] = (PCR_M(1, 0) - PCR_M(0, 1)) * t;
  } else {
    int i = 0;
    if (PCR_M(1, 1) > PCR_M(0, 0)) i = 1;
    if (PCR_M(2, 2) > PCR_M(i, i)) i = 2;
    int j = (i + 1) % 3, kk = (j + 1) % 3;
    t = sqrt(PCR_M(i, i) - PCR_M(j, j) - PCR_M(kk, kk) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (PCR_M(kk, j) - PCR_M(j, kk)) * t;
    q[j] = (PCR_M(j, i) + PCR_M(i, j)) * t;
    q[kk] = (PCR_M(kk, i) + PCR_M(i, kk)) * t;
  }
  double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double x = q[0] / nrm, y = q[1] / nrm, z = q[2] / nrm, w = q[3] / nrm;
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
         tzz = tz * z;
  PCR_M(0, 0) = 1 - (tyy + tzz); PCR_M(0, 1) = txy - twz; PCR_M(0, 2) = txz + twy;
  PCR_M(1, 0) = txy + twz; PCR_M(1, 1) = 1 - (txx + tzz); PCR_M(1, 2) = tyz - twx;
  PCR_M(2, 0) = txz - twy; PCR_M(2, 1) = tyz + twx; PCR_M(2, 2) = 1 - (txx + tyy);
#undef PCR_M
}

// ----------------------------------------------------------------------------------------------------------------
// symmetric 3x3 eigen-decomposition, cyclic Jacobi; ascending eigenvalues, eigenvectors in columns of V —
// the contract of Eigen::SelfAdjointEigenSolver<Matrix3d> relied on at voxel_grid_covariance_omp_impl.hpp:333-353
// and (as the SVD of a PSD matrix) at fast_gicp_impl.hpp:272.
// ----------------------------------------------------------------------------------------------------------------
PCR_HD void eig_sym3(const double (&Ain)[3][3], double (&w)[3], double (&V)[3][3]) {
  double a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { a[i][j] = 0.5 * (Ain[i][j] + Ain[j][i]); V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (w[j] > w[j + 1]) {
        double t = w[j]; w[j] = w[j + 1]; w[j + 1] = t;
        for (int k = 0; k < 3; k++) { double u = V[k][j]; V[k][j] = V[k][j + 1]; V[k][j + 1] = u; }
      }
}

// general 3x3 inverse by cofactors (Eigen Matrix3d::inverse()).
PCR_HD void inv3(const double (&m)[3][3], double (&o)[3][3]) {
  double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
  double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
  double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
  double id = 1.0 / det;
  o[0][0] = c00 * id; o[1][0] = c01 * id; o[2][0] = c02 * id;
  o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
}

}  // namespace pcr
