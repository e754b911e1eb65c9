// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY UNPINNED for this file: the reference ships no golden vectors for its dense linear
// algebra, and the routines live in an un-vendored dependency (Eigen, version unpinned by the
// reference's CMake: `find_package(Eigen3 REQUIRED)`, /root/reference/CMakeLists.txt:32; the
// author's platforms carry Eigen 3.3.4 / 3.3.7).  What follows restates Eigen 3.3's *published
// algorithms* (column-pivoted Householder QR, pivoted LDLT, Jacobi SVD / symmetric eigen) as used at
// the reference call sites cited on each function.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace orc {

// ------------------------------------------------------------------------------------------------
// 5x3 (generic rows x 3) column-pivoted Householder QR "solve", restating
// Eigen::ColPivHouseholderQR<Matrix<double,N,3>>::compute + ::solve as called at
// /root/reference/PCR/src/LoamRegister.cpp:34  (x = A.colPivHouseholderQr().solve(b)).
// Rank rule = Eigen 3.3 "nonzeroPivots()": pivot k is zero when the largest remaining *updated*
// column norm^2 < (eps * maxcolnorm)^2 / rows * (rows-k); the solve returns the *basic* solution.
// A is row-major rows x 3 (modified copy is internal). Returns number of nonzero pivots.
// ------------------------------------------------------------------------------------------------
template <int ROWS>
inline int cpqr_solve3(const double (&Ain)[ROWS][3], const double (&bin)[ROWS], double (&x)[3]) {
  constexpr int COLS = 3;
  double qr[ROWS][COLS];
  for (int i = 0; i < ROWS; i++)
    for (int j = 0; j < COLS; j++) qr[i][j] = Ain[i][j];
  double hcoef[COLS];
  int perm[COLS] = {0, 1, 2};
  double normsUpdated[COLS], normsDirect[COLS];
  for (int k = 0; k < COLS; k++) {
    double s = 0;
    for (int i = 0; i < ROWS; i++) s += qr[i][k] * qr[i][k];
    normsDirect[k] = normsUpdated[k] = std::sqrt(s);
  }
  const double eps = std::numeric_limits<double>::epsilon();
  double maxn = std::max(normsUpdated[0], std::max(normsUpdated[1], normsUpdated[2]));
  const double threshold_helper = (maxn * eps) * (maxn * eps) / double(ROWS);
  const double norm_downdate_threshold = std::sqrt(eps);
  int nonzero_pivots = COLS;
  for (int k = 0; k < COLS; k++) {
    int big = k;
    double bigv = normsUpdated[k];
    for (int j = k + 1; j < COLS; j++)
      if (normsUpdated[j] > bigv) { bigv = normsUpdated[j]; big = j; }
    double big_sq = bigv * bigv;
    if (nonzero_pivots == COLS && big_sq < threshold_helper * double(ROWS - k)) nonzero_pivots = k;
    if (big != k) {
      for (int i = 0; i < ROWS; i++) std::swap(qr[i][k], qr[i][big]);
      std::swap(normsUpdated[k], normsUpdated[big]);
      std::swap(normsDirect[k], normsDirect[big]);
      std::swap(perm[k], perm[big]);
    }
    // makeHouseholderInPlace on qr[k..ROWS-1][k]
    double tailSq = 0;
    for (int i = k + 1; i < ROWS; i++) tailSq += qr[i][k] * qr[i][k];
    double c0 = qr[k][k];
    double tau, beta;
    if (tailSq <= std::numeric_limits<double>::min()) {
      tau = 0; beta = c0;
      for (int i = k + 1; i < ROWS; i++) qr[i][k] = 0;
    } else {
      beta = std::sqrt(c0 * c0 + tailSq);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < ROWS; i++) qr[i][k] = qr[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    qr[k][k] = beta;
    hcoef[k] = tau;
    // apply H_k to the trailing columns
    if (tau != 0) {
      for (int j = k + 1; j < COLS; j++) {
        double tmp = 0;
        for (int i = k + 1; i < ROWS; i++) tmp += qr[i][k] * qr[i][j];
        tmp += qr[k][j];
        qr[k][j] -= tau * tmp;
        for (int i = k + 1; i < ROWS; i++) qr[i][j] -= tau * qr[i][k] * tmp;
      }
    }
    // norm downdate
    for (int j = k + 1; j < COLS; j++) {
      if (normsUpdated[j] != 0) {
        double temp = std::fabs(qr[k][j]) / normsUpdated[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0 ? 0 : temp;
        double r = normsUpdated[j] / normsDirect[j];
        double temp2 = temp * r * r;
        if (temp2 <= norm_downdate_threshold) {
          double s = 0;
          for (int i = k + 1; i < ROWS; i++) s += qr[i][j] * qr[i][j];
          normsDirect[j] = std::sqrt(s);
          normsUpdated[j] = normsDirect[j];
        } else {
          normsUpdated[j] *= std::sqrt(temp);
        }
      }
    }
  }
  // solve
  if (nonzero_pivots == 0) { x[0] = x[1] = x[2] = 0; return 0; }
  double c[ROWS];
  for (int i = 0; i < ROWS; i++) c[i] = bin[i];
  for (int k = 0; k < nonzero_pivots; k++) {
    double tau = hcoef[k];
    if (tau != 0) {
      double tmp = 0;
      for (int i = k + 1; i < ROWS; i++) tmp += qr[i][k] * c[i];
      tmp += c[k];
      c[k] -= tau * tmp;
      for (int i = k + 1; i < ROWS; i++) c[i] -= tau * qr[i][k] * tmp;
    }
  }
  // upper-triangular back substitution (column-oriented like Eigen's ColMajor small solver)
  for (int i = nonzero_pivots - 1; i >= 0; i--) {
    c[i] /= qr[i][i];
    for (int r = 0; r < i; r++) c[r] -= c[i] * qr[r][i];
  }
  for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = c[i];
  for (int i = nonzero_pivots; i < COLS; i++) x[perm[i]] = 0;
  return nonzero_pivots;
}

// ------------------------------------------------------------------------------------------------
// N x N symmetric LDLT with diagonal pivoting + solve, restating Eigen::LDLT<>::compute/solve as
// called at /root/reference/PCR/src/LoamRegister.cpp:198 and
// /root/reference/third_parties/pclomp/src/lsq_registration_impl.hpp:111,136.
// A row-major full symmetric (lower triangle is read). Returns false if the matrix is all zero.
// ------------------------------------------------------------------------------------------------
template <int N>
inline bool ldlt_solve(const double (&Ain)[N][N], const double (&b)[N], double (&x)[N]) {
  double m[N][N];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) m[i][j] = (j <= i) ? Ain[i][j] : Ain[j][i];  // use the lower triangle
  int tr[N];
  bool all_zero_tail = false;
  for (int k = 0; k < N; k++) {
    int big = k;
    double bigv = std::fabs(m[k][k]);
    for (int i = k + 1; i < N; i++)
      if (std::fabs(m[i][i]) > bigv) { bigv = std::fabs(m[i][i]); big = i; }
    tr[k] = big;
    if (big != k) {
      // symmetric row/col swap on the full (kept symmetric) matrix
      for (int j = 0; j < N; j++) std::swap(m[k][j], m[big][j]);
      for (int i = 0; i < N; i++) std::swap(m[i][k], m[i][big]);
    }
    // m(k,k) -= sum_j L(k,j)^2 D(j)
    double temp[N];
    for (int j = 0; j < k; j++) temp[j] = m[j][j] * m[k][j];
    double akk = m[k][k];
    for (int j = 0; j < k; j++) akk -= m[k][j] * temp[j];
    m[k][k] = akk;
    for (int i = k + 1; i < N; i++) {
      double v = m[i][k];
      for (int j = 0; j < k; j++) v -= m[i][j] * temp[j];
      m[i][k] = v;
    }
    bool pivot_valid = std::fabs(akk) > 0.0;
    if (k == 0 && !pivot_valid) {
      // whole matrix is zero
      for (int j = 1; j < N; j++) tr[j] = j;
      all_zero_tail = true;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) m[i][j] = 0;
      break;
    }
    if (pivot_valid)
      for (int i = k + 1; i < N; i++) m[i][k] /= akk;
    // mirror to keep the matrix symmetric for later swaps
    for (int i = k + 1; i < N; i++) m[k][i] = m[i][k];
  }
  // solve: x = P^T L^-T D^-1 L^-1 P b
  double y[N];
  for (int i = 0; i < N; i++) y[i] = b[i];
  for (int k = 0; k < N; k++)
    if (tr[k] != k) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < N; i++)
    for (int j = 0; j < i; j++) y[i] -= m[i][j] * y[j];
  const double tol = std::numeric_limits<double>::min();
  for (int i = 0; i < N; i++) {
    if (std::fabs(m[i][i]) > tol) y[i] /= m[i][i];
    else y[i] = 0;
  }
  for (int i = N - 1; i >= 0; i--)
    for (int j = i + 1; j < N; j++) y[i] -= m[j][i] * y[j];
  for (int k = N - 1; k >= 0; k--)
    if (tr[k] != k) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < N; i++) x[i] = y[i];
  return !all_zero_tail;
}

// ------------------------------------------------------------------------------------------------
// Symmetric 3x3 eigen-decomposition (cyclic Jacobi), eigenvalues ascending, eigenvectors in the
// columns of V. Restates the *contract* of Eigen::SelfAdjointEigenSolver<Matrix3d>::compute relied
// upon at /root/reference/third_parties/pclomp/src/voxel_grid_covariance_omp_impl.hpp:333-353
// (ascending eigenvalues, orthonormal eigenvectors); Eigen itself uses tridiagonal QL — results
// agree to rounding.
// ------------------------------------------------------------------------------------------------
inline void eig_sym3(const double (&Ain)[3][3], double (&w)[3], double (&V)[3][3]) {
  double a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { a[i][j] = 0.5 * (Ain[i][j] + Ain[j][i]); V[i][j] = (i == j); }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    double diag = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        // A <- J^T A J
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
  // sort ascending
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (w[j] > w[j + 1]) {
        std::swap(w[j], w[j + 1]);
        for (int k = 0; k < 3; k++) std::swap(V[k][j], V[k][j + 1]);
      }
}

// General 3x3 inverse by cofactors (Eigen's Matrix3d::inverse() is the cofactor formula).
inline bool inv3(const double (&m)[3][3], double (&o)[3][3]) {
  double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
  double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
  double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
  double id = 1.0 / det;
  o[0][0] = c00 * id;
  o[1][0] = c01 * id;
  o[2][0] = c02 * id;
  o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return std::isfinite(id);
}

// ------------------------------------------------------------------------------------------------
// N x N one-sided Jacobi SVD  A = U S V^T (singular values descending) and the rank-truncated
// solve  x = V_r S_r^-1 U_r^T b  with Eigen's SVDBase::rank() threshold (diagSize*eps*sigma_max).
// Restates Eigen::JacobiSVD<Matrix<double,6,6>>(H, FullU|FullV).solve(-g) at
// /root/reference/third_parties/pclomp/src/ndt_omp_impl.hpp:127-129 (Eigen uses two-sided Jacobi;
// the factorisation is unique up to signs so the solve agrees to rounding).
// ------------------------------------------------------------------------------------------------
template <int N>
inline void svd_solve(const double (&A)[N][N], const double (&b)[N], double (&x)[N]) {
  double U[N][N], V[N][N];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) { U[i][j] = A[i][j]; V[i][j] = (i == j); }
  for (int sweep = 0; sweep < 100; sweep++) {
    bool rotated = false;
    for (int p = 0; p < N - 1; p++)
      for (int q = p + 1; q < N; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < N; k++) {
          alpha += U[k][p] * U[k][p];
          beta += U[k][q] * U[k][q];
          gamma += U[k][p] * U[k][q];
        }
        if (gamma == 0.0) continue;
        if (std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < N; k++) {
          double up = U[k][p], uq = U[k][q];
          U[k][p] = c * up - s * uq;
          U[k][q] = s * up + c * uq;
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq;
          V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double S[N];
  int order[N];
  for (int j = 0; j < N; j++) {
    double s = 0;
    for (int k = 0; k < N; k++) s += U[k][j] * U[k][j];
    S[j] = std::sqrt(s);
    order[j] = j;
  }
  std::sort(order, order + N, [&](int a, int c) { return S[a] > S[c]; });
  double smax = S[order[0]];
  double thr = std::max(smax * double(N) * std::numeric_limits<double>::epsilon(),
                        std::numeric_limits<double>::min());
  for (int i = 0; i < N; i++) x[i] = 0;
  for (int r = 0; r < N; r++) {
    int j = order[r];
    if (S[j] < thr) break;
    // u_j = U[:,j]/S[j];  coefficient = (u_j . b) / S[j]
    double ub = 0;
    for (int k = 0; k < N; k++) ub += U[k][j] * b[k];
    double coef = ub / (S[j] * S[j]);
    for (int k = 0; k < N; k++) x[k] += V[k][j] * coef;
  }
}

}  // namespace orc
