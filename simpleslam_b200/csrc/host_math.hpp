// Scalar math of the registration drivers (rows N5 / V5 of SURVEY.md §8a): Euler / float pose assembly and angular
// derivative tables of pclomp NDT, More-Thuente line search, 6x6 solve, SO(3) exp for the VGICP LM driver.
// Host AND device: the NDT Newton / More-Thuente state machine runs in the tail of the evaluation kernels (ndt_logic.cuh),
// the same functions compiled for the host are unit-tested on the CPU (tests/test_product_linalg.py, tests/test_ndt_logic.py).
#pragma once
#include <math.h>
#include <float.h>
#include <string.h>

#if defined(__CUDACC__)
#define PCR_HM __host__ __device__ inline
#else
#define PCR_HM inline
#endif

namespace pcr {
namespace hm {

// std::min / std::max with their exact NaN behaviour (the line search feeds NaN trial values through them on flat cost
// surfaces: std::min(a, b) = (b < a) ? b : a, std::max(a, b) = (a < b) ? b : a)
PCR_HM double std_min(double a, double b) { return (b < a) ? b : a; }
PCR_HM double std_max(double a, double b) { return (a < b) ? b : a; }

// float sin / cos / atan2 / sqrt as glibc gives them on the reference's CPU path (nearly correctly rounded): on the device
// they are evaluated in double and rounded once, which is the correctly rounded float except on vanishingly rare inputs
PCR_HM float sin_f32(float x) {
#ifdef __CUDA_ARCH__
  return static_cast<float>(sin(static_cast<double>(x)));
#else
  return sinf(x);
#endif
}
PCR_HM float cos_f32(float x) {
#ifdef __CUDA_ARCH__
  return static_cast<float>(cos(static_cast<double>(x)));
#else
  return cosf(x);
#endif
}
PCR_HM float atan2_f32(float y, float x) {
#ifdef __CUDA_ARCH__
  return static_cast<float>(atan2(static_cast<double>(y), static_cast<double>(x)));
#else
  return atan2f(y, x);
#endif
}

// Eigen::AngleAxisf(angle, UnitAxis).toRotationMatrix(), float (Eigen/src/Geometry/AngleAxis.h algorithm)
PCR_HM void axis_rotation_f32(float angle, int axis, float R[9]) {
  float u[3] = {0.f, 0.f, 0.f};
  u[axis] = 1.f;
  const float sn = sin_f32(angle), cs = cos_f32(angle);
  float su[3], cu[3];
  for (int i = 0; i < 3; i++) { su[i] = sn * u[i]; cu[i] = (1.f - cs) * u[i]; }
  float t;
  t = cu[0] * u[1]; R[0 * 3 + 1] = t - su[2]; R[1 * 3 + 0] = t + su[2];
  t = cu[0] * u[2]; R[0 * 3 + 2] = t + su[1]; R[2 * 3 + 0] = t - su[1];
  t = cu[1] * u[2]; R[1 * 3 + 2] = t - su[0]; R[2 * 3 + 1] = t + su[0];
  for (int i = 0; i < 3; i++) R[i * 3 + i] = cu[i] * u[i] + cs;
}

PCR_HM void mul3_f32(const float* A, const float* B, float* C) {
  float t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
  for (int i = 0; i < 9; i++) C[i] = t[i];
}

// (Translation3f(p0,p1,p2) * AngleAxisf(p3, X) * AngleAxisf(p4, Y) * AngleAxisf(p5, Z)).matrix() — ndt_omp_impl.hpp:827-830.
// M: column-major float[16].
PCR_HM void ndt_pose_matrix_f32(const double p[6], float M[16]) {
  float Rx[9], Ry[9], Rz[9], L[9];
  axis_rotation_f32(static_cast<float>(p[3]), 0, Rx);
  axis_rotation_f32(static_cast<float>(p[4]), 1, Ry);
  axis_rotation_f32(static_cast<float>(p[5]), 2, Rz);
  mul3_f32(Rx, Ry, L);
  mul3_f32(L, Rz, L);
  for (int i = 0; i < 16; i++) M[i] = 0.f;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) M[c * 4 + r] = L[r * 3 + c];
    M[12 + r] = static_cast<float>(p[r]);
  }
  M[15] = 1.f;
}

// Matrix3f::eulerAngles(0,1,2), Eigen 3.3 convention (first angle in [0, pi] before the final sign flip).
// R row-major float[9].
PCR_HM void euler_xyz_f32(const float R[9], float e[3]) {
#define PCR_M(r, c) R[(r) * 3 + (c)]
  e[0] = atan2_f32(PCR_M(1, 2), PCR_M(2, 2));
  const float c2 = sqrtf(PCR_M(0, 0) * PCR_M(0, 0) + PCR_M(0, 1) * PCR_M(0, 1));
  if (e[0] > 0.f) {
    e[0] -= static_cast<float>(3.14159265358979323846);
    e[1] = atan2_f32(-PCR_M(0, 2), -c2);
  } else {
    e[1] = atan2_f32(-PCR_M(0, 2), c2);
  }
  const float s1 = sin_f32(e[0]), c1 = cos_f32(e[0]);
  e[2] = atan2_f32(s1 * PCR_M(2, 0) - c1 * PCR_M(1, 0), c1 * PCR_M(1, 1) - s1 * PCR_M(2, 1));
#undef PCR_M
  e[0] = -e[0]; e[1] = -e[1]; e[2] = -e[2];
}

// computeAngleDerivatives (ndt_omp_impl.hpp:289-395). jf/hf: float tables used by computeDerivatives (hf row 6 has
// +sy, :383); jd/hd: double tables used by computeHessian (-sy, :361).
PCR_HM void ndt_angle_tables(const double p[6], float jf[8][3], float hf[15][3], double jd[8][3], double hd[15][3]) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
  const double j[8][3] = {{(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
                          {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
                          {(-sy * cz), sy * sz, cy},
                          {sx * cy * cz, (-sx * cy * sz), sx * sy},
                          {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
                          {(-cy * sz), (-cy * cz), 0},
                          {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
                          {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
  const double h[15][3] = {{(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},    // a2
                           {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)}, // a3
                           {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},                       // b2
                           {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},                       // b3
                           {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},           // c2
                           {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},           // c3
                           {(-cy * cz), (cy * sz), (-sy)},                                     // d1 (double table)
                           {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},                       // d2
                           {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},                      // d3
                           {(sy * sz), (sy * cz), 0},                                          // e1
                           {(-sx * cy * sz), (-sx * cy * cz), 0},                              // e2
                           {(cx * cy * sz), (cx * cy * cz), 0},                                // e3
                           {(-cy * cz), (cy * sz), 0},                                         // f1
                           {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},          // f2
                           {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};         // f3
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 3; c++) { jd[r][c] = j[r][c]; jf[r][c] = static_cast<float>(j[r][c]); }
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 3; c++) { hd[r][c] = h[r][c]; hf[r][c] = static_cast<float>(h[r][c]); }
  hf[6][2] = static_cast<float>(sy);
}

// x = V_r S_r^-1 U_r^T b of a 6x6 matrix via one-sided Jacobi SVD with Eigen's rank threshold —
// Eigen::JacobiSVD<Matrix6d>(H, FullU|FullV).solve(b) at ndt_omp_impl.hpp:127-129.
PCR_HM void svd6_solve(const double* Arow, const double* b, double* x) {
  const int N = 6;
  double U[6][6], V[6][6];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) { U[i][j] = Arow[i * 6 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 100; sweep++) {
    bool rotated = false;
    for (int p = 0; p < N - 1; p++)
      for (int q = p + 1; q < N; q++) {
        double al = 0, be = 0, ga = 0;
        for (int k = 0; k < N; k++) { al += U[k][p] * U[k][p]; be += U[k][q] * U[k][q]; ga += U[k][p] * U[k][q]; }
        if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < N; k++) {
          const double up = U[k][p], uq = U[k][q];
          U[k][p] = c * up - s * uq; U[k][q] = s * up + c * uq;
          const double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double S[6];
  int order[6];
  for (int j = 0; j < N; j++) {
    double s = 0;
    for (int k = 0; k < N; k++) s += U[k][j] * U[k][j];
    S[j] = sqrt(s);
    order[j] = j;
  }
  for (int i = 1; i < N; i++) {  // insertion sort, descending singular values (stable like the comparator it replaces)
    const int o = order[i];
    int j = i - 1;
    while (j >= 0 && S[order[j]] < S[o]) { order[j + 1] = order[j]; j--; }
    order[j + 1] = o;
  }
  const double thr = std_max(S[order[0]] * double(N) * DBL_EPSILON, DBL_MIN);
  for (int i = 0; i < N; i++) x[i] = 0;
  for (int r = 0; r < N; r++) {
    const int j = order[r];
    if (S[j] < thr) break;
    double ub = 0;
    for (int k = 0; k < N; k++) ub += U[k][j] * b[k];
    const double coef = ub / (S[j] * S[j]);
    for (int k = 0; k < N; k++) x[k] += V[k][j] * coef;
  }
}

// The same solve for the device-side Newton step: when H is comfortably full rank (every pivot of a partially pivoted
// elimination above 1e-10 of the largest one — Eigen's truncation only starts at singular values below 6 eps of the
// largest) JacobiSVD::solve is H^-1 b, which the elimination delivers in ~150 dependent flops instead of ~10^4; a
// (near-)singular H falls back to the Jacobi SVD above, truncation rule included.
PCR_HM void solve6_newton(const double* Arow, const double* b, double* x) {
  double M[6][7];
  double amax = 0.0;
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) { M[i][j] = Arow[i * 6 + j]; const double a = fabs(M[i][j]); amax = a > amax ? a : amax; }
    M[i][6] = b[i];
  }
  bool ok = amax > 0.0 && amax == amax && amax < DBL_MAX;
  for (int k = 0; k < 6 && ok; k++) {
    int piv = k;
    double pv = fabs(M[k][k]);
    for (int i = k + 1; i < 6; i++) { const double a = fabs(M[i][k]); if (a > pv) { pv = a; piv = i; } }
    if (!(pv > 1e-10 * amax)) { ok = false; break; }
    if (piv != k)
      for (int j = k; j < 7; j++) { const double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
    const double inv = 1.0 / M[k][k];
    for (int i = k + 1; i < 6; i++) {
      const double f = M[i][k] * inv;
      for (int j = k + 1; j < 7; j++) M[i][j] -= f * M[k][j];
    }
  }
  if (!ok) { svd6_solve(Arow, b, x); return; }
  for (int i = 5; i >= 0; i--) {
    double v = M[i][6];
    for (int j = i + 1; j < 6; j++) v -= M[i][j] * x[j];
    x[i] = v / M[i][i];
  }
}

// ---- More-Thuente helpers (ndt_omp_impl.hpp:649-769, ndt_omp.h:430-447) ----
PCR_HM double mt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
PCR_HM double mt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

PCR_HM bool mt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t,
                               double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

// minimiser of the cubic through (a0,f0,g0),(a1,f1,g1) — Sun & Yuan eq. 2.4.52 / 2.4.56
PCR_HM double mt_cubic(double a0, double f0, double g0, double a1, double f1, double g1) {
  const double z = 3 * (f1 - f0) / (a1 - a0) - g1 - g0;
  const double w = sqrt(z * z - g1 * g0);
  return a0 + (a1 - a0) * (w - g0 - z) / (g1 - g0 + 2 * w);
}

PCR_HM double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t,
                             double g_t) {
  if (f_t > f_l) {
    const double a_c = mt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return (fabs(a_c - a_l) < fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
  }
  if (g_t * g_l < 0) {
    const double a_c = mt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return (fabs(a_c - a_t) >= fabs(a_s - a_t)) ? a_c : a_s;
  }
  if (fabs(g_t) <= fabs(g_l)) {
    const double a_c = mt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    const double a_next = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
    const double lim = a_t + 0.66 * (a_u - a_t);
    return (a_t > a_l) ? std_min(lim, a_next) : std_max(lim, a_next);
  }
  return mt_cubic(a_u, f_u, g_u, a_t, f_t, g_t);
}

// so3_exp (third_parties/pclomp/src/so3/so3.hpp:58-77) -> Quaterniond::toRotationMatrix(); R row-major
PCR_HM void so3_exp_matrix(const double* om, double R[9]) {
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double im, re;
  if (th2 < 1e-10) {
    const double th4 = th2 * th2;
    im = 0.5 - 1.0 / 48.0 * th2 + 1.0 / 3840.0 * th4;
    re = 1.0 - 1.0 / 8.0 * th2 + 1.0 / 384.0 * th4;
  } else {
    const double th = sqrt(th2), half = 0.5 * th;
    im = sin(half) / th;
    re = cos(half);
  }
  const double w = re, x = im * om[0], y = im * om[1], z = im * om[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
               tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

}  // namespace hm
}  // namespace pcr
