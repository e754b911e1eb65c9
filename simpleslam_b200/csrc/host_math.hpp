// Host-side scalar drivers' math: Euler/float pose assembly and angular derivative tables of pclomp NDT,
// More-Thuente line search, 6x6 SVD solve, SO(3) exp for the VGICP LM driver.
// These are the "negligible compute" rows N5 / V5 of SURVEY.md §8a: they stay on the host (C++), one evaluation
// kernel launch per trial, exactly where the reference serialises as well.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace pcr {
namespace hm {

// Eigen::AngleAxisf(angle, UnitAxis).toRotationMatrix(), float (Eigen/src/Geometry/AngleAxis.h algorithm)
inline void axis_rotation_f32(float angle, int axis, float R[9]) {
  float u[3] = {0.f, 0.f, 0.f};
  u[axis] = 1.f;
  const float sn = std::sin(angle), cs = std::cos(angle);
  float su[3], cu[3];
  for (int i = 0; i < 3; i++) { su[i] = sn * u[i]; cu[i] = (1.f - cs) * u[i]; }
  float t;
  t = cu[0] * u[1]; R[0 * 3 + 1] = t - su[2]; R[1 * 3 + 0] = t + su[2];
  t = cu[0] * u[2]; R[0 * 3 + 2] = t + su[1]; R[2 * 3 + 0] = t - su[1];
  t = cu[1] * u[2]; R[1 * 3 + 2] = t - su[0]; R[2 * 3 + 1] = t + su[0];
  for (int i = 0; i < 3; i++) R[i * 3 + i] = cu[i] * u[i] + cs;
}

inline void mul3_f32(const float* A, const float* B, float* C) {
  float t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
  std::memcpy(C, t, sizeof(t));
}

// (Translation3f(p0,p1,p2) * AngleAxisf(p3, X) * AngleAxisf(p4, Y) * AngleAxisf(p5, Z)).matrix() — ndt_omp_impl.hpp:827-830.
// M: column-major float[16].
inline void ndt_pose_matrix_f32(const double p[6], float M[16]) {
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
inline void euler_xyz_f32(const float R[9], float e[3]) {
  auto m = [&](int r, int c) { return R[r * 3 + c]; };
  e[0] = std::atan2(m(1, 2), m(2, 2));
  const float c2 = std::sqrt(m(0, 0) * m(0, 0) + m(0, 1) * m(0, 1));
  if (e[0] > 0.f) {
    e[0] -= static_cast<float>(M_PI);
    e[1] = std::atan2(-m(0, 2), -c2);
  } else {
    e[1] = std::atan2(-m(0, 2), c2);
  }
  const float s1 = std::sin(e[0]), c1 = std::cos(e[0]);
  e[2] = std::atan2(s1 * m(2, 0) - c1 * m(1, 0), c1 * m(1, 1) - s1 * m(2, 1));
  e[0] = -e[0]; e[1] = -e[1]; e[2] = -e[2];
}

// computeAngleDerivatives (ndt_omp_impl.hpp:289-395). jf/hf: float tables used by computeDerivatives (hf row 6 has
// +sy, :383); jd/hd: double tables used by computeHessian (-sy, :361).
inline void ndt_angle_tables(const double p[6], float jf[8][3], float hf[15][3], double jd[8][3], double hd[15][3]) {
  double cx, cy, cz, sx, sy, sz;
  if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
  if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
  if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
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
inline void svd6_solve(const double* Arow, const double* b, double* x) {
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
        if (ga == 0.0 || std::fabs(ga) <= 1e-17 * std::sqrt(al * be)) continue;
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
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
    S[j] = std::sqrt(s);
    order[j] = j;
  }
  std::sort(order, order + N, [&](int a, int c) { return S[a] > S[c]; });
  const double thr = std::max(S[order[0]] * double(N) * std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::min());
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

// ---- More-Thuente helpers (ndt_omp_impl.hpp:649-769, ndt_omp.h:430-447) ----
inline double mt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
inline double mt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

inline bool mt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t,
                               double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

inline double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t,
                             double g_t) {
  auto cubic = [](double a0, double f0, double g0, double a1, double f1, double g1) {
    // minimiser of the cubic through (a0,f0,g0),(a1,f1,g1) — Sun & Yuan eq. 2.4.52 / 2.4.56
    const double z = 3 * (f1 - f0) / (a1 - a0) - g1 - g0;
    const double w = std::sqrt(z * z - g1 * g0);
    return a0 + (a1 - a0) * (w - g0 - z) / (g1 - g0 + 2 * w);
  };
  if (f_t > f_l) {
    const double a_c = cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
  }
  if (g_t * g_l < 0) {
    const double a_c = cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) ? a_c : a_s;
  }
  if (std::fabs(g_t) <= std::fabs(g_l)) {
    const double a_c = cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    const double a_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
    return (a_t > a_l) ? std::min(a_t + 0.66 * (a_u - a_t), a_next) : std::max(a_t + 0.66 * (a_u - a_t), a_next);
  }
  return cubic(a_u, f_u, g_u, a_t, f_t, g_t);
}

// so3_exp (third_parties/pclomp/src/so3/so3.hpp:58-77) -> Quaterniond::toRotationMatrix(); R row-major
inline void so3_exp_matrix(const double* om, double R[9]) {
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double im, re;
  if (th2 < 1e-10) {
    const double th4 = th2 * th2;
    im = 0.5 - 1.0 / 48.0 * th2 + 1.0 / 3840.0 * th4;
    re = 1.0 - 1.0 / 8.0 * th2 + 1.0 / 384.0 * th4;
  } else {
    const double th = std::sqrt(th2), half = 0.5 * th;
    im = std::sin(half) / th;
    re = std::cos(half);
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
