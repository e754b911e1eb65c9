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
#define PCR_HM_NOINLINE inline __host__ __device__ __noinline__  // own register allocation on the device (single-thread tails)
#else
#define PCR_HM inline
#define PCR_HM_NOINLINE inline
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

// Eigen::AngleAxisf(angle, UnitAxis).toRotationMatrix(), float (Eigen/src/Geometry/AngleAxis.h algorithm), from the sine and
// cosine of the angle (so that the device can evaluate the six trigonometric functions of a pose in parallel lanes)
PCR_HM void axis_rotation_sc_f32(float sn, float cs, int axis, float R[9]) {
  float u[3] = {0.f, 0.f, 0.f};
  u[axis] = 1.f;
  float su[3], cu[3];
  for (int i = 0; i < 3; i++) { su[i] = sn * u[i]; cu[i] = (1.f - cs) * u[i]; }
  float t;
  t = cu[0] * u[1]; R[0 * 3 + 1] = t - su[2]; R[1 * 3 + 0] = t + su[2];
  t = cu[0] * u[2]; R[0 * 3 + 2] = t + su[1]; R[2 * 3 + 0] = t - su[1];
  t = cu[1] * u[2]; R[1 * 3 + 2] = t - su[0]; R[2 * 3 + 1] = t + su[0];
  for (int i = 0; i < 3; i++) R[i * 3 + i] = cu[i] * u[i] + cs;
}
PCR_HM void axis_rotation_f32(float angle, int axis, float R[9]) { axis_rotation_sc_f32(sin_f32(angle), cos_f32(angle), axis, R); }

PCR_HM void mul3_f32(const float* A, const float* B, float* C) {
  float t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
  for (int i = 0; i < 9; i++) C[i] = t[i];
}

// (Translation3f(p0,p1,p2) * AngleAxisf(p3, X) * AngleAxisf(p4, Y) * AngleAxisf(p5, Z)).matrix() — ndt_omp_impl.hpp:827-830.
// M: column-major float[16]. sc = {sin p3, cos p3, sin p4, cos p4, sin p5, cos p5} of the angles cast to float.
PCR_HM void ndt_pose_matrix_sc_f32(const double p[6], const float sc[6], float M[16]) {
  float Rx[9], Ry[9], Rz[9], L[9];
  axis_rotation_sc_f32(sc[0], sc[1], 0, Rx);
  axis_rotation_sc_f32(sc[2], sc[3], 1, Ry);
  axis_rotation_sc_f32(sc[4], sc[5], 2, Rz);
  mul3_f32(Rx, Ry, L);
  mul3_f32(L, Rz, L);
  for (int i = 0; i < 16; i++) M[i] = 0.f;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) M[c * 4 + r] = L[r * 3 + c];
    M[12 + r] = static_cast<float>(p[r]);
  }
  M[15] = 1.f;
}
PCR_HM void ndt_pose_matrix_f32(const double p[6], float M[16]) {
  float sc[6];
  for (int a = 0; a < 3; a++) { const float ang = static_cast<float>(p[3 + a]); sc[2 * a] = sin_f32(ang); sc[2 * a + 1] = cos_f32(ang); }
  ndt_pose_matrix_sc_f32(p, sc, M);
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
// +sy, :383); jd/hd: double tables used by computeHessian (-sy, :361). The sine / cosine of an angle below 10e-5 are
// replaced by 0 / 1 (:293-322): ndt_angle_trig applies that rule, ndt_angle_tables_trig builds the tables from the six values.
PCR_HM void ndt_angle_trig(double angle, double& c, double& s) {
  if (fabs(angle) < 10e-5) { c = 1.0; s = 0.0; } else { c = cos(angle); s = sin(angle); }
}
PCR_HM void ndt_angle_tables_trig(double cx, double cy, double cz, double sx, double sy, double sz, float jf[8][3], float hf[15][3], double jd[8][3],
                                  double hd[15][3]) {
#define PCR_ROW(T, r, a, b, c) { const double v0 = (a), v1 = (b), v2 = (c); T##d[r][0] = v0; T##d[r][1] = v1; T##d[r][2] = v2; \
                                 T##f[r][0] = static_cast<float>(v0); T##f[r][1] = static_cast<float>(v1); T##f[r][2] = static_cast<float>(v2); }
  PCR_ROW(j, 0, (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy))
  PCR_ROW(j, 1, (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy))
  PCR_ROW(j, 2, (-sy * cz), sy * sz, cy)
  PCR_ROW(j, 3, sx * cy * cz, (-sx * cy * sz), sx * sy)
  PCR_ROW(j, 4, (-cx * cy * cz), cx * cy * sz, (-cx * sy))
  PCR_ROW(j, 5, (-cy * sz), (-cy * cz), 0.0)
  PCR_ROW(j, 6, (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0.0)
  PCR_ROW(j, 7, (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0.0)
  PCR_ROW(h, 0, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy)     // a2
  PCR_ROW(h, 1, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy))  // a3
  PCR_ROW(h, 2, (cx * cy * cz), (-cx * cy * sz), (cx * sy))                        // b2
  PCR_ROW(h, 3, (sx * cy * cz), (-sx * cy * sz), (sx * sy))                        // b3
  PCR_ROW(h, 4, (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0.0)          // c2
  PCR_ROW(h, 5, (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0.0)          // c3
  PCR_ROW(h, 6, (-cy * cz), (cy * sz), (-sy))                                      // d1 (double table)
  PCR_ROW(h, 7, (-sx * sy * cz), (sx * sy * sz), (sx * cy))                        // d2
  PCR_ROW(h, 8, (cx * sy * cz), (-cx * sy * sz), (-cx * cy))                       // d3
  PCR_ROW(h, 9, (sy * sz), (sy * cz), 0.0)                                         // e1
  PCR_ROW(h, 10, (-sx * cy * sz), (-sx * cy * cz), 0.0)                            // e2
  PCR_ROW(h, 11, (cx * cy * sz), (cx * cy * cz), 0.0)                              // e3
  PCR_ROW(h, 12, (-cy * cz), (cy * sz), 0.0)                                       // f1
  PCR_ROW(h, 13, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0.0)        // f2
  PCR_ROW(h, 14, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0.0)        // f3
#undef PCR_ROW
  hf[6][2] = static_cast<float>(sy);
}
PCR_HM void ndt_angle_tables(const double p[6], float jf[8][3], float hf[15][3], double jd[8][3], double hd[15][3]) {
  double cx, cy, cz, sx, sy, sz;
  ndt_angle_trig(p[3], cx, sx);
  ndt_angle_trig(p[4], cy, sy);
  ndt_angle_trig(p[5], cz, sz);
  ndt_angle_tables_trig(cx, cy, cz, sx, sy, sz, jf, hf, jd, hd);
}

// x = V_r S_r^-1 U_r^T b of a 6x6 matrix via one-sided Jacobi SVD with Eigen's rank threshold —
// Eigen::JacobiSVD<Matrix6d>(H, FullU|FullV).solve(b) at ndt_omp_impl.hpp:127-129.
PCR_HM_NOINLINE void svd6_solve(const double* Arow, const double* b, double* x) {
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
PCR_HM_NOINLINE void solve6_newton(const double* Arow, const double* b, double* x) {
  // fully unrolled with static indices (partial pivoting through conditional row swaps): everything stays in registers
  double M[6][7];
  double amax = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j < 6; j++) { M[i][j] = Arow[i * 6 + j]; const double a = fabs(M[i][j]); amax = a > amax ? a : amax; }
    M[i][6] = b[i];
  }
  bool ok = amax > 0.0 && amax == amax && amax < DBL_MAX;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double pv = fabs(M[k][k]);
#pragma unroll
    for (int i = k + 1; i < 6; i++) {  // bring the largest remaining entry of column k into row k (first maximum wins)
      const double a = fabs(M[i][k]);
      const bool sw = a > pv;
      pv = sw ? a : pv;
#pragma unroll
      for (int j = k; j < 7; j++) { const double t = M[k][j]; M[k][j] = sw ? M[i][j] : t; M[i][j] = sw ? t : M[i][j]; }
    }
    if (!(pv > 1e-10 * amax)) ok = false;
    const double inv = 1.0 / M[k][k];
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      const double f = M[i][k] * inv;
#pragma unroll
      for (int j = k + 1; j < 7; j++) M[i][j] -= f * M[k][j];
    }
  }
  if (!ok) { svd6_solve(Arow, b, x); return; }
  double y[6];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double v = M[i][6];
#pragma unroll
    for (int j = i + 1; j < 6; j++) v -= M[i][j] * y[j];
    y[i] = v / M[i][i];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) x[i] = y[i];
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
