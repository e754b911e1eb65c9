// ORACLE — TEST INFRASTRUCTURE ONLY (see pcr_oracle.h). PARITY UNPINNED (no reference fixtures).
// CPU restatement of: pcl::VoxelGrid downsample (A1), exact kNN (A3/A4), LoamRegister::scan2Map (A5-A9).
#include "pcr_oracle.h"
#include "orc_common.hpp"
#include "orc_linalg.hpp"
#include <omp.h>
#include <numeric>
#include <cstdio>

using namespace orc;

// ---------------------------------------------------------------------------------------------------
// A1. pcl::VoxelGrid<PointXYZI>::applyFilter as reached from common/pcp/pcp.hpp:15-28 and
// frontend/src/LidarOdometry.cpp:170-171; key math restated in-tree at pcp.hpp:191-210.
// Oracle conventions (documented deviations): output voxels ascending by key; the
// float32 centroid is summed in ascending input index (what a stable sort yields — PCL's std::sort
// leaves the intra-voxel order unspecified).
// ---------------------------------------------------------------------------------------------------
extern "C" int orc_voxel_downsample(const float* pts, size_t n, size_t stride_f, float leaf,
                                    int32_t* keys_out, float* out_pts, int32_t* out_keys,
                                    int32_t* out_counts, size_t* m, int32_t grid_out[9]) {
  Cloud c{pts, n, stride_f};
  VoxelGridSpec g = voxel_grid_spec(c, leaf);
  if (grid_out) {
    for (int a = 0; a < 3; a++) { grid_out[a] = g.min_b[a]; grid_out[3 + a] = g.div_b[a]; grid_out[6 + a] = g.mul[a]; }
  }
  if (n == 0) { *m = 0; return 0; }
  if (g.overflow) {
    // PCL: "Leaf size is too small for the input dataset" -> output = input
    for (size_t i = 0; i < n; i++) {
      const float* p = c.at(i);
      float* o = out_pts + i * 8;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = 1.0f; o[4] = c.intensity(i); o[5] = o[6] = o[7] = 0.f;
      if (keys_out) keys_out[i] = -1;
    }
    *m = n;
    return 1;
  }
  std::vector<int32_t> key(n);
  for (size_t i = 0; i < n; i++) key[i] = voxel_key(g, c.at(i));
  if (keys_out) std::copy(key.begin(), key.end(), keys_out);
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  size_t nv = 0;
  size_t i = 0;
  while (i < n) {
    size_t j = i;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;  // pcl::CentroidPoint: float32 accumulators
    while (j < n && key[order[j]] == key[order[i]]) {
      const float* p = c.at(order[j]);
      sx += p[0]; sy += p[1]; sz += p[2]; si += c.intensity(order[j]);
      j++;
    }
    float cnt = static_cast<float>(j - i);
    float* o = out_pts + nv * 8;
    o[0] = sx / cnt; o[1] = sy / cnt; o[2] = sz / cnt; o[3] = 1.0f;
    o[4] = si / cnt; o[5] = o[6] = o[7] = 0.f;
    if (out_keys) out_keys[nv] = key[order[i]];
    if (out_counts) out_counts[nv] = static_cast<int32_t>(j - i);
    nv++;
    i = j;
  }
  *m = nv;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// A3/A4. exact kNN
// ---------------------------------------------------------------------------------------------------
extern "C" int orc_knn(const float* map, size_t nm, size_t stride_f, const double* queries, size_t nq,
                       int k, int metric_float, int brute, float cell, int threads, int64_t* idx_out,
                       double* d2_out) {
  Cloud c{map, nm, stride_f};
  KnnGrid grid;
  if (!brute) grid.build(c, cell > 0 ? cell : 1.0f);
  if (threads <= 0) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
  for (long long qi = 0; qi < (long long)nq; qi++) {
    std::vector<int64_t> idx(k, -1);
    int cnt;
    if (metric_float) {
      std::vector<float> d2(k, 0.f);
      float q[3] = {float(queries[qi * 3]), float(queries[qi * 3 + 1]), float(queries[qi * 3 + 2])};
      cnt = brute ? knn_brute<float>(c, q, k, idx.data(), d2.data()) : grid.knn<float>(q, k, idx.data(), d2.data());
      for (int j = 0; j < k; j++) { idx_out[qi * k + j] = j < cnt ? idx[j] : -1; d2_out[qi * k + j] = j < cnt ? double(d2[j]) : -1.0; }
    } else {
      std::vector<double> d2(k, 0.0);
      const double* q = queries + qi * 3;
      cnt = brute ? knn_brute<double>(c, q, k, idx.data(), d2.data()) : grid.knn<double>(q, k, idx.data(), d2.data());
      for (int j = 0; j < k; j++) { idx_out[qi * k + j] = j < cnt ? idx[j] : -1; d2_out[qi * k + j] = j < cnt ? d2[j] : -1.0; }
    }
  }
  return 0;
}

// C-bar of SURVEY.md §8(d): mean number of map points in the 27 cells (cell = gate radius, 1 m) around a query — the
// per-unit figure of the LOAM algorithmic-bytes formula. Queries are doubles (xyz). Returns the mean.
extern "C" double orc_neighbourhood27(const float* map, size_t nm, size_t stride_f, const double* queries, size_t nq, float cell,
                                      int threads) {
  Cloud c{map, nm, stride_f};
  KnnGrid grid;
  grid.build(c, cell > 0 ? cell : 1.0f);
  if (threads <= 0) threads = 1;
  double total = 0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : total)
  for (long long qi = 0; qi < (long long)nq; qi++) {
    int c0[3];
    bool far = false;
    for (int a = 0; a < 3; a++) {
      double f = std::floor((queries[qi * 3 + a] - double(grid.origin[a])) / double(grid.cell));
      if (f < -1 || f > double(grid.dim[a])) far = true;
      c0[a] = int(std::max(-2.0, std::min(f, double(grid.dim[a]) + 1)));
    }
    if (far) continue;
    long long cnt = 0;
    for (int z = std::max(c0[2] - 1, 0); z <= std::min(c0[2] + 1, grid.dim[2] - 1); z++)
      for (int y = std::max(c0[1] - 1, 0); y <= std::min(c0[1] + 1, grid.dim[1] - 1); y++)
        for (int x = std::max(c0[0] - 1, 0); x <= std::min(c0[0] + 1, grid.dim[0] - 1); x++) {
          size_t cid = size_t(x) + size_t(grid.dim[0]) * (size_t(y) + size_t(grid.dim[1]) * size_t(z));
          cnt += grid.start[cid + 1] - grid.start[cid];
        }
    total += double(cnt);
  }
  return nq ? total / double(nq) : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// geometry: manifolds::exp(V6 -> M4) (common/geometry/manifolds.hpp:33-60), trans::T2SE3
// (common/geometry/trans.hpp:54-65) via Eigen::Quaternion(R).normalized().toRotationMatrix().
// ---------------------------------------------------------------------------------------------------
extern "C" void orc_se3_exp(const double k[6], double T[16]) {
  const double p[3] = {k[0], k[1], k[2]};
  const double w[3] = {k[3], k[4], k[5]};
  double t = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  if (t < 1e-6) { T[12] = p[0]; T[13] = p[1]; T[14] = p[2]; return; }
  double a[3] = {w[0] / t, w[1] / t, w[2] / t};
  double ct = std::cos(t), st = std::sin(t);
  double ah[3][3] = {{0, -a[2], a[1]}, {a[2], 0, -a[0]}, {-a[1], a[0], 0}};
  double R[3][3], V[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double I = (i == j) ? 1.0 : 0.0;
      double aa = a[i] * a[j];
      R[i][j] = ct * I + (1.0 - ct) * aa + std::sin(t) * ah[i][j];
      V[i][j] = st / t * I + (1.0 - st / t) * aa + ((1 - ct) / t) * ah[i][j];
    }
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[j * 4 + i] = R[i][j];
    T[12 + i] = V[i][0] * p[0] + V[i][1] * p[1] + V[i][2] * p[2];
  }
}

extern "C" void orc_t2se3(double T[16]) {
  auto m = [&](int r, int c) -> double { return T[c * 4 + r]; };
  double q[4];  // x y z w
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
  double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double x = q[0] / nrm, y = q[1] / nrm, z = q[2] / nrm, w = q[3] / nrm;
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
         tyz = tz * y, tzz = tz * z;
  double R[3][3] = {{1 - (tyy + tzz), txy - twz, txz + twy},
                    {txy + twz, 1 - (txx + tzz), tyz - twx},
                    {txz - twy, tyz + twx, 1 - (txx + tyy)}};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[c * 4 + r] = R[r][c];
}

// ---------------------------------------------------------------------------------------------------
// A5-A8. one LOAM linearisation (PCR/src/LoamRegister.cpp:122-188)
// ---------------------------------------------------------------------------------------------------
namespace {
struct LoamPoint {
  int status;
  double E;
  double J[6];
};

inline LoamPoint loam_point(const Cloud& src, size_t i, const Cloud& dst, const KnnGrid& grid, const double* T,
                            int64_t* idx5, double* d25) {
  LoamPoint out{};
  const float* po = src.at(i);
  // :128-130  V4 ori = pointOri.cast<double>(); ori = res * ori; pointInMap = ori.cast<float>()
  double ori[3] = {double(po[0]), double(po[1]), double(po[2])};
  double pm[3];
  transform_f64(T, ori, pm);
  float pmf[3] = {float(pm[0]), float(pm[1]), float(pm[2])};
  // :53-56 query = float coords widened to double; 5-NN
  double q[3] = {double(pmf[0]), double(pmf[1]), double(pmf[2])};
  for (int j = 0; j < 5; j++) { idx5[j] = -1; d25[j] = -1.0; }
  int cnt = grid.knn<double>(q, 5, idx5, d25);
  // :59 gate on the squared distance of the 5th neighbour
  const double kMaxSearch = double(1.0f);
  if (cnt < 5 || !(d25[4] < kMaxSearch)) { out.status = 0; return out; }
  double A[5][3];
  for (int j = 0; j < 5; j++) {
    const float* m = dst.at(size_t(idx5[j]));
    A[j][0] = m[0]; A[j][1] = m[1]; A[j][2] = m[2];
  }
  // :29-45 plane fit
  double b[5] = {-1, -1, -1, -1, -1};
  double x[3];
  cpqr_solve3<5>(A, b, x);
  double x_norm = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const double kPlaneValid = double(0.2f);
  for (int j = 0; j < 5; j++) {
    double v = x[0] * A[j][0] + x[1] * A[j][1] + x[2] * A[j][2];
    if (std::fabs(v + 1.0) > kPlaneValid * x_norm) { out.status = 1; return out; }
  }
  // :144-151
  double dist = ((q[0] * x[0] + q[1] * x[1] + q[2] * x[2]) + 1.0) / x_norm;
  float r2 = po[0] * po[0] + po[1] * po[1] + po[2] * po[2];
  float rr = std::sqrt(std::sqrt(r2));  // float sqrt of float sqrt
  double s = 1 - 0.9 * std::fabs(dist) / double(rr);
  const double kPointValid = double(0.1f);
  if (!(s > kPointValid)) { out.status = 2; return out; }
  out.status = 3;
  out.E = s * dist;
  double sn[3] = {s * (x[0] / x_norm), s * (x[1] / x_norm), s * (x[2] / x_norm)};
  // J = (s n^T) [I | -p^]   with p = float-rounded map-frame point (:156-158, manifolds.hpp:63-68)
  out.J[0] = sn[0]; out.J[1] = sn[1]; out.J[2] = sn[2];
  out.J[3] = sn[1] * (-q[2]) + sn[2] * (q[1]);
  out.J[4] = sn[0] * (q[2]) + sn[2] * (-q[0]);
  out.J[5] = sn[0] * (-q[1]) + sn[1] * (q[0]);
  return out;
}

void loam_linearize_impl(const Cloud& src, const Cloud& dst, const KnnGrid& grid, const double* T, int threads,
                         int64_t* knn_idx, double* knn_d2, int32_t* status, double* resid, double* Jout,
                         double* JtJ, double* JtE, int64_t* n_acc) {
  std::vector<LoamPoint> pts(src.n);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256)
  for (long long i = 0; i < (long long)src.n; i++) {
    int64_t idx5[5];
    double d25[5];
    pts[i] = loam_point(src, size_t(i), dst, grid, T, idx5, d25);
    if (knn_idx) for (int j = 0; j < 5; j++) knn_idx[i * 5 + j] = idx5[j];
    if (knn_d2) for (int j = 0; j < 5; j++) knn_d2[i * 5 + j] = d25[j];
  }
  for (int i = 0; i < 36; i++) JtJ[i] = 0;
  for (int i = 0; i < 6; i++) JtE[i] = 0;
  int64_t n = 0;
  for (size_t i = 0; i < src.n; i++) {
    const LoamPoint& p = pts[i];
    if (status) status[i] = p.status;
    if (resid) resid[i] = p.status == 3 ? p.E : 0.0;
    if (Jout) for (int c = 0; c < 6; c++) Jout[i * 6 + c] = p.status == 3 ? p.J[c] : 0.0;
    if (p.status != 3) continue;
    n++;
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) JtJ[r * 6 + c] += p.J[r] * p.J[c];
      JtE[r] += p.J[r] * p.E;
    }
  }
  *n_acc = n;
}
}  // namespace

extern "C" int orc_loam_linearize(const float* src, size_t ns, size_t sstride, const float* dst, size_t nm,
                                  size_t dstride, const double T[16], int threads, int64_t* knn_idx,
                                  double* knn_d2, int32_t* status, double* resid, double* J, double JtJ[36],
                                  double JtE[6], int64_t* n_acc) {
  Cloud s{src, ns, sstride}, d{dst, nm, dstride};
  KnnGrid grid;
  grid.build(d, 1.0f);
  loam_linearize_impl(s, d, grid, T, threads > 0 ? threads : 1, knn_idx, knn_d2, status, resid, J, JtJ, JtE, n_acc);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// A8/A9. LoamRegister::scan2Map (PCR/src/LoamRegister.cpp:99-223)
// ---------------------------------------------------------------------------------------------------
extern "C" int orc_loam_align(const float* src, size_t ns, size_t sstride, const float* dst, size_t nm,
                              size_t dstride, double T[16], int threads, int max_iters,
                              orc_loam_iter_log* logs, int32_t* n_iters, int32_t* converged) {
  Cloud s{src, ns, sstride}, d{dst, nm, dstride};
  if (threads <= 0) threads = 1;
  KnnGrid grid;
  grid.build(d, 1.0f);  // the reference rebuilds its kd-tree here on every call (:110)
  bool conv = false;
  int it = 0;
  const double kPosConverge = double(5e-3f), kRotConverge = double(5e-3f);
  for (it = 0; it < max_iters; it++) {
    double JtJ[36], JtE[6];
    int64_t n;
    loam_linearize_impl(s, d, grid, T, threads, nullptr, nullptr, nullptr, nullptr, nullptr, JtJ, JtE, &n);
    orc_loam_iter_log* lg = logs ? &logs[it] : nullptr;
    if (lg) {
      std::memcpy(lg->T_before, T, sizeof(double) * 16);
      std::memcpy(lg->JtJ, JtJ, sizeof(JtJ));
      std::memcpy(lg->JtE, JtE, sizeof(JtE));
      lg->n = n; lg->converged = 0; lg->pad = 0;
      for (int i = 0; i < 6; i++) lg->x[i] = 0;
    }
    if (n < 6) { it++; break; }  // :173-176 "not enough valid point" -> break, not converged
    double A[6][6], b[6], x[6];
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) A[r][c] = JtJ[r * 6 + c];
      b[r] = -JtE[r];
    }
    ldlt_solve<6>(A, b, x);
    if (lg) for (int i = 0; i < 6; i++) lg->x[i] = x[i];
    double np = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    double nr = std::sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
    if (np <= kPosConverge && nr <= kRotConverge) {  // :202-206 converge before applying x
      conv = true;
      if (lg) lg->converged = 1;
      it++;
      break;
    }
    double E[16], Tn[16];
    orc_se3_exp(x, E);
    for (int c = 0; c < 4; c++)
      for (int r = 0; r < 4; r++) {
        double v = 0;
        for (int k = 0; k < 4; k++) v += E[k * 4 + r] * T[c * 4 + k];
        Tn[c * 4 + r] = v;
      }
    std::memcpy(T, Tn, sizeof(Tn));
  }
  orc_t2se3(T);  // :220
  if (n_iters) *n_iters = it;
  if (converged) *converged = conv ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// test hooks for the dense linear algebra restatements (tests/test_oracle_linalg.py)
// ---------------------------------------------------------------------------------------------------
extern "C" int orc_test_cpqr5x3(const double* A /*5x3 row-major*/, const double* b, double* x) {
  double a[5][3], bb[5], xx[3];
  for (int i = 0; i < 5; i++) { for (int j = 0; j < 3; j++) a[i][j] = A[i * 3 + j]; bb[i] = b[i]; }
  int r = cpqr_solve3<5>(a, bb, xx);
  for (int j = 0; j < 3; j++) x[j] = xx[j];
  return r;
}
extern "C" void orc_test_ldlt6(const double* A, const double* b, double* x) {
  double a[6][6], bb[6], xx[6];
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) a[i][j] = A[i * 6 + j]; bb[i] = b[i]; }
  ldlt_solve<6>(a, bb, xx);
  for (int j = 0; j < 6; j++) x[j] = xx[j];
}
extern "C" void orc_test_svd6(const double* A, const double* b, double* x) {
  double a[6][6], bb[6], xx[6];
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) a[i][j] = A[i * 6 + j]; bb[i] = b[i]; }
  svd_solve<6>(a, bb, xx);
  for (int j = 0; j < 6; j++) x[j] = xx[j];
}
extern "C" void orc_test_eig3(const double* A, double* w, double* V) {
  double a[3][3], ww[3], vv[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = A[i * 3 + j];
  eig_sym3(a, ww, vv);
  for (int i = 0; i < 3; i++) { w[i] = ww[i]; for (int j = 0; j < 3; j++) V[i * 3 + j] = vv[i][j]; }
}
