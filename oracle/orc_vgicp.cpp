// ORACLE — TEST INFRASTRUCTURE ONLY (see pcr_oracle.h). PARITY UNPINNED (no reference fixtures).
// CPU restatement of fast_gicp::FastVGICP as configured by PCR/src/VgicpRegister.cpp:
//   V1 FastGICP::calculate_covariances        third_parties/pclomp/src/fast_gicp_impl.hpp:241-298
//   V2 GaussianVoxelMap (ADDITIVE)            third_parties/pclomp/src/pclomp/fast_vgicp_voxel.hpp:105-174
//   V3 FastVGICP::update_correspondences      third_parties/pclomp/src/fast_vgicp_impl.hpp:73-116
//   V4 FastVGICP::linearize / compute_error   ...:119-204
//   V5 LsqRegistration (LM / GN)              third_parties/pclomp/src/lsq_registration_impl.hpp:53-172
//   V6 pcl::Registration::getFitnessScore     (un-vendored)
// Summation order convention: ascending source index (the reference's per-thread partial sums are
// schedule-dependent, SURVEY §7 hard part 7).
#include "pcr_oracle.h"
#include "orc_common.hpp"
#include "orc_linalg.hpp"
#include <omp.h>
#include <map>
#include <array>
#include <unordered_map>

using namespace orc;

// ---------------------------------------------------------------------------------------------------
// V1. k-NN covariances with PLANE regularisation: C = U diag(1,1,1e-3) V^T of cov = N N^T / k.
// For a symmetric PSD matrix the SVD equals the eigen-decomposition (U = V column-wise wherever the
// singular value is nonzero), so C = I - (1 - 1e-3) n n^T with n the eigenvector of the smallest
// eigenvalue; we form it through the full U diag V^T product to mirror the reference's arithmetic.
// ---------------------------------------------------------------------------------------------------
static void gicp_cov_from_neighbors(const Cloud& c, const int64_t* idx, int found, int k, double* cov9) {
  // neighbors (4 x k, double), columns beyond `found` stay uninitialised in the reference; k <= n is assumed
  double mean[3] = {0, 0, 0};
  for (int j = 0; j < found; j++) {
    const float* p = c.at(size_t(idx[j]));
    for (int a = 0; a < 3; a++) mean[a] += double(p[a]);
  }
  for (int a = 0; a < 3; a++) mean[a] /= double(k);
  double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < found; j++) {
    const float* p = c.at(size_t(idx[j]));
    double d[3] = {double(p[0]) - mean[0], double(p[1]) - mean[1], double(p[2]) - mean[2]};
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) cov[r][q] += d[r] * d[q];
  }
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) cov[r][q] /= double(k);
  double w[3], V[3][3];
  eig_sym3(cov, w, V);  // ascending: column 0 = smallest
  const double vals[3] = {1e-3, 1.0, 1.0};
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) {
      double v = 0;
      for (int e = 2; e >= 0; e--) v += (V[r][e] * vals[e]) * V[q][e];  // descending singular order
      cov9[r * 3 + q] = v;
    }
}

extern "C" int orc_gicp_covariances(const float* pts, size_t n, size_t stride_f, int k, int threads, double* covs_out,
                                    int64_t* knn_idx_out) {
  Cloud c{pts, n, stride_f};
  if (threads <= 0) threads = 1;
  KnnGrid grid;
  grid.build(c, 0.5f);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
  for (long long i = 0; i < (long long)n; i++) {
    std::vector<int64_t> idx(k, -1);
    std::vector<float> d2(k, 0.f);
    const float* q = c.at(size_t(i));
    int found = grid.knn<float>(q, k, idx.data(), d2.data());
    if (knn_idx_out) for (int j = 0; j < k; j++) knn_idx_out[i * k + j] = j < found ? idx[j] : -1;
    gicp_cov_from_neighbors(c, idx.data(), found, k, covs_out + i * 9);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
namespace {
struct Voxel { int n; double mean[3]; double cov[9]; };
struct Coord { int32_t x, y, z; bool operator<(const Coord& o) const { return z != o.z ? z < o.z : (y != o.y ? y < o.y : x < o.x); } };
}  // namespace

struct orc_vgicp {
  std::vector<float> dst_copy;
  size_t nm;
  double resolution;
  std::vector<double> target_covs;
  std::map<Coord, Voxel> voxels;
};

static inline Coord voxel_coord(const double* x, double res) {
  // fast_vgicp_voxel.hpp:158-160  (x / res - 0.5).floor().cast<int>()
  return Coord{int32_t(std::floor(x[0] / res - 0.5)), int32_t(std::floor(x[1] / res - 0.5)), int32_t(std::floor(x[2] / res - 0.5))};
}

extern "C" orc_vgicp* orc_vgicp_create(const float* dst, size_t nm, size_t dstride, double resolution, int k, int threads,
                                       const double* target_covs) {
  orc_vgicp* h = new orc_vgicp();
  h->nm = nm;
  h->resolution = resolution;
  h->dst_copy.resize(nm * 4);
  Cloud d{dst, nm, dstride};
  for (size_t i = 0; i < nm; i++) {
    const float* p = d.at(i);
    h->dst_copy[i * 4] = p[0]; h->dst_copy[i * 4 + 1] = p[1]; h->dst_copy[i * 4 + 2] = p[2]; h->dst_copy[i * 4 + 3] = 1.f;
  }
  h->target_covs.resize(nm * 9);
  if (target_covs) std::copy(target_covs, target_covs + nm * 9, h->target_covs.begin());
  else orc_gicp_covariances(h->dst_copy.data(), nm, 4, k, threads, h->target_covs.data(), nullptr);
  // create_voxelmap, ADDITIVE (fast_vgicp_voxel.hpp:129-156, 105-122), serial in index order
  for (size_t i = 0; i < nm; i++) {
    double x[3] = {double(h->dst_copy[i * 4]), double(h->dst_copy[i * 4 + 1]), double(h->dst_copy[i * 4 + 2])};
    Coord co = voxel_coord(x, resolution);
    auto it = h->voxels.find(co);
    if (it == h->voxels.end()) {
      Voxel v{};
      it = h->voxels.insert({co, v}).first;
    }
    Voxel& v = it->second;
    v.n++;
    for (int a = 0; a < 3; a++) v.mean[a] += x[a];
    for (int a = 0; a < 9; a++) v.cov[a] += h->target_covs[i * 9 + a];
  }
  for (auto& kv : h->voxels) {
    Voxel& v = kv.second;
    for (int a = 0; a < 3; a++) v.mean[a] /= v.n;
    for (int a = 0; a < 9; a++) v.cov[a] /= v.n;
  }
  return h;
}
extern "C" void orc_vgicp_destroy(orc_vgicp* h) { delete h; }
extern "C" size_t orc_vgicp_num_voxels(const orc_vgicp* h) { return h->voxels.size(); }
extern "C" void orc_vgicp_get_voxels(const orc_vgicp* h, int32_t* coords, int32_t* npts, double* mean, double* cov) {
  size_t i = 0;
  for (auto& kv : h->voxels) {
    if (coords) { coords[i * 3] = kv.first.x; coords[i * 3 + 1] = kv.first.y; coords[i * 3 + 2] = kv.first.z; }
    if (npts) npts[i] = kv.second.n;
    if (mean) for (int a = 0; a < 3; a++) mean[i * 3 + a] = kv.second.mean[a];
    if (cov) for (int a = 0; a < 9; a++) cov[i * 9 + a] = kv.second.cov[a];
    i++;
  }
}

namespace {
struct PairTerm { bool valid; double cost; double H[36]; double b[6]; };

// One source point against the voxel found with T0 (DIRECT1), Mahalanobis from T0, error at Ti.
inline void vgicp_point(const orc_vgicp* h, const float* p, const double* covA, const double* T0, const double* Ti,
                        bool want_Hb, PairTerm& out) {
  out.valid = false;
  double mean_A[3] = {double(p[0]), double(p[1]), double(p[2])};
  double tA0[3];
  transform_f64(T0, mean_A, tA0);
  Coord co = voxel_coord(tA0, h->resolution);
  auto it = h->voxels.find(co);
  if (it == h->voxels.end()) return;
  const Voxel& v = it->second;
  // RCR = cov_B + T cov_A T^T ; (3,3) = 1 ; inverse ; (3,3) = 0  -> upper-left 3x3 inverse
  double R[3][3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = T0[c * 4 + r];
  double RC[3][3], RCR[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) RC[r][c] = (R[r][0] * covA[0 * 3 + c] + R[r][1] * covA[1 * 3 + c]) + R[r][2] * covA[2 * 3 + c];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) RCR[r][c] = v.cov[r * 3 + c] + ((RC[r][0] * R[c][0] + RC[r][1] * R[c][1]) + RC[r][2] * R[c][2]);
  double M[3][3];
  inv3(RCR, M);
  double tA[3];
  transform_f64(Ti, mean_A, tA);
  double e[3] = {v.mean[0] - tA[0], v.mean[1] - tA[1], v.mean[2] - tA[2]};
  double Me[3];
  for (int r = 0; r < 3; r++) Me[r] = (M[r][0] * e[0] + M[r][1] * e[1]) + M[r][2] * e[2];
  double w = std::sqrt(double(v.n));
  out.valid = true;
  out.cost = w * ((e[0] * Me[0] + e[1] * Me[1]) + e[2] * Me[2]);
  if (!want_Hb) return;
  // J (3x6) = [ skew(T p) | -I ]
  double J[3][6] = {{0, -tA[2], tA[1], -1, 0, 0}, {tA[2], 0, -tA[0], 0, -1, 0}, {-tA[1], tA[0], 0, 0, 0, -1}};
  double MJ[3][6];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 6; c++) MJ[r][c] = (M[r][0] * J[0][c] + M[r][1] * J[1][c]) + M[r][2] * J[2][c];
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) out.H[r * 6 + c] = w * ((J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c]) + J[2][r] * MJ[2][c]);
    out.b[r] = w * ((J[0][r] * Me[0] + J[1][r] * Me[1]) + J[2][r] * Me[2]);
  }
}

double vgicp_eval(const orc_vgicp* h, const Cloud& src, const double* src_covs, const double* T0, const double* Ti, int threads,
                  double* H, double* b, int64_t* n_corr) {
  bool want = H != nullptr && b != nullptr;
  std::vector<PairTerm> terms(src.n);
#pragma omp parallel for num_threads(threads) schedule(guided, 8)
  for (long long i = 0; i < (long long)src.n; i++) vgicp_point(h, src.at(size_t(i)), src_covs + i * 9, T0, Ti, want, terms[i]);
  double sum = 0;
  int64_t n = 0;
  if (want) { for (int i = 0; i < 36; i++) H[i] = 0; for (int i = 0; i < 6; i++) b[i] = 0; }
  for (size_t i = 0; i < src.n; i++) {
    if (!terms[i].valid) continue;
    n++;
    sum += terms[i].cost;
    if (want) { for (int k = 0; k < 36; k++) H[k] += terms[i].H[k]; for (int k = 0; k < 6; k++) b[k] += terms[i].b[k]; }
  }
  if (n_corr) *n_corr = n;
  return sum;
}

// so3_exp (so3/so3.hpp:58-77) -> Quaterniond::toRotationMatrix()
void so3_exp_matrix(const double* omega, double R[3][3]) {
  double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double theta = std::sqrt(theta_sq);
    double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  double w = real_factor, x = imag_factor * omega[0], y = imag_factor * omega[1], z = imag_factor * omega[2];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}

// delta (rot from d[0:3], trans d[3:6]) ; xi = delta * x0
void make_delta(const double* d, double D[16]) {
  double R[3][3];
  so3_exp_matrix(d, R);
  for (int i = 0; i < 16; i++) D[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) D[c * 4 + r] = R[r][c]; D[12 + r] = d[3 + r]; }
}
void matmul4(const double* A, const double* B, double* C) {
  double t[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      double v = 0;
      for (int k = 0; k < 4; k++) v += A[k * 4 + r] * B[c * 4 + k];
      t[c * 4 + r] = v;
    }
  for (int i = 0; i < 16; i++) C[i] = t[i];
}
bool is_converged(const double* D, double rot_eps, double trans_eps) {
  double m = 0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) m = std::max(m, 1.0 / rot_eps * std::fabs(D[c * 4 + r] - (r == c ? 1.0 : 0.0)));
    m = std::max(m, 1.0 / trans_eps * std::fabs(D[12 + r]));
  }
  return m < 1;
}
}  // namespace

extern "C" double orc_vgicp_linearize(const orc_vgicp* h, const float* src, size_t ns, size_t sstride, const double* src_covs,
                                      const double T[16], int threads, double H[36], double b[6], int64_t* n_corr) {
  Cloud s{src, ns, sstride};
  return vgicp_eval(h, s, src_covs, T, T, threads > 0 ? threads : 1, H, b, n_corr);
}
extern "C" double orc_vgicp_error(const orc_vgicp* h, const float* src, size_t ns, size_t sstride, const double* src_covs,
                                  const double T0[16], const double Ti[16], int threads) {
  Cloud s{src, ns, sstride};
  return vgicp_eval(h, s, src_covs, T0, Ti, threads > 0 ? threads : 1, nullptr, nullptr, nullptr);
}

extern "C" int orc_vgicp_align(const orc_vgicp* h, const float* src, size_t ns, size_t sstride, const double* src_covs_in,
                               const double Tguess[16], int threads, int optimizer, int max_iterations, double rot_eps,
                               double trans_eps, orc_vgicp_result* out) {
  Cloud s{src, ns, sstride};
  if (threads <= 0) threads = 1;
  std::vector<double> covs_own;
  const double* src_covs = src_covs_in;
  if (!src_covs) {
    covs_own.resize(ns * 9);
    orc_gicp_covariances(src, ns, sstride, 20, threads, covs_own.data(), nullptr);
    src_covs = covs_own.data();
  }
  // VgicpRegister.cpp:36: guess = res.matrix().cast<float>(); lsq_registration_impl.hpp:54 x0 = Isometry3d(guess.cast<double>())
  double x0[16];
  for (int i = 0; i < 16; i++) x0[i] = double(static_cast<float>(Tguess[i]));
  double lm_lambda = -1.0;
  bool converged = false;
  int nr_iterations = 0, n_lin = 0, n_err = 0;
  const int lm_max_iterations = 10;
  const double lm_init_lambda_factor = 1e-9;
  for (int i = 0; i < max_iterations && !converged; i++) {
    nr_iterations = i;
    double delta[16];
    double H[36], b[6];
    double y0 = vgicp_eval(h, s, src_covs, x0, x0, threads, H, b, nullptr);
    n_lin++;
    bool ok = false;
    if (optimizer == 1) {
      // step_gn :106-123
      double A[6][6], rhs[6], d[6];
      for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) A[r][c] = H[r * 6 + c]; rhs[r] = -b[r]; }
      ldlt_solve<6>(A, rhs, d);
      make_delta(d, delta);
      matmul4(delta, x0, x0);
      ok = true;
    } else {
      // step_lm :125-172
      if (lm_lambda < 0.0) {
        double mx = 0;
        for (int r = 0; r < 6; r++) mx = std::max(mx, std::fabs(H[r * 6 + r]));
        lm_lambda = lm_init_lambda_factor * mx;
      }
      double nu = 2.0;
      for (int li = 0; li < lm_max_iterations; li++) {
        double A[6][6], rhs[6], d[6];
        for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) A[r][c] = H[r * 6 + c] + (r == c ? lm_lambda : 0.0); rhs[r] = -b[r]; }
        ldlt_solve<6>(A, rhs, d);
        make_delta(d, delta);
        double xi[16];
        matmul4(delta, x0, xi);
        double yi = vgicp_eval(h, s, src_covs, x0, xi, threads, nullptr, nullptr, nullptr);
        n_err++;
        double den = 0;
        for (int r = 0; r < 6; r++) den += d[r] * (lm_lambda * d[r] - b[r]);
        double rho = (y0 - yi) / den;
        if (rho < 0) {
          if (is_converged(delta, rot_eps, trans_eps)) { ok = true; break; }
          lm_lambda = nu * lm_lambda;
          nu = 2 * nu;
          continue;
        }
        std::memcpy(x0, xi, sizeof(xi));
        lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
        ok = true;
        break;
      }
    }
    if (!ok) break;  // "lm not converged!!"
    converged = is_converged(delta, rot_eps, trans_eps);
  }
  for (int i = 0; i < 16; i++) out->T[i] = double(static_cast<float>(x0[i]));  // final_transformation_ = x0.cast<float>()
  out->converged = converged ? 1 : 0;
  out->nr_iterations = nr_iterations;
  out->n_linearize = n_lin;
  out->n_error_evals = n_err;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// V6 pcl::Registration::getFitnessScore(max_range): float transform, 1-NN (FLANN float metric), mean d2
// ---------------------------------------------------------------------------------------------------
extern "C" double orc_fitness(const float* src, size_t ns, size_t sstride, const float* dst, size_t nm, size_t dstride,
                              const double T[16], double max_range, int threads) {
  Cloud s{src, ns, sstride}, d{dst, nm, dstride};
  if (threads <= 0) threads = 1;
  KnnGrid grid;
  grid.build(d, 1.0f);
  float Tf[16];
  for (int i = 0; i < 16; i++) Tf[i] = static_cast<float>(T[i]);
  std::vector<float> d2s(ns, -1.f);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256)
  for (long long i = 0; i < (long long)ns; i++) {
    float q[3];
    transform_f32(Tf, s.at(size_t(i)), q);
    int64_t idx;
    float d2;
    if (grid.knn<float>(q, 1, &idx, &d2) == 1) d2s[i] = d2;
  }
  double sum = 0;
  int nr = 0;
  for (size_t i = 0; i < ns; i++)
    if (d2s[i] >= 0 && double(d2s[i]) <= max_range) { sum += double(d2s[i]); nr++; }
  if (nr > 0) return sum / nr;
  return std::numeric_limits<double>::max();
}
