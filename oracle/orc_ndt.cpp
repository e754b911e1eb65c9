// ORACLE — TEST INFRASTRUCTURE ONLY (see pcr_oracle.h). PARITY UNPINNED (no reference fixtures).
// CPU restatement of pclomp::NormalDistributionsTransform as configured by PCR/src/NdtRegister.cpp:
//   N1 VoxelGridCovariance::applyFilter   third_parties/pclomp/src/voxel_grid_covariance_omp_impl.hpp:49-370
//   N2 getNeighborhoodAtPoint{7,1,26}     ...:374-442
//   N3 computeDerivatives / updateDerivatives (float path)  third_parties/pclomp/src/ndt_omp_impl.hpp:180-537
//   N4 computeHessian / updateHessian (double path)         ...:541-645
//   N5 computeTransformation + More-Thuente line search     ...:81-171, 649-932
//   N6 KDTREE radius search over voxel centroids            pclomp/voxel_grid_covariance_omp.h:476-505
// Documented deviations: (i) the initial Euler angles are taken from the guess's linear part directly
// (Eigen's Affine `.rotation()` runs a float SVD polar decomposition first — differs by ~1e-7 rad for an
// orthonormal guess); (ii) Eigen 3.3 eulerAngles() range convention.
#include "pcr_oracle.h"
#include "orc_common.hpp"
#include "orc_linalg.hpp"
#include <omp.h>
#include <unordered_map>
#include <map>
#include <cstdio>

using namespace orc;

namespace {
struct Leaf {
  int32_t key;
  int nr_points;
  double mean[3];
  double cov[3][3];
  double icov[3][3];
  float centroid[3];
  bool in_centroid_cloud;  // had >= min_points (pushed to voxel_centroids_, even if later rejected)
};
}  // namespace

struct orc_ndt {
  VoxelGridSpec g;
  float leaf_size[3];
  std::vector<Leaf> leaves;  // ascending key (std::map order)
  std::unordered_map<int32_t, int32_t> lookup;
  std::vector<float> centroid_pts;  // stride 4, for KDTREE mode
  std::vector<int32_t> centroid_leaf;
  KnnGrid cgrid;
  double d1, d2, d3;
  float resolution;
};

static void gauss_params(double resolution, double outlier_ratio, double& d1, double& d2, double& d3) {
  // ndt_omp_impl.hpp:86-93
  double c1 = 10 * (1 - outlier_ratio);
  double c2 = outlier_ratio / std::pow(resolution, 3);
  d3 = -std::log(c2);
  d1 = -std::log(c1 + c2) - d3;
  d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / d1);
}

extern "C" orc_ndt* orc_ndt_create(const float* dst, size_t nm, size_t dstride, float resolution) {
  orc_ndt* h = new orc_ndt();
  Cloud c{dst, nm, dstride};
  h->resolution = resolution;
  h->g = voxel_grid_spec(c, resolution);
  for (int a = 0; a < 3; a++) h->leaf_size[a] = resolution;
  gauss_params(double(resolution), 0.55, h->d1, h->d2, h->d3);
  if (h->g.overflow) return h;  // :79-84 output cleared, no leaves
  const int min_points = 6;
  const double eig_mult = 0.01;
  struct Acc { int n = 0; double sum[3] = {0, 0, 0}; double cov[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; float csum[3] = {0, 0, 0}; };
  std::map<int32_t, Acc> acc;  // Leaf() starts cov_ at Identity (voxel_grid_covariance_omp.h:107)
  for (size_t i = 0; i < nm; i++) {
    const float* p = c.at(i);
    int32_t key = voxel_key(h->g, p);
    Acc& a = acc[key];
    double pd[3] = {p[0], p[1], p[2]};
    for (int r = 0; r < 3; r++) {
      a.sum[r] += pd[r];
      for (int q = 0; q < 3; q++) a.cov[r][q] += pd[r] * pd[q];
      a.csum[r] += p[r];
    }
    a.n++;
  }
  for (auto& kv : acc) {
    Acc& a = kv.second;
    Leaf L{};
    L.key = kv.first;
    L.nr_points = a.n;
    for (int r = 0; r < 3; r++) {
      L.centroid[r] = a.csum[r] / static_cast<float>(a.n);
      L.mean[r] = a.sum[r] / a.n;
    }
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) { L.cov[r][q] = (r == q); L.icov[r][q] = 0; }
    L.in_centroid_cloud = false;
    if (a.n >= min_points) {
      L.in_centroid_cloud = true;
      // :329-330 single pass covariance
      double n = a.n;
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++)
          L.cov[r][q] = (a.cov[r][q] - 2 * (a.sum[r] * L.mean[q])) / n + L.mean[r] * L.mean[q];
      double f = (n - 1.0) / n;
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) L.cov[r][q] *= f;
      double w[3], V[3][3];
      eig_sym3(L.cov, w, V);
      if (w[0] < 0 || w[1] < 0 || w[2] <= 0) {
        L.nr_points = -1;  // :337-341
      } else {
        double mn = eig_mult * w[2];
        if (w[0] < mn) {
          w[0] = mn;
          if (w[1] < mn) w[1] = mn;
          double Vi[3][3];
          inv3(V, Vi);
          for (int r = 0; r < 3; r++)
            for (int q = 0; q < 3; q++) {
              double v = 0;
              for (int k = 0; k < 3; k++) v += (V[r][k] * w[k]) * Vi[k][q];
              L.cov[r][q] = v;
            }
        }
        inv3(L.cov, L.icov);
        double mx = -1e300, mi = 1e300;
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) { mx = std::max(mx, L.icov[r][q]); mi = std::min(mi, L.icov[r][q]); }
        if (mx == std::numeric_limits<double>::infinity() || mi == -std::numeric_limits<double>::infinity()) L.nr_points = -1;
      }
    }
    h->lookup[L.key] = static_cast<int32_t>(h->leaves.size());
    h->leaves.push_back(L);
  }
  for (size_t i = 0; i < h->leaves.size(); i++) {
    const Leaf& L = h->leaves[i];
    if (!L.in_centroid_cloud) continue;
    h->centroid_pts.push_back(L.centroid[0]);
    h->centroid_pts.push_back(L.centroid[1]);
    h->centroid_pts.push_back(L.centroid[2]);
    h->centroid_pts.push_back(1.f);
    h->centroid_leaf.push_back(static_cast<int32_t>(i));
  }
  Cloud cc{h->centroid_pts.data(), h->centroid_leaf.size(), 4};
  h->cgrid.build(cc, resolution);
  return h;
}

extern "C" void orc_ndt_destroy(orc_ndt* h) { delete h; }

extern "C" size_t orc_ndt_num_leaves(const orc_ndt* h, int32_t grid_out[9]) {
  if (grid_out)
    for (int a = 0; a < 3; a++) { grid_out[a] = h->g.min_b[a]; grid_out[3 + a] = h->g.max_b[a]; grid_out[6 + a] = h->g.div_b[a]; }
  return h->leaves.size();
}

extern "C" void orc_ndt_get_leaves(const orc_ndt* h, int32_t* keys, int32_t* npts, double* mean, double* cov, double* icov) {
  for (size_t i = 0; i < h->leaves.size(); i++) {
    const Leaf& L = h->leaves[i];
    if (keys) keys[i] = L.key;
    if (npts) npts[i] = L.nr_points;
    for (int r = 0; r < 3; r++) {
      if (mean) mean[i * 3 + r] = L.mean[r];
      for (int q = 0; q < 3; q++) {
        if (cov) cov[i * 9 + r * 3 + q] = L.cov[r][q];
        if (icov) icov[i * 9 + r * 3 + q] = L.icov[r][q];
      }
    }
  }
}

namespace {
// Eigen::AngleAxisf(angle, Unit{X,Y,Z}).toRotationMatrix() in float
void angle_axis_f32(float angle, int axis, float R[3][3]) {
  float ax[3] = {0, 0, 0};
  ax[axis] = 1.f;
  float s = std::sin(angle), c = std::cos(angle);
  float sin_axis[3] = {s * ax[0], s * ax[1], s * ax[2]};
  float cos1_axis[3] = {(1.f - c) * ax[0], (1.f - c) * ax[1], (1.f - c) * ax[2]};
  float tmp;
  tmp = cos1_axis[0] * ax[1]; R[0][1] = tmp - sin_axis[2]; R[1][0] = tmp + sin_axis[2];
  tmp = cos1_axis[0] * ax[2]; R[0][2] = tmp + sin_axis[1]; R[2][0] = tmp - sin_axis[1];
  tmp = cos1_axis[1] * ax[2]; R[1][2] = tmp - sin_axis[0]; R[2][1] = tmp + sin_axis[0];
  for (int i = 0; i < 3; i++) R[i][i] = cos1_axis[i] * ax[i] + c;
}
void matmul3_f32(const float A[3][3], const float B[3][3], float C[3][3]) {
  float t[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i][j] = (A[i][0] * B[0][j] + A[i][1] * B[1][j]) + A[i][2] * B[2][j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i][j] = t[i][j];
}
// (Translation3f(x,y,z) * AngleAxisf(rx,X) * AngleAxisf(ry,Y) * AngleAxisf(rz,Z)).matrix()  ndt_omp_impl.hpp:827-830
void pose_vec_to_matrix_f32(const double p[6], float M[16]) {
  float Rx[3][3], Ry[3][3], Rz[3][3], L[3][3];
  angle_axis_f32(static_cast<float>(p[3]), 0, Rx);
  angle_axis_f32(static_cast<float>(p[4]), 1, Ry);
  angle_axis_f32(static_cast<float>(p[5]), 2, Rz);
  matmul3_f32(Rx, Ry, L);
  matmul3_f32(L, Rz, L);
  for (int i = 0; i < 16; i++) M[i] = 0.f;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) M[c * 4 + r] = L[r][c];
    M[12 + r] = static_cast<float>(p[r]);
  }
  M[15] = 1.f;
}

struct AngleTables {
  float j_ang[8][3];
  float h_ang[15][3];
  double j_ang_d[8][3];
  double h_ang_d[15][3];
};
// computeAngleDerivatives ndt_omp_impl.hpp:289-395 (float table row 6 = d1 carries +sy, double table -sy)
void angle_tables(const double p[6], AngleTables& t) {
  double cx, cy, cz, sx, sy, sz;
  if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
  if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
  if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
  double j[8][3] = {{(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
                    {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
                    {(-sy * cz), sy * sz, cy},
                    {sx * cy * cz, (-sx * cy * sz), sx * sy},
                    {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
                    {(-cy * sz), (-cy * cz), 0},
                    {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
                    {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
  double h[15][3] = {{(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},
                     {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},
                     {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},
                     {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},
                     {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},
                     {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},
                     {(-cy * cz), (cy * sz), (-sy)},
                     {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},
                     {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},
                     {(sy * sz), (sy * cz), 0},
                     {(-sx * cy * sz), (-sx * cy * cz), 0},
                     {(cx * cy * sz), (cx * cy * cz), 0},
                     {(-cy * cz), (cy * sz), 0},
                     {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},
                     {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 3; c++) { t.j_ang_d[r][c] = j[r][c]; t.j_ang[r][c] = static_cast<float>(j[r][c]); }
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 3; c++) { t.h_ang_d[r][c] = h[r][c]; t.h_ang[r][c] = static_cast<float>(h[r][c]); }
  t.h_ang[6][2] = static_cast<float>(sy);  // :383 float-path quirk (+sy)
}

// neighbourhoods (N2 / N6). Returns leaf indices in reference order.
int neighbourhood(const orc_ndt* h, const float* pt, int search, int32_t* out /* cap 64 */) {
  int n = 0;
  if (h->g.overflow) return 0;
  if (search == 0) {
    // KDTREE: radius search over centroids, sorted by distance, strict d2 < r^2 (FLANN RadiusResultSet)
    float r2 = h->resolution * h->resolution;
    struct Hit { float d; int32_t i; };
    std::vector<Hit> hits;
    const KnnGrid& G = h->cgrid;
    int c0[3], reach[3];
    for (int a = 0; a < 3; a++) {
      c0[a] = int(std::floor((double(pt[a]) - double(G.origin[a])) / double(G.cell)));
      reach[a] = int(std::ceil(double(h->resolution) / double(G.cell))) + 1;
    }
    for (int z = std::max(c0[2] - reach[2], 0); z <= std::min(c0[2] + reach[2], G.dim[2] - 1); z++)
      for (int y = std::max(c0[1] - reach[1], 0); y <= std::min(c0[1] + reach[1], G.dim[1] - 1); y++)
        for (int x = std::max(c0[0] - reach[0], 0); x <= std::min(c0[0] + reach[0], G.dim[0] - 1); x++) {
          size_t cid = size_t(x) + size_t(G.dim[0]) * (size_t(y) + size_t(G.dim[1]) * size_t(z));
          for (uint32_t s = G.start[cid]; s < G.start[cid + 1]; s++) {
            uint32_t ci = G.order[s];
            float d = KnnGrid::dist2<float>(pt, G.c.at(ci));
            if (d < r2) hits.push_back({d, int32_t(ci)});
          }
        }
    std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); });
    for (auto& hh : hits) if (n < 64) out[n++] = h->centroid_leaf[hh.i];
    return n;
  }
  // DIRECT*: voxel_grid_covariance_omp_impl.hpp:374-404
  int ijk[3];
  for (int a = 0; a < 3; a++) ijk[a] = static_cast<int>(std::floor(pt[a] / h->leaf_size[a]));
  static const int off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  int offs[27][3];
  int noff = 0;
  if (search == 2) { noff = 7; for (int i = 0; i < 7; i++) for (int a = 0; a < 3; a++) offs[i][a] = off7[i][a]; }
  else if (search == 3) { noff = 1; offs[0][0] = offs[0][1] = offs[0][2] = 0; }
  else {
    // pcl::getAllNeighborCellIndices(): 27 cells (dx outer ... dz inner) minus the last (1,1,1)?  PCL builds
    // all 3^3 combinations then drops the centre by "conservativeResize(3, 26)" — recalled; the oracle uses
    // the 26 non-centre offsets followed order x-major. DIRECT26 is not the configured path.
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
          if (dx == 0 && dy == 0 && dz == 0) continue;
          offs[noff][0] = dx; offs[noff][1] = dy; offs[noff][2] = dz; noff++;
        }
  }
  for (int ni = 0; ni < noff; ni++) {
    bool inside = true;
    for (int a = 0; a < 3; a++) {
      int d2min = h->g.min_b[a] - ijk[a], d2max = h->g.max_b[a] - ijk[a];
      if (!(d2min <= offs[ni][a] && d2max >= offs[ni][a])) inside = false;
    }
    if (!inside) continue;
    int key = 0;
    for (int a = 0; a < 3; a++) key += (ijk[a] + offs[ni][a] - h->g.min_b[a]) * h->g.mul[a];
    auto it = h->lookup.find(key);
    if (it != h->lookup.end() && h->leaves[it->second].nr_points >= 6) out[n++] = it->second;
  }
  return n;
}

struct PointDeriv { double score; double g[6]; double H[6][6]; };

// float path: computePointDerivatives (:399-440) + updateDerivatives (:485-537)
inline void point_derivs_f32(const orc_ndt* h, const AngleTables& tb, const float* x_orig, const float* x_trans_pt,
                             const int32_t* nb, int nnb, bool compute_hessian, PointDeriv& out) {
  out.score = 0;
  for (int i = 0; i < 6; i++) { out.g[i] = 0; for (int j = 0; j < 6; j++) out.H[i][j] = 0; }
  if (nnb == 0) return;
  // x4 = float(double(x_pt)) = x_pt
  float x4[3] = {x_orig[0], x_orig[1], x_orig[2]};
  float J[4][6];
  for (int r = 0; r < 4; r++) for (int c = 0; c < 6; c++) J[r][c] = 0.f;
  J[0][0] = J[1][1] = J[2][2] = 1.f;
  float xj[8];
  for (int r = 0; r < 8; r++) xj[r] = (tb.j_ang[r][0] * x4[0] + tb.j_ang[r][1] * x4[1]) + tb.j_ang[r][2] * x4[2];
  J[1][3] = xj[0]; J[2][3] = xj[1]; J[0][4] = xj[2]; J[1][4] = xj[3]; J[2][4] = xj[4]; J[0][5] = xj[5]; J[1][5] = xj[6]; J[2][5] = xj[7];
  float HE[24][6];
  for (int r = 0; r < 24; r++) for (int c = 0; c < 6; c++) HE[r][c] = 0.f;
  if (compute_hessian) {
    float xh[15];
    for (int r = 0; r < 15; r++) xh[r] = (tb.h_ang[r][0] * x4[0] + tb.h_ang[r][1] * x4[1]) + tb.h_ang[r][2] * x4[2];
    float a[4] = {0, xh[0], xh[1], 0}, b[4] = {0, xh[2], xh[3], 0}, c[4] = {0, xh[4], xh[5], 0};
    float d[4] = {xh[6], xh[7], xh[8], 0}, e[4] = {xh[9], xh[10], xh[11], 0}, f[4] = {xh[12], xh[13], xh[14], 0};
    for (int k = 0; k < 4; k++) {
      HE[12 + k][3] = a[k]; HE[16 + k][3] = b[k]; HE[20 + k][3] = c[k];
      HE[12 + k][4] = b[k]; HE[16 + k][4] = d[k]; HE[20 + k][4] = e[k];
      HE[12 + k][5] = c[k]; HE[16 + k][5] = e[k]; HE[20 + k][5] = f[k];
    }
  }
  const float gauss_d2 = static_cast<float>(h->d2);
  for (int ni = 0; ni < nnb; ni++) {
    const Leaf& L = h->leaves[nb[ni]];
    float xt[4];
    for (int k = 0; k < 3; k++) xt[k] = static_cast<float>(double(x_trans_pt[k]) - L.mean[k]);
    xt[3] = 0.f;
    float C[4][4];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) C[r][c] = (r < 3 && c < 3) ? static_cast<float>(L.icov[r][c]) : 0.f;
    float xC[4];
    for (int c = 0; c < 4; c++) xC[c] = ((xt[0] * C[0][c] + xt[1] * C[1][c]) + xt[2] * C[2][c]) + xt[3] * C[3][c];
    float q = ((xt[0] * xC[0] + xt[1] * xC[1]) + xt[2] * xC[2]) + xt[3] * xC[3];
    float e_x_cov_x = std::exp(-gauss_d2 * q * 0.5f);
    float score_inc = static_cast<float>(-h->d1 * double(e_x_cov_x));
    e_x_cov_x = gauss_d2 * e_x_cov_x;
    if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;  // returns 0: no score either
    e_x_cov_x = static_cast<float>(double(e_x_cov_x) * h->d1);
    float CJ[4][6];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 6; c++) CJ[r][c] = ((C[r][0] * J[0][c] + C[r][1] * J[1][c]) + C[r][2] * J[2][c]) + C[r][3] * J[3][c];
    float xCJ[6];
    for (int c = 0; c < 6; c++) xCJ[c] = ((xt[0] * CJ[0][c] + xt[1] * CJ[1][c]) + xt[2] * CJ[2][c]) + xt[3] * CJ[3][c];
    for (int c = 0; c < 6; c++) out.g[c] += double(e_x_cov_x * xCJ[c]);
    if (compute_hessian) {
      float JCJ[6][6];
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) JCJ[r][c] = ((J[0][r] * CJ[0][c] + J[1][r] * CJ[1][c]) + J[2][r] * CJ[2][c]) + J[3][r] * CJ[3][c];
      for (int i = 0; i < 6; i++) {
        float xH[6];
        for (int j = 0; j < 6; j++)
          xH[j] = ((xC[0] * HE[i * 4 + 0][j] + xC[1] * HE[i * 4 + 1][j]) + xC[2] * HE[i * 4 + 2][j]) + xC[3] * HE[i * 4 + 3][j];
        for (int j = 0; j < 6; j++)
          out.H[i][j] += double(e_x_cov_x * (-gauss_d2 * xCJ[i] * xCJ[j] + xH[j] + JCJ[j][i]));
      }
    }
    out.score += double(score_inc);
  }
}

double derivatives_impl(const orc_ndt* h, const Cloud& src, const float* Tf, const double* p, int search,
                        bool compute_hessian, int threads, double* g, double* H, int32_t* nb_count) {
  AngleTables tb;
  angle_tables(p, tb);
  std::vector<PointDeriv> per(src.n);
#pragma omp parallel for num_threads(threads) schedule(guided, 8)
  for (long long i = 0; i < (long long)src.n; i++) {
    const float* po = src.at(size_t(i));
    float pt[3];
    transform_f32(Tf, po, pt);
    int32_t nb[64];
    int nnb = neighbourhood(h, pt, search, nb);
    if (nb_count) nb_count[i] = nnb;
    point_derivs_f32(h, tb, po, pt, nb, nnb, compute_hessian, per[i]);
  }
  double score = 0;
  for (int i = 0; i < 6; i++) { g[i] = 0; for (int j = 0; j < 6; j++) H[i * 6 + j] = 0; }
  for (size_t i = 0; i < src.n; i++) {  // :278-282 serial sum in index order
    score += per[i].score;
    for (int r = 0; r < 6; r++) { g[r] += per[i].g[r]; for (int c = 0; c < 6; c++) H[r * 6 + c] += per[i].H[r][c]; }
  }
  return score;
}

void hessian_impl(const orc_ndt* h, const Cloud& src, const float* Tf, const double* p, int search, double* H) {
  AngleTables tb;
  angle_tables(p, tb);
  for (int i = 0; i < 36; i++) H[i] = 0;
  for (size_t idx = 0; idx < src.n; idx++) {
    const float* po = src.at(idx);
    float pt[3];
    transform_f32(Tf, po, pt);
    int32_t nb[64];
    int nnb = neighbourhood(h, pt, search, nb);
    if (!nnb) continue;
    double x[3] = {po[0], po[1], po[2]};
    double J[3][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}};
    auto dot = [&](const double* r) { return x[0] * r[0] + x[1] * r[1] + x[2] * r[2]; };
    J[1][3] = dot(tb.j_ang_d[0]); J[2][3] = dot(tb.j_ang_d[1]); J[0][4] = dot(tb.j_ang_d[2]); J[1][4] = dot(tb.j_ang_d[3]);
    J[2][4] = dot(tb.j_ang_d[4]); J[0][5] = dot(tb.j_ang_d[5]); J[1][5] = dot(tb.j_ang_d[6]); J[2][5] = dot(tb.j_ang_d[7]);
    double HE[18][6];
    for (int r = 0; r < 18; r++) for (int c = 0; c < 6; c++) HE[r][c] = 0;
    double a[3] = {0, dot(tb.h_ang_d[0]), dot(tb.h_ang_d[1])}, b[3] = {0, dot(tb.h_ang_d[2]), dot(tb.h_ang_d[3])};
    double c[3] = {0, dot(tb.h_ang_d[4]), dot(tb.h_ang_d[5])};
    double d[3] = {dot(tb.h_ang_d[6]), dot(tb.h_ang_d[7]), dot(tb.h_ang_d[8])};
    double e[3] = {dot(tb.h_ang_d[9]), dot(tb.h_ang_d[10]), dot(tb.h_ang_d[11])};
    double f[3] = {dot(tb.h_ang_d[12]), dot(tb.h_ang_d[13]), dot(tb.h_ang_d[14])};
    for (int k = 0; k < 3; k++) {
      HE[9 + k][3] = a[k]; HE[12 + k][3] = b[k]; HE[15 + k][3] = c[k];
      HE[9 + k][4] = b[k]; HE[12 + k][4] = d[k]; HE[15 + k][4] = e[k];
      HE[9 + k][5] = c[k]; HE[12 + k][5] = e[k]; HE[15 + k][5] = f[k];
    }
    for (int ni = 0; ni < nnb; ni++) {
      const Leaf& L = h->leaves[nb[ni]];
      double xt[3];
      for (int k = 0; k < 3; k++) xt[k] = double(pt[k]) - L.mean[k];
      double Cx[3];
      for (int r = 0; r < 3; r++) Cx[r] = L.icov[r][0] * xt[0] + L.icov[r][1] * xt[1] + L.icov[r][2] * xt[2];
      double e_x_cov_x = h->d2 * std::exp(-h->d2 * (xt[0] * Cx[0] + xt[1] * Cx[1] + xt[2] * Cx[2]) / 2);
      if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;
      e_x_cov_x *= h->d1;
      for (int i = 0; i < 6; i++) {
        double cov_dxd_pi[3];
        for (int r = 0; r < 3; r++) cov_dxd_pi[r] = L.icov[r][0] * J[0][i] + L.icov[r][1] * J[1][i] + L.icov[r][2] * J[2][i];
        for (int j = 0; j < 6; j++) {
          double cj[3], ch[3];
          for (int r = 0; r < 3; r++) {
            cj[r] = L.icov[r][0] * J[0][j] + L.icov[r][1] * J[1][j] + L.icov[r][2] * J[2][j];
            ch[r] = L.icov[r][0] * HE[3 * i + 0][j] + L.icov[r][1] * HE[3 * i + 1][j] + L.icov[r][2] * HE[3 * i + 2][j];
          }
          double xd = xt[0] * cov_dxd_pi[0] + xt[1] * cov_dxd_pi[1] + xt[2] * cov_dxd_pi[2];
          double xcj = xt[0] * cj[0] + xt[1] * cj[1] + xt[2] * cj[2];
          double xch = xt[0] * ch[0] + xt[1] * ch[1] + xt[2] * ch[2];
          double jd = J[0][j] * cov_dxd_pi[0] + J[1][j] * cov_dxd_pi[1] + J[2][j] * cov_dxd_pi[2];
          H[i * 6 + j] += e_x_cov_x * (-h->d2 * xd * xcj + xch + jd);
        }
      }
    }
  }
}

// More-Thuente helpers ndt_omp_impl.hpp:649-769
bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  else if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  else return true;
}
double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
    else return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
    else return a_s;
  } else if (std::fabs(g_t) <= std::fabs(g_l)) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
    else return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
  } else {
    double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}
inline double psiMT(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
inline double dpsiMT(double g_a, double g_0, double mu) { return g_a - mu * g_0; }
}  // namespace

// Eigen 3.3 Matrix3f::eulerAngles(0,1,2)
extern "C" void orc_euler_xyz_f32(const float R[9], float res[3]) {
  auto m = [&](int r, int c) { return R[r * 3 + c]; };
  const int i = 0, j = 1, k = 2;
  res[0] = std::atan2(m(j, k), m(k, k));
  float c2 = std::sqrt(m(i, i) * m(i, i) + m(i, j) * m(i, j));
  if (res[0] > 0.f) {  // !odd && res[0] > 0
    if (res[0] > 0.f) res[0] -= static_cast<float>(M_PI); else res[0] += static_cast<float>(M_PI);
    res[1] = std::atan2(-m(i, k), -c2);
  } else {
    res[1] = std::atan2(-m(i, k), c2);
  }
  float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * m(k, i) - c1 * m(j, i), c1 * m(j, j) - s1 * m(k, j));
  res[0] = -res[0]; res[1] = -res[1]; res[2] = -res[2];
}

extern "C" double orc_ndt_derivatives(const orc_ndt* h, const float* src, size_t ns, size_t sstride, const double p[6],
                                      int search, int compute_hessian, int threads, double g[6], double H[36],
                                      int32_t* nb_count) {
  float Tf[16];
  pose_vec_to_matrix_f32(p, Tf);
  Cloud s{src, ns, sstride};
  return derivatives_impl(h, s, Tf, p, search, compute_hessian != 0, threads > 0 ? threads : 1, g, H, nb_count);
}

extern "C" double orc_ndt_derivatives_T(const orc_ndt* h, const float* src, size_t ns, size_t sstride, const float Tf[16],
                                        const double p[6], int search, int compute_hessian, int threads, double g[6],
                                        double H[36], int32_t* nb_count) {
  Cloud s{src, ns, sstride};
  return derivatives_impl(h, s, Tf, p, search, compute_hessian != 0, threads > 0 ? threads : 1, g, H, nb_count);
}

extern "C" void orc_ndt_hessian(const orc_ndt* h, const float* src, size_t ns, size_t sstride, const double p[6], int search,
                                double H[36]) {
  float Tf[16];
  pose_vec_to_matrix_f32(p, Tf);
  Cloud s{src, ns, sstride};
  hessian_impl(h, s, Tf, p, search, H);
}

extern "C" int orc_ndt_align(const orc_ndt* h, const float* src, size_t ns, size_t sstride, const double Tguess[16], int search,
                             int threads, int max_iterations, double trans_eps, double step_size, orc_ndt_result* out) {
  Cloud s{src, ns, sstride};
  if (threads <= 0) threads = 1;
  // NdtRegister.cpp:27: guess = res.matrix().cast<float>()
  float guess[16], final_T[16];
  bool guess_is_identity = true;
  for (int i = 0; i < 16; i++) {
    guess[i] = static_cast<float>(Tguess[i]);
    if (guess[i] != ((i % 5 == 0) ? 1.f : 0.f)) guess_is_identity = false;
  }
  for (int i = 0; i < 16; i++) final_T[i] = (i % 5 == 0) ? 1.f : 0.f;  // align(): final_transformation_ = Identity
  float cloudT[16];
  std::memcpy(cloudT, final_T, sizeof(cloudT));
  if (!guess_is_identity) { std::memcpy(final_T, guess, sizeof(guess)); std::memcpy(cloudT, guess, sizeof(guess)); }
  // :103-111 p = [translation, eulerAngles(0,1,2)]
  double p[6];
  {
    float Rm[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rm[r * 3 + c] = final_T[c * 4 + r];
    float eul[3];
    orc_euler_xyz_f32(Rm, eul);
    p[0] = final_T[12]; p[1] = final_T[13]; p[2] = final_T[14];
    p[3] = eul[0]; p[4] = eul[1]; p[5] = eul[2];
  }
  int n_deriv = 0, n_hess = 0;
  double g[6], H[36];
  double score = derivatives_impl(h, s, cloudT, p, search, true, threads, g, H, nullptr);
  n_deriv++;
  int nr_iterations = 0;
  bool converged = false;
  while (!converged) {
    double Hm[6][6], b[6], delta_p[6];
    for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) Hm[r][c] = H[r * 6 + c]; b[r] = -g[r]; }
    svd_solve<6>(Hm, b, delta_p);
    double delta_p_norm = 0;
    for (int i = 0; i < 6; i++) delta_p_norm += delta_p[i] * delta_p[i];
    delta_p_norm = std::sqrt(delta_p_norm);
    if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
      converged = delta_p_norm == delta_p_norm;
      break;
    }
    for (int i = 0; i < 6; i++) delta_p[i] /= delta_p_norm;
    // ---- computeStepLengthMT(p, delta_p, delta_p_norm, step_size, trans_eps/2, score, g, H, cloud) :773-932
    double a_t_final;
    {
      double* step_dir = delta_p;
      double step_init = delta_p_norm, step_max = step_size, step_min = trans_eps / 2;
      double phi_0 = -score;
      double d_phi_0 = 0;
      for (int i = 0; i < 6; i++) d_phi_0 += g[i] * step_dir[i];
      d_phi_0 = -d_phi_0;
      bool bail = false;
      if (d_phi_0 >= 0) {
        if (d_phi_0 == 0) { a_t_final = 0; bail = true; }
        else { d_phi_0 *= -1; for (int i = 0; i < 6; i++) step_dir[i] *= -1; }
      }
      if (!bail) {
        const int max_step_iterations = 10;
        int step_iterations = 0;
        const double mu = 1.e-4, nu = 0.9;
        double a_l = 0, a_u = 0;
        double f_l = psiMT(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsiMT(d_phi_0, d_phi_0, mu);
        double f_u = psiMT(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsiMT(d_phi_0, d_phi_0, mu);
        bool interval_converged = (step_max - step_min) < 0, open_interval = true;
        double a_t = step_init;
        a_t = std::min(a_t, step_max);
        a_t = std::max(a_t, step_min);
        double x_t[6];
        for (int i = 0; i < 6; i++) x_t[i] = p[i] + step_dir[i] * a_t;
        pose_vec_to_matrix_f32(x_t, final_T);
        score = derivatives_impl(h, s, final_T, x_t, search, true, threads, g, H, nullptr);
        n_deriv++;
        double phi_t = -score, d_phi_t = 0;
        for (int i = 0; i < 6; i++) d_phi_t += g[i] * step_dir[i];
        d_phi_t = -d_phi_t;
        double psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
        while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
          if (open_interval) a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
          else a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
          a_t = std::min(a_t, step_max);
          a_t = std::max(a_t, step_min);
          for (int i = 0; i < 6; i++) x_t[i] = p[i] + step_dir[i] * a_t;
          pose_vec_to_matrix_f32(x_t, final_T);
          score = derivatives_impl(h, s, final_T, x_t, search, false, threads, g, H, nullptr);
          n_deriv++;
          phi_t = -score;
          d_phi_t = 0;
          for (int i = 0; i < 6; i++) d_phi_t += g[i] * step_dir[i];
          d_phi_t = -d_phi_t;
          psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
          d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
          if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
            open_interval = false;
            f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
            g_l = g_l + mu * d_phi_0;
            f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
            g_u = g_u + mu * d_phi_0;
          }
          if (open_interval) interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
          else interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
          step_iterations++;
        }
        if (step_iterations) { hessian_impl(h, s, final_T, x_t, search, H); n_hess++; }
        a_t_final = a_t;
      }
    }
    delta_p_norm = a_t_final;
    for (int i = 0; i < 6; i++) { delta_p[i] *= delta_p_norm; p[i] += delta_p[i]; }
    if (nr_iterations > max_iterations || (nr_iterations && (std::fabs(delta_p_norm) < trans_eps))) converged = true;
    nr_iterations++;
  }
  for (int i = 0; i < 16; i++) out->T[i] = double(final_T[i]);
  out->trans_probability = score / double(ns);
  out->converged = converged ? 1 : 0;
  out->nr_iterations = nr_iterations;
  out->n_derivative_evals = n_deriv;
  out->n_hessian_evals = n_hess;
  for (int i = 0; i < 6; i++) out->p_final[i] = p[i];
  return 0;
}
