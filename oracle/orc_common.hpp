// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.hpp header). PARITY UNPINNED by reference tests:
// the reference ships no fixtures for this path; kNN is cross-checked against the
// reference's own vendored nanoflann compiled into oracle/_ref (see ref_nanoflann.cpp).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <limits>

namespace orc {

// A cloud is a strided array of floats; xyz at float offset 0..2, intensity at offset 4 when
// stride_f >= 5 (pcl::PointXYZI layout: x y z pad | intensity pad pad pad = 8 floats).
struct Cloud {
  const float* p;
  size_t n;
  size_t stride_f;
  const float* at(size_t i) const { return p + i * stride_f; }
  float intensity(size_t i) const { return stride_f >= 5 ? p[i * stride_f + 4] : 0.f; }
};

// pcl::transformPointCloud(in, out, Matrix4f) — PCL <= 1.9 scalar form
// ((m00*x + m01*y) + m02*z) + m03, all float.  M is column-major 4x4 float.
inline void transform_f32(const float* M, const float* p, float* o) {
  for (int r = 0; r < 3; r++) o[r] = ((M[r] * p[0] + M[4 + r] * p[1]) + M[8 + r] * p[2]) + M[12 + r];
}

// Eigen::Isometry3d * Vector4d(x,y,z,w=1): ((R0*x + R1*y) + R2*z) + t*w, double
// (/root/reference/PCR/src/LoamRegister.cpp:128-129, fast_vgicp_impl.hpp:85).  T column-major 4x4.
inline void transform_f64(const double* T, const double* p, double* o) {
  for (int r = 0; r < 3; r++) o[r] = ((T[r] * p[0] + T[4 + r] * p[1]) + T[8 + r] * p[2]) + T[12 + r] * 1.0;
}

// ------------------------------------------------------------------------------------------------
// Integer voxel grid shared by pcl::VoxelGrid (A1) and pclomp::VoxelGridCovariance (N1):
// /root/reference/common/pcp/pcp.hpp:191-210 (in-tree restatement of PCL's key math) and
// /root/reference/third_parties/pclomp/src/voxel_grid_covariance_omp_impl.hpp:67-103,218-223.
// ------------------------------------------------------------------------------------------------
struct VoxelGridSpec {
  float leaf, inv_leaf;
  float min_p[3], max_p[3];
  int min_b[3], max_b[3], div_b[3], mul[3];
  bool overflow;  // dx*dy*dz > INT32_MAX
};

inline VoxelGridSpec voxel_grid_spec(const Cloud& c, float leaf) {
  VoxelGridSpec g{};
  g.leaf = leaf;
  g.inv_leaf = 1.0f / leaf;
  for (int a = 0; a < 3; a++) { g.min_p[a] = std::numeric_limits<float>::max(); g.max_p[a] = -std::numeric_limits<float>::max(); }
  for (size_t i = 0; i < c.n; i++) {
    const float* p = c.at(i);
    for (int a = 0; a < 3; a++) { g.min_p[a] = std::min(g.min_p[a], p[a]); g.max_p[a] = std::max(g.max_p[a], p[a]); }
  }
  int64_t d[3];
  for (int a = 0; a < 3; a++) d[a] = static_cast<int64_t>((g.max_p[a] - g.min_p[a]) * g.inv_leaf) + 1;
  g.overflow = c.n == 0 || (d[0] * d[1] * d[2]) > static_cast<int64_t>(std::numeric_limits<int32_t>::max());
  for (int a = 0; a < 3; a++) {
    g.min_b[a] = static_cast<int>(std::floor(g.min_p[a] * g.inv_leaf));
    g.max_b[a] = static_cast<int>(std::floor(g.max_p[a] * g.inv_leaf));
    g.div_b[a] = g.max_b[a] - g.min_b[a] + 1;
  }
  g.mul[0] = 1; g.mul[1] = g.div_b[0]; g.mul[2] = g.div_b[0] * g.div_b[1];
  return g;
}

inline int32_t voxel_key(const VoxelGridSpec& g, const float* p) {
  int ijk0 = static_cast<int>(std::floor(p[0] * g.inv_leaf) - static_cast<float>(g.min_b[0]));
  int ijk1 = static_cast<int>(std::floor(p[1] * g.inv_leaf) - static_cast<float>(g.min_b[1]));
  int ijk2 = static_cast<int>(std::floor(p[2] * g.inv_leaf) - static_cast<float>(g.min_b[2]));
  return ijk0 * g.mul[0] + ijk1 * g.mul[1] + ijk2 * g.mul[2];
}

// ------------------------------------------------------------------------------------------------
// Exact kNN with the oracle's documented tie-break (d2, index) ascending.
//  metric double: nanoflann L2_Simple<double>::evalMetric (nanoflann.hpp:522-533): sum_{x,y,z} (q-p)^2
//                 with q double (already float-rounded by the caller), p float promoted to double.
//  metric float : FLANN L2_Simple<float> behind pcl::search::KdTree: float diff,
//                 float square, float accumulate, x->y->z.
// The search structure is a uniform hash grid with ring expansion; exactness does not depend on it
// (cross-checked against brute force and against the reference's nanoflann in tests).
// ------------------------------------------------------------------------------------------------
struct KnnGrid {
  Cloud c;
  float cell;
  float origin[3];
  int dim[3];
  std::vector<uint32_t> start;  // dim0*dim1*dim2 + 1
  std::vector<uint32_t> order;  // point indices sorted by cell (ascending index inside a cell)

  void build(const Cloud& cloud, float cell_size) {
    c = cloud; cell = cell_size;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (size_t i = 0; i < c.n; i++) {
      const float* p = c.at(i);
      for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
    }
    if (c.n == 0) { for (int a = 0; a < 3; a++) { mn[a] = 0; mx[a] = 0; } }
    // adapt the cell so the dense table stays small
    for (;;) {
      double cells = 1;
      for (int a = 0; a < 3; a++) {
        origin[a] = mn[a];
        dim[a] = static_cast<int>(std::floor((mx[a] - mn[a]) / cell)) + 1;
        cells *= dim[a];
      }
      if (cells <= 6.4e7) break;
      cell *= 2.f;
    }
    size_t ncell = size_t(dim[0]) * dim[1] * dim[2];
    start.assign(ncell + 1, 0);
    std::vector<uint32_t> key(c.n);
    for (size_t i = 0; i < c.n; i++) {
      key[i] = cell_of(c.at(i));
      start[key[i] + 1]++;
    }
    for (size_t k = 0; k < ncell; k++) start[k + 1] += start[k];
    order.resize(c.n);
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    for (size_t i = 0; i < c.n; i++) order[fill[key[i]]++] = static_cast<uint32_t>(i);
  }
  int coord(float v, int a) const {
    double f = std::floor((double(v) - double(origin[a])) / double(cell));
    int q = f < 0 ? 0 : (f > double(dim[a] - 1) ? dim[a] - 1 : int(f));
    return q;
  }
  uint32_t cell_of(const float* p) const {
    return uint32_t(coord(p[0], 0)) + uint32_t(dim[0]) * (uint32_t(coord(p[1], 1)) + uint32_t(dim[1]) * uint32_t(coord(p[2], 2)));
  }

  template <typename D>
  static inline D dist2(const D* q, const float* p);

  // k nearest: results ascending by (d2, idx). Returns count found (min(k, n)).
  template <typename D>
  int knn(const D* q, int k, int64_t* idx, D* d2) const {
    int cnt = 0;
    auto consider = [&](uint32_t pi) {
      D d = dist2<D>(q, c.at(pi));
      if (cnt == k && !(d < d2[k - 1] || (d == d2[k - 1] && int64_t(pi) < idx[k - 1]))) return;
      int i = (cnt < k) ? cnt : k - 1;
      while (i > 0 && (d2[i - 1] > d || (d2[i - 1] == d && idx[i - 1] > int64_t(pi)))) {
        d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; i--;
      }
      d2[i] = d; idx[i] = pi;
      if (cnt < k) cnt++;
    };
    // unclamped cell coordinates of the query
    double qc[3];
    int c0[3];
    for (int a = 0; a < 3; a++) {
      qc[a] = (double(q[a]) - double(origin[a])) / double(cell);
      double f = std::floor(qc[a]);
      c0[a] = f < -1e9 ? -1000000000 : (f > 1e9 ? 1000000000 : int(f));
    }
    int maxring = 0;
    for (int a = 0; a < 3; a++) maxring = std::max(maxring, std::max(std::abs(c0[a]), std::abs(dim[a] - 1 - c0[a])));
    int rstart = 0;  // rings closer than the grid's bounding box are empty
    for (int a = 0; a < 3; a++) rstart = std::max(rstart, std::max(-c0[a], c0[a] - (dim[a] - 1)));
    for (int r = rstart; r <= maxring; r++) {
      // visit the shell of Chebyshev radius r
      int lo[3], hi[3];
      for (int a = 0; a < 3; a++) { lo[a] = c0[a] - r; hi[a] = c0[a] + r; }
      for (int z = std::max(lo[2], 0); z <= std::min(hi[2], dim[2] - 1); z++)
        for (int y = std::max(lo[1], 0); y <= std::min(hi[1], dim[1] - 1); y++) {
          bool edge_zy = (z == lo[2] || z == hi[2] || y == lo[1] || y == hi[1]);
          if (edge_zy) {
            for (int x = std::max(lo[0], 0); x <= std::min(hi[0], dim[0] - 1); x++) {
              size_t cid = size_t(x) + size_t(dim[0]) * (size_t(y) + size_t(dim[1]) * size_t(z));
              for (uint32_t s = start[cid]; s < start[cid + 1]; s++) consider(order[s]);
            }
          } else {
            for (int xi = 0; xi < 2; xi++) {
              int x = xi == 0 ? lo[0] : hi[0];
              if (xi == 1 && hi[0] == lo[0]) break;
              if (x < 0 || x >= dim[0]) continue;
              size_t cid = size_t(x) + size_t(dim[0]) * (size_t(y) + size_t(dim[1]) * size_t(z));
              for (uint32_t s = start[cid]; s < start[cid + 1]; s++) consider(order[s]);
            }
          }
        }
      if (cnt == k) {
        // every unvisited point lies outside the cube of half-width (r + frac) cells around q
        double margin = 1e300;
        for (int a = 0; a < 3; a++) {
          double fl = qc[a] - std::floor(qc[a]);
          margin = std::min(margin, std::min(fl, 1.0 - fl));
        }
        double reach = (double(r) + margin) * double(cell) * (1.0 - 1e-6);
        if (reach > 0 && double(d2[k - 1]) < reach * reach) break;
      }
    }
    return cnt;
  }
};

template <>
inline double KnnGrid::dist2<double>(const double* q, const float* p) {
  double r = 0;
  for (int a = 0; a < 3; a++) { double d = q[a] - double(p[a]); r += d * d; }
  return r;
}
template <>
inline float KnnGrid::dist2<float>(const float* q, const float* p) {
  float r = 0;
  for (int a = 0; a < 3; a++) { float d = q[a] - p[a]; r += d * d; }
  return r;
}

template <typename D>
inline int knn_brute(const Cloud& c, const D* q, int k, int64_t* idx, D* d2) {
  int cnt = 0;
  for (size_t pi = 0; pi < c.n; pi++) {
    D d = KnnGrid::dist2<D>(q, c.at(pi));
    if (cnt == k && !(d < d2[k - 1])) continue;  // ascending scan: equal d2 keeps the lower index
    int i = (cnt < k) ? cnt : k - 1;
    while (i > 0 && d2[i - 1] > d) { d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; i--; }
    d2[i] = d; idx[i] = int64_t(pi);
    if (cnt < k) cnt++;
  }
  return cnt;
}

}  // namespace orc
