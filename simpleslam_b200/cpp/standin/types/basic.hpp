// Minimal stand-ins for the reference's <types/basic.hpp> (common/types/basic.hpp:14-21 -> PCL + Eigen types), so the
// PCR adaptor can be compiled where PCL / Eigen are absent. Layout-compatible with what the adaptor touches:
//   pt_t   = pcl::PointXYZI  : 32 bytes, x y z pad(=1) | intensity pad pad pad, 16-byte aligned
//   pc_t   = pcl::PointCloud<pt_t> : `points` (contiguous), size(), Ptr / ConstPtr (shared_ptr)
//   pose_t = Eigen::Isometry3d : matrix().data() -> 16 doubles, column-major
// When building inside SimpleSLAM, put the reference's own common/ on the include path instead of this directory.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

namespace standin {
struct alignas(16) PointXYZI {
  float x{0}, y{0}, z{0}, pad{1.0f};
  float intensity{0}, p1{0}, p2{0}, p3{0};
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");

template <typename PointT>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void push_back(const PointT& p) { points.push_back(p); }
};

struct Matrix4d {
  double m[16];  // column-major
  double* data() { return m; }
  const double* data() const { return m; }
  double& operator()(int r, int c) { return m[c * 4 + r]; }
  double operator()(int r, int c) const { return m[c * 4 + r]; }
};
struct Isometry3d {
  Matrix4d mat;
  Isometry3d() { setIdentity(); }
  void setIdentity() { for (int i = 0; i < 16; i++) mat.m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  Matrix4d& matrix() { return mat; }
  const Matrix4d& matrix() const { return mat; }
};
}  // namespace standin

using index_t = size_t;
using scalar_t = double;
using pt_t = standin::PointXYZI;
using pc_t = standin::PointCloud<pt_t>;
using pose_t = standin::Isometry3d;
