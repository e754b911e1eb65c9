// Headless frontend over the C ABI: frontend::LidarOdometry::generateOdom
// (frontend/src/LidarOdometry.cpp:89-246) and frontend::MapManager::{setCurPose, putKeyFrame, updateMap}
// (frontend/src/MapManager.cpp:109-201) without ROS and without threads. The per-frame voxel downsample, the submap
// assembly (transform + concat + downsample, pcr_submap_build) and the registration run on the GPU; the submap never
// leaves the device (it is the register's target). Same synchronous threading model and ascending-index keyframe order
// as simpleslam_b200/frontend.py (see its module docstring); tests/cpp/test_frontend.cpp checks the two against each other.
// Only pose_t::matrix().data() (16 doubles, column-major) and pc_t::points / size() are used from the host types, so
// the header works with the reference's Eigen / PCL types and with the stand-ins alike.
#pragma once
#include <types/basic.hpp>
#include <pcr_cuda.h>
#include <cstdio>

#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace frontend {

namespace m4 {  // column-major 4x4 helpers on raw doubles
inline void identity(double* T) { for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0; }
inline void mul(const double* A, const double* B, double* C) {
  double t[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      double v = 0;
      for (int k = 0; k < 4; k++) v += A[k * 4 + r] * B[c * 4 + k];
      t[c * 4 + r] = v;
    }
  std::memcpy(C, t, sizeof(t));
}
inline void inverse_rigid(const double* T, double* I) {  // [R t; 0 1]^-1 = [R^T  -R^T t]
  double t[16];
  identity(t);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[c * 4 + r] = T[r * 4 + c];
  for (int r = 0; r < 3; r++) t[12 + r] = -(t[0 * 4 + r] * T[12] + t[1 * 4 + r] * T[13] + t[2 * 4 + r] * T[14]);
  std::memcpy(I, t, sizeof(t));
}
inline double dist(const double* A, const double* B) {
  const double dx = A[12] - B[12], dy = A[13] - B[13], dz = A[14] - B[14];
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
}  // namespace m4

// geometry::trans::SixDof2Mobile (common/geometry/trans.hpp:68-86) on a column-major 4x4
inline void sixDof2Mobile(const double* T, double* out) {
  auto R = [&](int r, int c) { return T[c * 4 + r]; };
  double w, v[3];
  const double tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr > 0) {  // Eigen::Quaternion(Matrix3)
    double s = std::sqrt(tr + 1.0);
    w = 0.5 * s;
    s = 0.5 / s;
    v[0] = (R(2, 1) - R(1, 2)) * s; v[1] = (R(0, 2) - R(2, 0)) * s; v[2] = (R(1, 0) - R(0, 1)) * s;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    v[i] = 0.5 * s;
    s = 0.5 / s;
    w = (R(k, j) - R(j, k)) * s;
    v[j] = (R(j, i) + R(i, j)) * s;
    v[k] = (R(k, i) + R(i, k)) * s;
  }
  double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  m4::identity(out);
  out[12] = T[12];
  out[13] = T[13];
  if (n > 0) {  // Eigen::AngleAxis(Quaternion)
    const double angle = 2.0 * std::atan2(n, std::fabs(w));
    if (w < 0) n = -n;
    const double az = v[2] / n;
    if (std::fabs(az) > 0.95) {
      const double c = std::cos(angle), s = std::sin(angle) * std::copysign(1.0, az);
      out[0] = c; out[1] = s; out[4] = -s; out[5] = c;
    }
  }
}

struct KeyFrame {
  pc_t::ConstPtr pc;
  pose_t pose;
};

class MapManager {
 public:
  static constexpr double minKFGap = 1.0;
  static constexpr double mSurroundingKeyframeSearchRadius = 8.0;

  MapManager(pcr_ctx* ctx, float grid_size) : ctx_(ctx), grid_(grid_size) {}

  bool isSubmapEmpty() const { return submap_idx_.empty(); }
  void notifyUpdateMap() { update_requested_ = true; }
  bool updateRequested() const { return update_requested_; }
  size_t submapSize() const { return submap_size_; }
  const std::vector<KeyFrame>& keyframes() const { return keyframes_; }
  const std::vector<size_t>& submapIdx() const { return submap_idx_; }

  void setCurPose(const pose_t& p) {  // MapManager.cpp:109-119
    cur_ = p;
    if (m4::dist(last_.matrix().data(), p.matrix().data()) > minKFGap) {
      last_ = p;
      notifyUpdateMap();
    }
  }
  bool putKeyFrame(const KeyFrame& kf) {  // MapManager.cpp:122-149
    if (keyframes_.empty()) { keyframes_.push_back(kf); return true; }
    double best = 1e300;
    for (const auto& k : keyframes_) {
      const double d = m4::dist(k.pose.matrix().data(), kf.pose.matrix().data());
      best = std::min(best, d * d);
    }
    if (best > minKFGap) { keyframes_.push_back(kf); return true; }  // squared distance vs minKFGap, as the reference (:141)
    return false;
  }
  void updateMap() {  // MapManager.cpp:151-201
    update_requested_ = false;
    if (keyframes_.empty()) return;
    std::vector<const void*> clouds;
    std::vector<size_t> counts;
    std::vector<double> poses;
    std::vector<int64_t> ids;  // keyframe index = cache key of the device copy (a keyframe crosses PCIe once)
    submap_idx_.clear();
    for (size_t i = 0; i < keyframes_.size(); i++) {
      const double d = m4::dist(keyframes_[i].pose.matrix().data(), cur_.matrix().data());
      if (d * d < mSurroundingKeyframeSearchRadius * mSurroundingKeyframeSearchRadius) {
        submap_idx_.push_back(i);
        clouds.push_back(keyframes_[i].pc->points.data());
        counts.push_back(keyframes_[i].pc->size());
        ids.push_back(int64_t(i));
        const double* T = keyframes_[i].pose.matrix().data();
        poses.insert(poses.end(), T, T + 16);
      }
    }
    size_t m = 0;
    if (pcr_submap_build(ctx_, clouds.data(), counts.data(), ids.data(), clouds.size(), sizeof(pt_t), poses.data(), grid_, nullptr, 0, &m) != PCR_OK)
      throw std::runtime_error(std::string("pcr_submap_build: ") + pcr_last_error(ctx_));
    submap_size_ = m;
  }

 private:
  pcr_ctx* ctx_;
  float grid_;
  std::vector<KeyFrame> keyframes_;
  std::vector<size_t> submap_idx_;
  size_t submap_size_{0};
  pose_t cur_, last_;
  bool update_requested_{false};
};

class LidarOdometry {
 public:
  explicit LidarOdometry(const std::string& pcr_type, float grid_size = 0.5f, int device = 0) : grid_(grid_size) {
    int method;
    if (pcr_type == "loam") method = PCR_LOAM;
    else if (pcr_type == "ndt") method = PCR_NDT;
    else if (pcr_type == "vgicp") method = PCR_VGICP;
    else throw std::runtime_error("such pcr type(" + pcr_type + ") is not exist, please implemented your self!");
    pcr_params p;
    pcr_default_params(method, &p);
    p.device = device;
    if (pcr_create(&p, &ctx_) != PCR_OK) throw std::runtime_error(std::string("PCR CUDA register: ") + pcr_last_error(nullptr));
    map_.reset(new MapManager(ctx_, grid_size));
  }
  ~LidarOdometry() { map_.reset(); if (ctx_) pcr_destroy(ctx_); }
  LidarOdometry(const LidarOdometry&) = delete;
  LidarOdometry& operator=(const LidarOdometry&) = delete;

  MapManager& map() { return *map_; }
  bool lastConverged() const { return last_conv_; }

  // one frame (LidarOdometry.cpp:89-246); local_odom may be null. Returns the global pose of the frame.
  pose_t generateOdom(const pc_t::ConstPtr& scan, double stamp, const pose_t* local_odom) {
    pose_t init;  // mRelocPose = identity
    if (local_odom && odom2map_init_) {
      m4::mul(odom2map_.matrix().data(), local_odom->matrix().data(), init.matrix().data());
    } else if (global_.size() >= 2) {  // Frontend::getClosestItem: frames arrive in stamp order -> the newest entry, used iff cidx > 0
      size_t cidx = global_.size() - 1;
      double m = std::fabs(stamp - global_.back().first);
      for (size_t i = global_.size() - 1; i-- > 0;) {
        const double t = std::fabs(global_[i].first - stamp);
        if (t < m) { m = t; cidx = i; } else break;
      }
      if (cidx > 0) init = global_[cidx].second;
    }
    last_conv_ = true;
    if (!map_->isSubmapEmpty()) {
      // mVoxelGrid.filter (:170-171) + mPcr->scan2Map (:184): one call, the downsampled scan stays on the device
      int32_t conv = 0;
      if (pcr_downsample_align(ctx_, scan->points.data(), scan->size(), sizeof(pt_t), grid_, init.matrix().data(), &conv, nullptr) != PCR_OK) {
        std::fprintf(stderr, "[PCR] %s\n", pcr_last_error(ctx_));
        conv = 0;  // scan2Map failure is not fatal (:184-199)
      }
      last_conv_ = conv != 0;
    }
    pose_t mob;
    sixDof2Mobile(init.matrix().data(), mob.matrix().data());
    map_->setCurPose(mob);
    KeyFrame kf{scan, mob};
    if (map_->isSubmapEmpty()) {
      map_->putKeyFrame(kf);
      map_->notifyUpdateMap();
    } else {
      const double* t = mob.matrix().data();
      const double dx = t[12] - last_pos_[0], dy = t[13] - last_pos_[1], dz = t[14] - last_pos_[2];
      if (std::sqrt(dx * dx + dy * dy + dz * dz) > MapManager::minKFGap) {
        map_->putKeyFrame(kf);
        last_pos_[0] = t[12]; last_pos_[1] = t[13]; last_pos_[2] = t[14];
      }
    }
    global_.emplace_back(stamp, mob);
    if (local_odom) {
      double inv[16];
      m4::inverse_rigid(local_odom->matrix().data(), inv);
      m4::mul(mob.matrix().data(), inv, odom2map_.matrix().data());
      odom2map_init_ = true;
    }
    if (map_->updateRequested()) map_->updateMap();  // the map thread, run synchronously
    return mob;
  }

 private:
  pcr_ctx* ctx_{nullptr};
  float grid_;
  std::unique_ptr<MapManager> map_;
  std::vector<std::pair<double, pose_t>> global_;
  pose_t odom2map_;
  bool odom2map_init_{false};
  bool last_conv_{true};
  double last_pos_[3] = {0, 0, 0};
};

}  // namespace frontend
