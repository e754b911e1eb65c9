// PCR::PointCloudRegister and the three config-selected registers with the reference's exact virtual interface,
// implemented over the C ABI of include/pcr_cuda.h (libpcr_cuda.so). Header-only; drop it in place of the reference's
// PCR/include/PCR/*.hpp + PCR/src/*.cpp.
//
// Reference interface mirrored here:
//   PCR/include/PCR/PointCloudRegister.hpp:12-38  (abstract base: scan2Map, getFitnessScore, isConverge, cores)
//   PCR/include/PCR/LoamRegister.hpp:15-87, NdtRegister.hpp:7-26, VgicpRegister.hpp:6-26 (+ initForLC)
// The host types (pt_t = pcl::PointXYZI, pc_t = pcl::PointCloud<pt_t>, pose_t = Eigen::Isometry3d, scalar_t = double)
// come from the reference's <types/basic.hpp>; simpleslam_b200/cpp/standin/ provides minimal stand-ins so that this
// adaptor can be compiled and exercised where PCL / Eigen are not installed.
#pragma once
#include <types/basic.hpp>
#include <pcr_cuda.h>

#include <cstdio>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>

namespace PCR {

class PointCloudRegister {
 protected:
  bool isConverge{false};
  int cores{4};

 public:
  using Ptr = std::shared_ptr<PointCloudRegister>;
  using cPtr = std::shared_ptr<const PointCloudRegister>;
  using PC_Ptr = typename pc_t::Ptr;
  using PC_cPtr = typename pc_t::ConstPtr;

  PointCloudRegister() = default;
  virtual scalar_t getFitnessScore() { return 0; }
  virtual bool scan2Map(const PC_cPtr& src, const PC_cPtr& dst, pose_t& res) = 0;
  virtual ~PointCloudRegister() {}
};

// Common GPU plumbing of the three registers: one pcr_ctx (CUDA stream + device buffers) per instance, like one
// reference register instance per thread (SURVEY.md §8b "Threading").
class CudaRegister : public PointCloudRegister {
 protected:
  pcr_ctx* ctx_{nullptr};
  pcr_params prm_{};
  // optional static-map cache (the reference's VGICP caches by pointer identity, fast_vgicp_impl.hpp:57; loc.cpp feeds the
  // same submap pointer on every call). Off by default = reference semantics of LOAM / NDT: index rebuilt on every call.
  bool cache_target_{false};
  bool pin_target_{false};              // page-lock the cached target's points (pcr_host_register) for its upload
  const void* pinned_points_{nullptr};
  const void* cached_ptr_{nullptr};
  size_t cached_size_{0};
  unsigned long cached_generation_{0}, generation_{0};

  explicit CudaRegister(int method, int cores_ = 4, int device = 0) {
    cores = cores_;
    pcr_default_params(method, &prm_);
    prm_.device = device;
    prm_.cores = cores_;
    open();
  }
  void open() {
    if (ctx_) { pcr_destroy(ctx_); ctx_ = nullptr; }
    const int rc = pcr_create(&prm_, &ctx_);
    if (rc != PCR_OK)  // no CPU fallback: construction fails loudly (north_star)
      throw std::runtime_error(std::string("PCR CUDA register: ") + pcr_last_error(nullptr));
    cached_ptr_ = nullptr;
  }

 public:
  ~CudaRegister() override {
    unpin();
    if (ctx_) pcr_destroy(ctx_);
  }
  CudaRegister(const CudaRegister&) = delete;
  CudaRegister& operator=(const CudaRegister&) = delete;

  // localisation mode: keep the index of an unchanged target (same pointer, size and generation). pin_target: page-lock the
  // target's points before they are uploaded (a static map is large and, with bumpTargetGeneration, uploaded again and again)
  void enableTargetCache(bool on, bool pin_target = false) {
    cache_target_ = on;
    pin_target_ = on && pin_target;
    cached_ptr_ = nullptr;
    if (!pin_target_) unpin();
  }
  void bumpTargetGeneration() { ++generation_; }

  bool scan2Map(const PC_cPtr& src, const PC_cPtr& dst, pose_t& res) override {
    this->isConverge = false;
    int32_t conv = 0;
    int rc;
    const bool hit = cache_target_ && cached_ptr_ == dst.get() && cached_size_ == dst->size() && cached_generation_ == generation_;
    if (!hit) {
      if (pin_target_ && pinned_points_ != dst->points.data()) {
        unpin();
        if (dst->size() && pcr_host_register(dst->points.data(), dst->size() * sizeof(pt_t)) == PCR_OK) pinned_points_ = dst->points.data();
      }
      rc = pcr_set_target(ctx_, dst->points.data(), dst->size(), sizeof(pt_t));
      if (rc != PCR_OK) return fail("set_target");
      cached_ptr_ = dst.get(); cached_size_ = dst->size(); cached_generation_ = generation_;
    }
    // res.matrix().data(): 16 doubles, column-major (Eigen default) == the C ABI's pose layout; refined in place
    rc = pcr_align(ctx_, src->points.data(), src->size(), sizeof(pt_t), res.matrix().data(), &conv);
    if (rc != PCR_OK) return fail("align");
    this->isConverge = conv != 0;
    return this->isConverge;
  }

  bool stats(pcr_stats& s) const { return pcr_get_stats(ctx_, &s) == PCR_OK; }

 protected:
  void unpin() {
    if (pinned_points_) pcr_host_unregister(pinned_points_);
    pinned_points_ = nullptr;
  }

 private:
  bool fail(const char* what) {
    // CUDA / argument failures must not throw through the virtual call: log + false, pose left as it was
    std::fprintf(stderr, "[PCR] %s failed: %s\n", what, pcr_last_error(ctx_));
    cached_ptr_ = nullptr;
    return false;
  }
};

class LoamRegister : public CudaRegister {
 public:
  explicit LoamRegister(int cores_ = 4, int device = 0) : CudaRegister(PCR_LOAM, cores_, device) {}
};

class NdtRegister : public CudaRegister {
 public:
  explicit NdtRegister(int cores_ = 4, int device = 0) : CudaRegister(PCR_NDT, cores_, device) {}
};

class VgicpRegister : public CudaRegister {
 public:
  explicit VgicpRegister(int cores_ = 4, int device = 0) : CudaRegister(PCR_VGICP, cores_, device) {}
  void initForLC() { pcr_vgicp_init_for_lc(ctx_); }  // PCR/src/VgicpRegister.cpp:21-28
  scalar_t getFitnessScore() override {             // PCR/src/VgicpRegister.cpp:42-45
    // fails CLOSED: pcl::Registration::getFitnessScore answers max() when it has nothing to measure, and the loop-closure
    // gate is `fitness < 0.3` (backend/src/LoopClosureManager.cpp:98) — a failed computation must not pass as a perfect match
    double s = std::numeric_limits<double>::max();
    if (pcr_fitness(ctx_, &s) != PCR_OK) {
      std::fprintf(stderr, "[PCR] fitness failed: %s\n", pcr_last_error(ctx_));
      return std::numeric_limits<double>::max();
    }
    return s;
  }
};

// The same factory straight from the reference's configuration object (config::Params::getInstance(), an nlohmann::json):
// reads cfg["frontend"]["pcr"] (frontend/src/LidarOdometry.cpp:44), cfg["cores"] (PointCloudRegister.hpp:30-31, kept for
// compatibility) and the new optional key cfg["gpu"]["device"]. Any json type with operator[], contains()
// and get<T>() works; nothing is included here.
template <class Json>
inline std::shared_ptr<PointCloudRegister> makeRegisterFromConfig(const Json& cfg);

// cfg["frontend"]["pcr"] -> register, as frontend/src/LidarOdometry.cpp:44-53 (unknown string -> runtime_error)
inline PointCloudRegister::Ptr makeRegister(const std::string& pcr_type, int cores = 4, int device = 0) {
  if (pcr_type == "loam") return std::make_shared<LoamRegister>(cores, device);
  if (pcr_type == "ndt") return std::make_shared<NdtRegister>(cores, device);
  if (pcr_type == "vgicp") return std::make_shared<VgicpRegister>(cores, device);
  throw std::runtime_error("such pcr type(" + pcr_type + ") is not exist, please implemented your self!");
}

template <class Json>
inline std::shared_ptr<PointCloudRegister> makeRegisterFromConfig(const Json& cfg) {
  const std::string type = cfg["frontend"]["pcr"].template get<std::string>();
  const int cores = cfg.contains("cores") ? cfg["cores"].template get<int>() : 4;
  int device = 0;
  if (cfg.contains("gpu") && cfg["gpu"].contains("device")) device = cfg["gpu"]["device"].template get<int>();
  return makeRegister(type, cores, device);
}

}  // namespace PCR
