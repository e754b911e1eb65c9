// Host side of an upload from host memory — always for PAGEABLE sources (what a pcl::PointCloud is), for pinned sources when
// the register was given at least 8 `cores`: `cores` host threads — the reference's `cores`
// setting (PCR/include/PCR/PointCloudRegister.hpp:30-31, config/params.json:5), which there sizes the OpenMP teams of the
// registration itself — turn the caller's AoS records (pcl::PointXYZI: 32 bytes) into the 16-byte float4 records the
// kernels read, into pinned staging memory, chunk by chunk; each chunk crosses PCIe while the next one is packed. A plain
// cudaMemcpyAsync from pageable memory is staged by the driver on one thread (~11 GB/s measured) and moves the 16 bytes of
// padding / ring / time of every record as well; a pinned source crosses PCIe at ~50 GB/s as it is, which 8 or more packing
// threads beat by halving the bytes (api.cu: use_host_pack). Device sources never come here.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__) || defined(__x86_64__)
#include <emmintrin.h>
#define PCR_HOSTPACK_SSE2 1
#endif

namespace pcr {

// the host twin of pack_kernel (voxel.cu): (x, y, z, intensity | 0)
inline void host_pack_slice(const unsigned char* src, size_t n, size_t stride, float* out) {
  if (stride == 32) {
#ifdef PCR_HOSTPACK_SSE2
    if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
      // non-temporal stores: the staging buffer is only read by the copy engine, and a cached store would first READ the
      // line it overwrites (16 more bytes of memory traffic per record)
      for (size_t i = 0; i < n; i++) {
        const float* r = reinterpret_cast<const float*>(src + i * 32);
        const __m128 a = _mm_loadu_ps(r);                                  // x y z pad
        const __m128 b = _mm_load_ss(r + 4);                               // intensity 0 0 0
        const __m128 zi = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 2, 2));   // z z i i
        _mm_stream_ps(out + i * 4, _mm_shuffle_ps(a, zi, _MM_SHUFFLE(2, 0, 1, 0)));  // x y z i
      }
      _mm_sfence();
      return;
    }
#endif
    for (size_t i = 0; i < n; i++) {
      const float* r = reinterpret_cast<const float*>(src + i * 32);
      float* o = out + i * 4;
      o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[4];
    }
  } else if (stride == 16) {
    for (size_t i = 0; i < n; i++) {
      const float* r = reinterpret_cast<const float*>(src + i * 16);
      float* o = out + i * 4;
      o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = 0.f;
    }
  } else {
    for (size_t i = 0; i < n; i++) {
      float r[5] = {0, 0, 0, 0, 0};
      std::memcpy(r, src + i * stride, stride >= 20 ? 20 : 12);
      float* o = out + i * 4;
      o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = stride >= 20 ? r[4] : 0.f;
    }
  }
}

// A few persistent worker threads; pack() cuts the records into equal slices (the caller takes one) and returns when all
// of them are done.
class HostPacker {
 public:
  explicit HostPacker(int threads) {
    const int hw = int(std::thread::hardware_concurrency());
    n_ = threads < 1 ? 1 : threads;
    if (hw > 0 && n_ > hw) n_ = hw;
    if (n_ > 64) n_ = 64;
    for (int w = 1; w < n_; w++) workers_.emplace_back([this, w] { loop(w); });
  }
  ~HostPacker() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  HostPacker(const HostPacker&) = delete;
  HostPacker& operator=(const HostPacker&) = delete;
  int threads() const { return n_; }

  void pack(const unsigned char* src, size_t n, size_t stride, float* out) {
    if (n_ == 1 || n < 4096) { host_pack_slice(src, n, stride, out); return; }
    {
      std::lock_guard<std::mutex> lk(mu_);
      src_ = src; cnt_ = n; stride_ = stride; out_ = out;
      pending_ = n_ - 1;
      gen_++;
    }
    cv_.notify_all();
    slice(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void slice(int w) {
    const size_t per = (cnt_ + size_t(n_) - 1) / size_t(n_);
    const size_t a = per * size_t(w) < cnt_ ? per * size_t(w) : cnt_;
    const size_t b = a + per < cnt_ ? a + per : cnt_;
    if (b > a) host_pack_slice(src_ + a * stride_, b - a, stride_, out_ + a * 4);
  }
  void loop(int w) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      slice(w);
      {
        std::lock_guard<std::mutex> lk(mu_);
        pending_--;
      }
      done_.notify_one();
    }
  }
  int n_ = 1;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  uint64_t gen_ = 0;
  bool stop_ = false;
  int pending_ = 0;
  const unsigned char* src_ = nullptr;
  size_t cnt_ = 0, stride_ = 0;
  float* out_ = nullptr;
};

}  // namespace pcr
