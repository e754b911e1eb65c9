// Shared device/host utilities for the PCR CUDA library (sm_100a).
#pragma once
#include <exception>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <stdexcept>
#include <utility>
#include <vector>

namespace pcr {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define PCR_CUDA_CHECK(expr)                                                                                   \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess)                                                                                     \
      throw ::pcr::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                             std::to_string(__LINE__) + ")");                                                  \
  } while (0)

// Process-wide cache of device allocations, bucketed by size class and device. cudaMalloc / cudaFree cost milliseconds
// (occasionally tens of milliseconds, and cudaFree synchronises the device): a register that rebuilds its index on every
// scan2Map call, or a frontend whose submap grows frame by frame, would pay that inside the hot loop. Freed buffers are
// parked here instead and handed out again to the next request of the same size class — also across contexts.
class DevPool {
 public:
  static size_t size_class(size_t bytes) {
    if (bytes <= 256) return 256;
    if (bytes > (size_t(1) << 30)) return (bytes + (size_t(1) << 28) - 1) & ~((size_t(1) << 28) - 1);  // > 1 GiB: 256 MiB steps
    size_t c = 256;
    while (c < bytes) c <<= 1;
    return c;
  }
  // Small size classes are carved out of slabs: a miss for a class of at most 4 MiB allocates 16 MiB worth of buffers of that
  // class with ONE cudaMalloc (2..16 of them) and parks the others. A frontend that caches one device copy per keyframe
  // (pcr_submap_build) otherwise paid a 1-3 ms cudaMalloc on every other keyframe (PCR_TRACE=1, profiles/README.md).
  static constexpr size_t kSlabMaxClass = size_t(4) << 20;
  static constexpr size_t kSlabBytes = size_t(16) << 20;

  static void* alloc(size_t bytes, size_t& got, int& dev) {
    got = size_class(bytes);
    dev = 0;
    PCR_CUDA_CHECK(cudaGetDevice(&dev));
    {
      std::lock_guard<std::mutex> lk(mu());
      auto& lst = free_list()[std::make_pair(dev, got)];
      if (!lst.empty()) {
        void* p = lst.back();
        lst.pop_back();
        auto it = child_slab().find(p);
        if (it != child_slab().end()) slabs()[it->second].n_free--;
        return p;
      }
    }
    const size_t k = got <= kSlabMaxClass ? std::min<size_t>(16, std::max<size_t>(2, kSlabBytes / got)) : 1;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, got * k);
    if (e == cudaErrorMemoryAllocation) {  // out of memory: give the parked buffers back to the driver and retry once
      cudaGetLastError();
      trim();
      e = cudaMalloc(&p, got * k);
    }
    if (e != cudaSuccess) throw CudaError(std::string("cudaMalloc(") + std::to_string(got * k) + " bytes) failed: " + cudaGetErrorString(e));
    if (k > 1) {
      std::lock_guard<std::mutex> lk(mu());
      const int id = int(slabs().size());
      slabs().push_back(Slab{p, int(k), int(k) - 1, true});
      auto& lst = free_list()[std::make_pair(dev, got)];
      for (size_t c = 0; c < k; c++) {
        void* child = static_cast<unsigned char*>(p) + c * got;
        child_slab()[child] = id;
        if (c) lst.push_back(child);
      }
    }
    return p;
  }
  // the caller guarantees that no work touching p is still in flight; dev = the device the allocation came from
  static void release(void* p, size_t got, int dev) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu());
    auto it = child_slab().find(p);
    if (it != child_slab().end()) slabs()[it->second].n_free++;
    free_list()[std::make_pair(dev, got)].push_back(p);
  }
  // bytes parked in the cache (all devices)
  static size_t cached_bytes() {
    std::lock_guard<std::mutex> lk(mu());
    size_t t = 0;
    for (auto& kv : free_list()) t += kv.first.second * kv.second.size();
    return t;
  }
  // hands parked buffers back to the driver; a slab goes back once all of its buffers are parked
  static void trim() {
    std::lock_guard<std::mutex> lk(mu());
    int cur = 0;
    const bool have = cudaGetDevice(&cur) == cudaSuccess;
    for (auto& kv : free_list()) {
      if (kv.second.empty()) continue;
      cudaSetDevice(kv.first.first);
      std::vector<void*> keep;
      for (void* p : kv.second) {
        auto it = child_slab().find(p);
        if (it == child_slab().end()) { cudaFree(p); continue; }
        Slab& sl = slabs()[it->second];
        if (sl.n_free < sl.n_children) { keep.push_back(p); continue; }  // a sibling is still in use
        if (sl.live) { cudaFree(sl.base); sl.live = false; }
        child_slab().erase(it);
      }
      kv.second.swap(keep);
    }
    if (have) cudaSetDevice(cur);
  }

 private:
  struct Slab { void* base; int n_children; int n_free; bool live; };
  static std::vector<Slab>& slabs() { static auto* v = new std::vector<Slab>(); return *v; }
  static std::map<void*, int>& child_slab() { static auto* m = new std::map<void*, int>(); return *m; }
  static std::mutex& mu() { static std::mutex m; return m; }
  static std::map<std::pair<int, size_t>, std::vector<void*>>& free_list() {
    static auto* m = new std::map<std::pair<int, size_t>, std::vector<void*>>();  // leaked on purpose: outlives the CUDA context teardown
    return *m;
  }
};

// Grow-only device buffer (allocations come from DevPool).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;        // elements
  size_t bytes_ = 0;     // size class the allocation was taken from
  int dev_ = 0;          // device the allocation lives on
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  T* ensure(size_t n) {
    if (n > cap) {
      if (p) {
        cudaDeviceSynchronize();  // the old buffer may still be read by queued work (what cudaFree's implicit sync used to cover)
        DevPool::release(p, bytes_, dev_);
        p = nullptr;
        cap = 0;
      }
      const size_t want = n + n / 4 + 64;
      p = static_cast<T*>(DevPool::alloc(want * sizeof(T), bytes_, dev_));
      cap = bytes_ / sizeof(T);
    }
    return p;
  }
  void release() {
    if (p) {
      if (std::uncaught_exceptions() > 0) {  // unwinding: queued work may still touch the buffer; the pool hands it to anyone
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != dev_) cudaSetDevice(dev_);
        cudaDeviceSynchronize();
        if (cur >= 0 && cur != dev_) cudaSetDevice(cur);
      }
      DevPool::release(p, bytes_, dev_);
    }
    p = nullptr;
    cap = 0;
  }
};

// Pinned host staging buffer.
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  T* ensure(size_t n) {
    if (n > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      size_t want = n + n / 4 + 64;
      PCR_CUDA_CHECK(cudaMallocHost(&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
};

constexpr int kNumSMs = 148;  // B200

// ---- integer voxel grid (PCL VoxelGrid key math; see voxel.cu) ------------------------------------------------
struct GridSpec {
  float inv_leaf[3];
  float leaf[3];
  int min_b[3];
  int max_b[3];
  int div_b[3];
  int mul[3];
  long long ncell;
};

// ---- device helpers ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned enc_f32(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_f32(unsigned u) {
  unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// PCL key: ijk = (int)(floorf(x * inv_leaf) - (float)min_b)   (pcp.hpp:206-210) — single-rounded float ops only.
__device__ __forceinline__ int voxel_axis(float v, float inv_leaf, int min_b) {
  return static_cast<int>(__fsub_rn(floorf(__fmul_rn(v, inv_leaf)), static_cast<float>(min_b)));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of NV doubles held per thread: warp shuffles, then fixed-order sum over warps.
// Result valid in threads [0, NV) of the block (thread t holds component t). smem: NV * (BLOCK/32) doubles.
template <int NV, int BLOCK>
__device__ __forceinline__ double block_reduce_vec(double (&v)[NV], double* smem) {
  constexpr int NW = BLOCK / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double s = warp_sum(v[k]);
    if (lane == 0) smem[warp * NV + k] = s;
  }
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < NV) {
#pragma unroll
    for (int w = 0; w < NW; w++) r += smem[w * NV + threadIdx.x];
  }
  return r;
}
#endif

}  // namespace pcr
