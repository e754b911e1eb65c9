// Shared device/host utilities for the PCR CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <stdexcept>
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

// Grow-only device buffer.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  T* ensure(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      p = nullptr;
      size_t want = n + n / 4 + 64;
      PCR_CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
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
