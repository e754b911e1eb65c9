// Build of the nested multi-resolution kNN grid (knn.cuh): Morton codes -> 64-bit radix sort -> per-level hash tables.
#include "knn.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cmath>

namespace pcr {

__global__ void __launch_bounds__(256) morton_code_kernel(const float4* __restrict__ pts, size_t n, float mn0, float mn1, float mn2, float inv_h0,
                                                          int d0, int d1, int d2, unsigned long long* __restrict__ codes,
                                                          uint32_t* __restrict__ vals) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + i);
  const int x = min(max(int(floorf(knn_scaled(p.x, mn0, inv_h0))), 0), d0 - 1);
  const int y = min(max(int(floorf(knn_scaled(p.y, mn1, inv_h0))), 0), d1 - 1);
  const int z = min(max(int(floorf(knn_scaled(p.z, mn2, inv_h0))), 0), d2 - 1);
  codes[i] = morton3(unsigned(x), unsigned(y), unsigned(z));
  vals[i] = uint32_t(i);
}

// slot of `key` in table `tab`, claiming an empty one if needed (open addressing, linear probing)
__device__ __forceinline__ uint4* knn_find_or_insert(uint4* tab, uint32_t mask, unsigned long long key) {
  uint32_t h = knn_hash(key) & mask;
  for (;;) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(tab + h);
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
    if (cur == 0ull) cur = atomicCAS(kp, 0ull, key);
    if (cur == 0ull || cur == key) return tab + h;
    h = (h + 1) & mask;
  }
}

// one thread per sorted position: gather the point, and at every level where it starts / ends a cell write start / end
__global__ void __launch_bounds__(256) morton_table_kernel(const float4* __restrict__ pts, const unsigned long long* __restrict__ codes,
                                                           const uint32_t* __restrict__ vals, size_t n, float4* __restrict__ sorted,
                                                           uint4* __restrict__ tables, uint32_t cap) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = vals[i];
  float4 p = __ldg(pts + v);
  p.w = __int_as_float(int(v));
  sorted[i] = p;
  const unsigned long long c = codes[i];
  // number of levels (from the finest) at which this point starts / ends a cell: levels whose prefix differs from the neighbour's
  int nh = kKnnLevels, nt = kKnnLevels;
  if (i > 0) {
    const unsigned long long d = c ^ codes[i - 1];
    nh = d ? min(kKnnLevels, (63 - __clzll((long long)d)) / 3 + 1) : 0;
  }
  if (i + 1 < n) {
    const unsigned long long d = c ^ codes[i + 1];
    nt = d ? min(kKnnLevels, (63 - __clzll((long long)d)) / 3 + 1) : 0;
  }
  const int nl = max(nh, nt);
  for (int l = 0; l < nl; l++) {
    uint4* slot = knn_find_or_insert(tables + size_t(l) * cap, cap - 1, (c >> (3 * l)) + 1ull);
    if (l < nh) slot->z = uint32_t(i);
    if (l < nt) slot->w = uint32_t(i + 1);
  }
}

int build_morton_grid(const float4* pts, size_t n, MortonGrid& grid, BBoxWork& bw, cudaStream_t s) {
  grid.built = false;
  grid.n = n;
  if (n == 0) return 0;
  float mn[3], mx[3];
  bbox_blocking(pts, n, mn, mx, bw, s);
  if (bw.n_nonfinite) return kRetryNonFinite;
  double m = 1.0;
  for (int a = 0; a < 3; a++) {
    grid.mn[a] = mn[a];
    volatile float span = mx[a] - mn[a];
    volatile float sc = span * grid.inv_h0;
    const double cells = std::floor(double(sc)) + 1.0;
    if (!(cells < double(1 << 21))) return -5;  // 21 bits per axis at 1/32 m: 65 km
    grid.dim0[a] = int(cells);
    grid.smax[a] = std::nextafter(float(grid.dim0[a]), 0.0f);
    m = std::max(m, cells);
  }
  // subtraction + multiplication rounding of both the query and the map point: ~4 ulp of the scaled coordinate
  grid.slack0 = std::max(1e-3, m * 4.8e-7);
  uint32_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  grid.cap = cap;
  grid.pts.ensure(n);
  grid.c0.ensure(n); grid.c1.ensure(n); grid.v0.ensure(n); grid.v1.ensure(n);
  grid.tables.ensure(size_t(kKnnLevels) * cap);
  PCR_CUDA_CHECK(cudaMemsetAsync(grid.tables.p, 0, size_t(kKnnLevels) * cap * sizeof(uint4), s));
  const unsigned blocks = unsigned((n + 255) / 256);
  morton_code_kernel<<<blocks, 256, 0, s>>>(pts, n, mn[0], mn[1], mn[2], grid.inv_h0, grid.dim0[0], grid.dim0[1], grid.dim0[2], grid.c0.p, grid.v0.p);
  // only the bits the extent uses take part in the sort: a 200 m cloud at 1/32 m needs 13 bits per axis = 5 radix passes, not 8
  int axis_bits = 1;
  while ((1ll << axis_bits) < (long long)m) axis_bits++;
  const int end_bit = std::min(63, 3 * axis_bits);
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, grid.c0.p, grid.c1.p, grid.v0.p, grid.v1.p, int(n), 0, end_bit, s);
  grid.tmp.ensure(bytes);
  PCR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(grid.tmp.p, bytes, grid.c0.p, grid.c1.p, grid.v0.p, grid.v1.p, int(n), 0, end_bit, s));
  morton_table_kernel<<<blocks, 256, 0, s>>>(pts, grid.c1.p, grid.v1.p, n, grid.pts.p, grid.tables.p, cap);
  PCR_CUDA_CHECK(cudaGetLastError());
  grid.built = true;
  return 0;
}

}  // namespace pcr
