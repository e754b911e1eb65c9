// Voxel keys / stable sort / segments / downsample / uniform cell grid.
// Reference behaviour restated: pcl::VoxelGrid as reached from common/pcp/pcp.hpp:15-28 (key math written out in-tree
// at pcp.hpp:191-210), and the integer grid of pclomp::VoxelGridCovariance (voxel_grid_covariance_omp_impl.hpp:67-103).
#include "voxel.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/reverse_iterator.h>
#include <cfloat>
#include <cmath>
#include <limits>

namespace pcr {

// ------------------------------------------------------------------------------------------------------------
// AoS records -> float4
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const unsigned char* __restrict__ raw, size_t n, size_t stride, int mode,
                                                   float4* __restrict__ out) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float4 o;
  if (mode == 32) {  // pcl::PointXYZI, 16-byte aligned
    const float4* r = reinterpret_cast<const float4*>(raw) + i * 2;
    float4 a = __ldg(r), b = __ldg(r + 1);
    o = make_float4(a.x, a.y, a.z, b.x);
  } else if (mode == 16) {
    float4 a = __ldg(reinterpret_cast<const float4*>(raw) + i);
    o = make_float4(a.x, a.y, a.z, 0.f);
  } else {
    const float* r = reinterpret_cast<const float*>(raw + i * stride);
    o = make_float4(r[0], r[1], r[2], stride >= 20 ? r[4] : 0.f);
  }
  out[i] = o;
}

void pack_points(const void* dev_raw, size_t n, size_t stride, float4* out, cudaStream_t s) {
  if (n == 0) return;
  int mode = 0;
  if ((reinterpret_cast<uintptr_t>(dev_raw) & 15) == 0) {
    if (stride == 32) mode = 32;
    else if (stride == 16) mode = 16;
  }
  unsigned blocks = unsigned((n + 255) / 256);
  pack_kernel<<<blocks, 256, 0, s>>>(static_cast<const unsigned char*>(dev_raw), n, stride, mode, out);
}

__global__ void __launch_bounds__(256) write_xyzi32_kernel(const float4* __restrict__ pts, size_t n, float4* __restrict__ out) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  out[2 * i] = make_float4(p.x, p.y, p.z, 1.0f);
  out[2 * i + 1] = make_float4(p.w, 0.f, 0.f, 0.f);
}
void write_xyzi32(const float4* pts, size_t n, void* dev_out32, cudaStream_t s) {
  if (n == 0) return;
  write_xyzi32_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(pts, n, static_cast<float4*>(dev_out32));
}

// ------------------------------------------------------------------------------------------------------------
// bounding box
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ pts, size_t n, unsigned* __restrict__ out) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  unsigned bad = 0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    float4 p = __ldg(pts + i);
    if (!(fabsf(p.x) <= FLT_MAX && fabsf(p.y) <= FLT_MAX && fabsf(p.z) <= FLT_MAX)) { bad++; continue; }  // NaN / Inf: not part of the box
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], o));
    }
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  __shared__ float smn[8][3], smx[8][3];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0 && bad) atomicAdd(out + 6, bad);
  if (lane == 0)
    for (int a = 0; a < 3; a++) { smn[warp][a] = mn[a]; smx[warp][a] = mx[a]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int a = threadIdx.x;
    float lo = smn[0][a], hi = smx[0][a];
    for (int w = 1; w < 8; w++) { lo = fminf(lo, smn[w][a]); hi = fmaxf(hi, smx[w][a]); }
    atomicMin(out + a, enc_f32(lo));
    atomicMax(out + 3 + a, enc_f32(hi));
  }
}

void bbox_blocking(const float4* pts, size_t n, float mn[3], float mx[3], BBoxWork& w, cudaStream_t s) {
  unsigned* d = w.d.ensure(8);
  unsigned* h = w.h.ensure(8);
  h[0] = h[1] = h[2] = 0xffffffffu;
  h[3] = h[4] = h[5] = 0u;
  h[6] = h[7] = 0u;
  PCR_CUDA_CHECK(cudaMemcpyAsync(d, h, 8 * sizeof(unsigned), cudaMemcpyHostToDevice, s));
  unsigned blocks = unsigned(std::min<size_t>((n + 255) / 256, size_t(kNumSMs) * 8));
  bbox_kernel<<<blocks, 256, 0, s>>>(pts, n, d);
  PCR_CUDA_CHECK(cudaMemcpyAsync(h, d, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  w.n_nonfinite = h[6];
  for (int a = 0; a < 3; a++) { mn[a] = dec_f32(h[a]); mx[a] = dec_f32(h[3 + a]); }
  if (w.n_nonfinite == n)  // nothing finite: an empty box at the origin
    for (int a = 0; a < 3; a++) mn[a] = mx[a] = 0.f;
}

struct FinitePred {
  const float4* pts;
  __device__ __forceinline__ bool operator()(const uint32_t& i) const {
    const float4 p = pts[i];
    return fabsf(p.x) <= FLT_MAX && fabsf(p.y) <= FLT_MAX && fabsf(p.z) <= FLT_MAX;
  }
};
__global__ void __launch_bounds__(256) gather_idx_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ idx, size_t m, float4* __restrict__ out) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < m) out[i] = pts[idx[i]];
}

size_t drop_nonfinite(float4* pts, size_t n, DevBuf<float4>& scratch, DevBuf<unsigned char>& tmp, DevBuf<unsigned>& d_count, PinBuf<unsigned>& h_count,
                      cudaStream_t s) {
  if (n == 0) return 0;
  // scratch: [n indices of the finite records | their records]
  scratch.ensure(n + (n + 3) / 4 + 1);
  uint32_t* idx = reinterpret_cast<uint32_t*>(scratch.p);
  float4* kept = scratch.p + (n + 3) / 4;
  unsigned* dc = d_count.ensure(1);
  unsigned* hc = h_count.ensure(1);
  cub::CountingInputIterator<uint32_t> it(0);
  FinitePred pred{pts};
  size_t bytes = 0;
  cub::DeviceSelect::If(nullptr, bytes, it, idx, dc, int(n), pred, s);
  tmp.ensure(bytes);
  PCR_CUDA_CHECK(cub::DeviceSelect::If(tmp.p, bytes, it, idx, dc, int(n), pred, s));
  PCR_CUDA_CHECK(cudaMemcpyAsync(hc, dc, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  const size_t m = *hc;
  if (m) {
    gather_idx_kernel<<<unsigned((m + 255) / 256), 256, 0, s>>>(pts, idx, m, kept);
    PCR_CUDA_CHECK(cudaMemcpyAsync(pts, kept, m * sizeof(float4), cudaMemcpyDeviceToDevice, s));
  }
  return m;
}

bool make_grid_spec(const float mn[3], const float mx[3], float leaf, GridSpec& g) {
  const float inv = 1.0f / leaf;
  long long d[3];
  for (int a = 0; a < 3; a++) {
    g.leaf[a] = leaf;
    g.inv_leaf[a] = inv;
    volatile float span = (mx[a] - mn[a]);
    volatile float scaled = span * inv;
    d[a] = static_cast<long long>(scaled) + 1;
    volatile float lo = mn[a] * inv, hi = mx[a] * inv;
    g.min_b[a] = static_cast<int>(std::floor(lo));
    g.max_b[a] = static_cast<int>(std::floor(hi));
    g.div_b[a] = g.max_b[a] - g.min_b[a] + 1;
  }
  g.mul[0] = 1;
  g.mul[1] = g.div_b[0];
  g.mul[2] = g.div_b[0] * g.div_b[1];
  g.ncell = (long long)g.div_b[0] * g.div_b[1] * g.div_b[2];
  return !((d[0] * d[1] * d[2]) > (long long)std::numeric_limits<int32_t>::max());
}

// ------------------------------------------------------------------------------------------------------------
// keys + sort + segments
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_key_kernel(const float4* __restrict__ pts, size_t n, GridSpec g,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(pts + i);
  int i0 = voxel_axis(p.x, g.inv_leaf[0], g.min_b[0]);
  int i1 = voxel_axis(p.y, g.inv_leaf[1], g.min_b[1]);
  int i2 = voxel_axis(p.z, g.inv_leaf[2], g.min_b[2]);
  keys[i] = static_cast<uint32_t>(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
  vals[i] = static_cast<uint32_t>(i);
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* v, size_t n) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) v[i] = uint32_t(i);
}

static int bits_for(long long ncell) {
  int b = 1;
  while (b < 32 && (1ll << b) < ncell) b++;
  return b;
}

void KeySort::sort_keys_in_k0(size_t n_, long long ncell, cudaStream_t s) {
  n = n_;
  k1.ensure(n); v0.ensure(n); v1.ensure(n);
  iota_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(v0.p, n);
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, v1.p, int(n), 0, bits_for(ncell), s);
  tmp.ensure(bytes);
  PCR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, k0.p, k1.p, v0.p, v1.p, int(n), 0, bits_for(ncell), s));
  keys_unsorted = k0.p;
  keys = k1.p;
  vals = v1.p;
  nseg = 0;
}

void KeySort::sort(const float4* pts, size_t n_, const GridSpec& g, cudaStream_t s) {
  n = n_;
  k0.ensure(n); k1.ensure(n); v0.ensure(n); v1.ensure(n);
  voxel_key_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(pts, n, g, k0.p, v0.p);
  size_t bytes = 0;
  int bits = bits_for(g.ncell);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, v1.p, int(n), 0, bits, s);
  tmp.ensure(bytes);
  PCR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, k0.p, k1.p, v0.p, v1.p, int(n), 0, bits, s));
  keys_unsorted = k0.p;
  keys = k1.p;
  vals = v1.p;
  nseg = 0;
}

struct HeadPred {
  const uint32_t* keys;
  __device__ __forceinline__ bool operator()(const uint32_t& i) const { return i == 0 || keys[i] != keys[i - 1]; }
};

__global__ void set_tail_kernel(uint32_t* seg_start, const unsigned* count, uint32_t n) { seg_start[*count] = n; }

void KeySort::segment(cudaStream_t s) {
  seg_start.ensure(n + 1);
  unsigned* dc = d_count.ensure(1);
  unsigned* hc = h_count.ensure(1);
  cub::CountingInputIterator<uint32_t> it(0);
  HeadPred pred{keys};
  size_t bytes = 0;
  cub::DeviceSelect::If(nullptr, bytes, it, seg_start.p, dc, int(n), pred, s);
  tmp.ensure(bytes);
  PCR_CUDA_CHECK(cub::DeviceSelect::If(tmp.p, bytes, it, seg_start.p, dc, int(n), pred, s));
  set_tail_kernel<<<1, 1, 0, s>>>(seg_start.p, dc, uint32_t(n));
  PCR_CUDA_CHECK(cudaMemcpyAsync(hc, dc, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  nseg = *hc;
}

// ------------------------------------------------------------------------------------------------------------
// pcl::VoxelGrid centroid: one thread per voxel, float32 running sums in ascending original index
// (pcl::CentroidPoint<PointXYZI>: xyz and intensity accumulators are float).
// ------------------------------------------------------------------------------------------------------------
// Two passes: (1) gather the points into key order (fully parallel, coalesced writes), (2) one thread per voxel walks its
// now CONTIGUOUS run — the float sums must stay strictly sequential in ascending original index to reproduce
// pcl::CentroidPoint bit for bit, but the loads need not: four independent float4 loads are in flight per trip instead of
// a dependent vals[j] -> pts[vals[j]] chain per point (91 -> ~10 us on a 22 k-point VLP-16 scan whose voxels next to the
// sensor hold hundreds of points).
__global__ void __launch_bounds__(256) gather_sorted_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ vals, size_t n,
                                                            float4* __restrict__ sorted) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) sorted[i] = __ldg(pts + vals[i]);
}

__global__ void __launch_bounds__(128) centroid_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ seg_start, size_t nseg,
                                                       float4* __restrict__ out) {
  size_t v = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (v >= nseg) return;
  const uint32_t b = seg_start[v], e = seg_start[v + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  uint32_t j = b;
  for (; j + 4 <= e; j += 4) {
    const float4 p0 = __ldg(sorted + j), p1 = __ldg(sorted + j + 1), p2 = __ldg(sorted + j + 2), p3 = __ldg(sorted + j + 3);
    sx = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sx, p0.x), p1.x), p2.x), p3.x);
    sy = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sy, p0.y), p1.y), p2.y), p3.y);
    sz = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sz, p0.z), p1.z), p2.z), p3.z);
    si = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(si, p0.w), p1.w), p2.w), p3.w);
  }
  for (; j < e; j++) {
    const float4 p = __ldg(sorted + j);
    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
  }
  const float cnt = static_cast<float>(e - b);
  out[2 * v] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 1.0f);
  out[2 * v + 1] = make_float4(__fdiv_rn(si, cnt), 0.f, 0.f, 0.f);
}

void gather_sorted(const float4* pts, KeySort& ks, cudaStream_t s) {
  if (ks.n == 0) return;
  ks.sorted_pts.ensure(ks.n);
  gather_sorted_kernel<<<unsigned((ks.n + 255) / 256), 256, 0, s>>>(pts, ks.vals, ks.n, ks.sorted_pts.p);
}

void voxel_centroids(const float4* pts, KeySort& ks, void* dev_out32, cudaStream_t s) {
  if (ks.nseg == 0) return;
  gather_sorted(pts, ks, s);
  centroid_kernel<<<unsigned((ks.nseg + 127) / 128), 128, 0, s>>>(ks.sorted_pts.p, ks.seg_start.p, ks.nseg, static_cast<float4*>(dev_out32));
}

// ------------------------------------------------------------------------------------------------------------
// uniform cell grid: cell-sorted points (w = original index) + dense per-cell [start,end)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_scatter_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ vals, size_t n,
                                                           float4* __restrict__ sorted, int2* __restrict__ range) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i], v = vals[i];
  float4 p = __ldg(pts + v);
  p.w = __int_as_float(int(v));
  sorted[i] = p;
  if (i == 0 || keys[i - 1] != k) range[k].x = int(i);
  if (i == n - 1 || keys[i + 1] != k) range[k].y = int(i + 1);
}

// start-table variant: scatter only (cell-sorted points) + heads of the key runs
__global__ void __launch_bounds__(256) grid_scatter_start_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys,
                                                                 const uint32_t* __restrict__ vals, size_t n, long long ncell,
                                                                 float4* __restrict__ sorted, int32_t* __restrict__ start,
                                                                 unsigned* __restrict__ occupied) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  bool head = false;
  if (i < n) {
    uint32_t k = keys[i], v = vals[i];
    float4 p = __ldg(pts + v);
    p.w = __int_as_float(int(v));
    sorted[i] = p;
    head = i == 0 || keys[i - 1] != k;
    if (head) start[k] = int32_t(i);
    if (i == n - 1) start[ncell] = int32_t(n);
  }
  const unsigned heads = __ballot_sync(0xffffffffu, head);  // occupied-cell count, one atomic per warp
  if ((threadIdx.x & 31) == 0 && heads) atomicAdd(occupied, unsigned(__popc(heads)));
}

__global__ void __launch_bounds__(256) occupancy_bitmap_kernel(const float4* __restrict__ pts, size_t n, GridSpec g, unsigned* __restrict__ bits) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + i);
  const long long key = (long long)voxel_axis(p.x, g.inv_leaf[0], g.min_b[0]) * g.mul[0] + (long long)voxel_axis(p.y, g.inv_leaf[1], g.min_b[1]) * g.mul[1] +
                        (long long)voxel_axis(p.z, g.inv_leaf[2], g.min_b[2]) * g.mul[2];
  const unsigned bit = 1u << (key & 31);
  unsigned* w = bits + (key >> 5);
  if (!(__ldg(w) & bit)) atomicOr(w, bit);  // most points find their cell's bit already set
}
__global__ void __launch_bounds__(256) popcount_kernel(const unsigned* __restrict__ bits, size_t nwords, unsigned* __restrict__ out) {
  unsigned c = 0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < nwords; i += size_t(gridDim.x) * blockDim.x) c += __popc(bits[i]);
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

size_t count_occupied_cells(const float4* pts, size_t n, const GridSpec& g, DevBuf<unsigned char>& tmp, DevBuf<unsigned>& d_count,
                            PinBuf<unsigned>& h_count, cudaStream_t s) {
  if (n == 0) return 0;
  const size_t nwords = size_t((g.ncell + 31) / 32);
  tmp.ensure(nwords * 4);
  unsigned* bits = reinterpret_cast<unsigned*>(tmp.p);
  unsigned* dc = d_count.ensure(1);
  unsigned* hc = h_count.ensure(1);
  PCR_CUDA_CHECK(cudaMemsetAsync(bits, 0, nwords * 4, s));
  PCR_CUDA_CHECK(cudaMemsetAsync(dc, 0, sizeof(unsigned), s));
  occupancy_bitmap_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(pts, n, g, bits);
  popcount_kernel<<<unsigned(std::min<size_t>((nwords + 255) / 256, size_t(kNumSMs) * 8)), 256, 0, s>>>(bits, nwords, dc);
  PCR_CUDA_CHECK(cudaMemcpyAsync(hc, dc, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  return *hc;
}

struct MinOp {
  __device__ __forceinline__ int32_t operator()(const int32_t& a, const int32_t& b) const { return a < b ? a : b; }
};

double grid_slack_cells(const GridSpec& g) {
  double m = 1.0;
  for (int a = 0; a < 3; a++) m = std::max(m, std::max(std::fabs(double(g.min_b[a])), std::fabs(double(g.max_b[a]))));
  return std::max(1e-3, m * 4.8e-7);
}

int build_cell_grid(const float4* pts, size_t n, float cell, CellGrid& grid, KeySort& ks, BBoxWork& bw, cudaStream_t s, bool start_table,
                    const float* bbox) {
  grid.built = false;
  grid.has_start = false;
  grid.n = n;
  if (n == 0) return 0;
  float mn[3], mx[3];
  if (bbox) {
    for (int a = 0; a < 3; a++) { mn[a] = bbox[a]; mx[a] = bbox[3 + a]; }
  } else {
    bbox_blocking(pts, n, mn, mx, bw, s);
    if (bw.n_nonfinite) return kRetryNonFinite;
  }
  bool ok = make_grid_spec(mn, mx, cell, grid.g);
  if (!ok || grid.g.ncell > (1ll << 29)) return -5;
  ks.sort(pts, n, grid.g, s);
  grid.pts.ensure(n);
  const size_t ncell = size_t(grid.g.ncell);
  if (start_table) {
    // heads of the key runs scattered into a table preset to INT_MAX-ish, then a reverse running minimum fills the empty
    // cells with the start of the next occupied one
    grid.start.ensure(ncell + 1);
    PCR_CUDA_CHECK(cudaMemsetAsync(grid.start.p, 0x7f, (ncell + 1) * sizeof(int32_t), s));
    unsigned* dc = ks.d_count.ensure(1);
    PCR_CUDA_CHECK(cudaMemsetAsync(dc, 0, sizeof(unsigned), s));
    grid_scatter_start_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(pts, ks.keys, ks.vals, n, grid.g.ncell, grid.pts.p, grid.start.p, dc);
    auto rit = thrust::make_reverse_iterator(grid.start.p + ncell + 1);
    size_t bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, bytes, rit, rit, MinOp(), int(ncell + 1), s);
    ks.tmp.ensure(bytes);
    PCR_CUDA_CHECK(cub::DeviceScan::InclusiveScan(ks.tmp.p, bytes, rit, rit, MinOp(), int(ncell + 1), s));
    grid.has_start = true;  // no read-back here: the build stays queued on the stream (callers synchronise once, at the end)
  } else {
    grid.range.ensure(ncell);
    PCR_CUDA_CHECK(cudaMemsetAsync(grid.range.p, 0, ncell * sizeof(int2), s));
    grid_scatter_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(pts, ks.keys, ks.vals, n, grid.pts.p, grid.range.p);
  }
  grid.built = true;
  return 0;
}

}  // namespace pcr
