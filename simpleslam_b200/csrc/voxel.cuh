// Voxel keys, stable key sort, segments, PCL-VoxelGrid downsample and the uniform cell grid used by the kNN kernels.
#pragma once
#include "common.cuh"

namespace pcr {

// float4 point: (x, y, z, intensity). In cell-sorted arrays w carries the ORIGINAL index (int bits).
void pack_points(const void* dev_raw, size_t n, size_t stride, float4* out, cudaStream_t s);
void write_xyzi32(const float4* pts, size_t n, void* dev_out32, cudaStream_t s);

// Bounding box over the FINITE points (blocking: returns host values). n must be > 0. w.n_nonfinite = records with a NaN / Inf
// coordinate (the reference strips them before a register sees a cloud, dataproxy/src/LidarDataProxy.cpp:47, and
// pcl::VoxelGrid / VoxelGridCovariance skip them): index builds answer kRetryNonFinite so that the API layer compacts
// the context-owned copy of the cloud (drop_nonfinite) and builds again.
struct BBoxWork { DevBuf<unsigned> d; PinBuf<unsigned> h; size_t n_nonfinite = 0; };
void bbox_blocking(const float4* pts, size_t n, float mn[3], float mx[3], BBoxWork& w, cudaStream_t s);
constexpr int kRetryNonFinite = -100;  // internal: never crosses the C ABI
// order-preserving removal of records with a non-finite coordinate, in place (blocking). Returns the new count.
size_t drop_nonfinite(float4* pts, size_t n, DevBuf<float4>& scratch, DevBuf<unsigned char>& tmp, DevBuf<unsigned>& d_count, PinBuf<unsigned>& h_count,
                      cudaStream_t s);

// PCL VoxelGrid grid parameters from a bounding box (pcp.hpp:191-200; voxel_grid_covariance_omp_impl.hpp:75-103).
// Returns false when dx*dy*dz overflows int32 (PCL's "leaf size too small").
bool make_grid_spec(const float mn[3], const float mx[3], float leaf, GridSpec& g);

// (key, original index) pairs sorted by key, stable => ascending original index inside a voxel.
struct KeySort {
  DevBuf<uint32_t> k0, k1, v0, v1, seg_start;
  DevBuf<float4> sorted_pts;          // points gathered into key order (centroid pass)
  DevBuf<unsigned char> tmp;
  DevBuf<unsigned> d_count;
  PinBuf<unsigned> h_count;
  uint32_t* keys_unsorted = nullptr;  // per input point
  uint32_t* keys = nullptr;           // sorted
  uint32_t* vals = nullptr;           // original indices, sorted by key
  size_t n = 0;
  size_t nseg = 0;                    // valid after segment()
  // keys -> sort
  void sort(const float4* pts, size_t n, const GridSpec& g, cudaStream_t s);
  // generic: sort precomputed keys living in k0 (vals are generated)
  void sort_keys_in_k0(size_t n, long long ncell, cudaStream_t s);
  // run starts of equal keys -> seg_start[0..nseg] (seg_start[nseg] = n). Blocking (reads nseg back).
  void segment(cudaStream_t s);
};

// pcl::VoxelGrid centroid per segment: float32 sums in ascending original index, / count.
// out32: 32-byte PointXYZI records.
void voxel_centroids(const float4* pts, KeySort& ks, void* dev_out32, cudaStream_t s);
// ks.sorted_pts[i] = pts[ks.vals[i]]: the points in key order (ascending original index inside a voxel)
void gather_sorted(const float4* pts, KeySort& ks, cudaStream_t s);

// Uniform grid over a cloud for neighbour search.
struct CellGrid {
  GridSpec g{};
  DevBuf<float4> pts;      // cell-sorted, w = original index bits
  DevBuf<int2> range;      // per cell [start, end)
  DevBuf<int32_t> start;   // optional dense prefix table: start[k] = first sorted position whose key >= k, start[ncell] = n;
                           // a run of consecutive cells [k0, k1] is the contiguous slice [start[k0], start[k1 + 1])
  bool has_start = false;
  size_t occupied = 0;     // occupied cells (start-table builds only)
  int max_ring = 1;        // LOAM search: rings of cells (1 = 27 cells, 2 = 125 cells) needed to cover the gate radius
  size_t n = 0;
  bool built = false;
};
// Kernel-side view of a CellGrid.
struct CellGridView {
  const float4* pts;
  const int2* range;
  const int32_t* start;
  GridSpec g;
};
inline CellGridView view_of(const CellGrid& grid) {
  CellGridView v;
  v.pts = grid.pts.p;
  v.range = grid.range.p;
  v.start = grid.has_start ? grid.start.p : nullptr;
  v.g = grid.g;
  return v;
}
// Returns PCR error code (0 ok, -5 grid too large).
// start_table: build `start` (LOAM row-run lookups) instead of `range`.
// bbox (nullable): min[3], max[3] of the finite points if the caller has them already (saves the bounding-box pass + sync).
int build_cell_grid(const float4* pts, size_t n, float cell, CellGrid& grid, KeySort& ks, BBoxWork& bw, cudaStream_t s,
                    bool start_table = false, const float* bbox = nullptr);
// number of occupied cells of the grid `g` would have over `pts` (bitmap + popcount, no sort): lets a caller pick the cell
// size from the density BEFORE building the index. Blocking (one read-back).
size_t count_occupied_cells(const float4* pts, size_t n, const GridSpec& g, DevBuf<unsigned char>& tmp, DevBuf<unsigned>& d_count,
                            PinBuf<unsigned>& h_count, cudaStream_t s);
// float rounding of x * inv_leaf moves a point by at most ~|cell index| * 2^-23 cells across a cell face: slack (in cells)
// that exactness arguments about "every point outside the scanned cells is farther than ..." have to subtract
double grid_slack_cells(const GridSpec& g);

}  // namespace pcr
