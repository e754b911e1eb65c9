// pclomp NDT on the GPU: voxel-covariance target build (N1), neighbourhood lookup (N2/N6), derivative kernels (N3/N4)
// and the host-side Newton + More-Thuente driver (N5), batched over independent scans.
#pragma once
#include "common.cuh"
#include "voxel.cuh"
#include "../../include/pcr_cuda.h"
#include <memory>

namespace pcr {

struct __align__(16) NdtLeafRec {  // 64 B: one leaf = two 32-byte sectors
  double mean[3];
  float icov[9];
  int npts;
};

struct NdtTarget {
  GridSpec g{};
  bool built = false;
  bool overflow = false;  // PCL: grid would overflow int32 -> no leaves
  size_t nleaves = 0;     // all occupied voxels (segments)
  float resolution = 1.f;
  double d1 = 0, d2 = 0, d3 = 0;
  DevBuf<NdtLeafRec> recs;
  DevBuf<int32_t> keys, npts;
  DevBuf<double> mean, cov, icov;  // double copies (computeHessian path + introspection)
  DevBuf<int32_t> table;           // dense cell -> leaf index (-1: none / < min points / rejected)
  // KDTREE mode: centroid grid (float centroids of leaves with >= min points)
  DevBuf<float4> centroids;        // per leaf (w = leaf index bits), only meaningful where in_cloud
  CellGrid cgrid;
};

struct NdtEvalParams {  // per scan, per evaluation
  float Tf[16];         // cloud transform (column-major float)
  float j_ang[8][3];
  float h_ang[15][3];
  double j_ang_d[8][3];
  double h_ang_d[15][3];
  int compute_hessian;
  int kind;             // 0 = computeDerivatives (float path), 1 = computeHessian (double path)
  int scan;             // which scan of the batch
  int pad;
};

struct NdtEvalResult { double v[30]; };  // score, g[6], H upper-triangular 21 (row-major order r<=c), pairs, pad

struct NdtDriver {
  DevBuf<NdtEvalParams> d_params;
  DevBuf<NdtEvalResult> d_results;
  DevBuf<double> partials;
  DevBuf<unsigned> tickets;
  DevBuf<uint32_t> offsets;
  PinBuf<NdtEvalParams> h_params;
  PinBuf<NdtEvalResult> h_results;
  PinBuf<uint32_t> h_offsets;
  size_t partial_cap_blocks = 0;
  long long launches = 0;
  float hot_ms = 0.f;
  int hot_launches = 0;
  int total_evals = 0, total_hess = 0;
  long long total_pairs = 0;
  long long point_evals = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  NdtEvalResult* mapped_results = nullptr;  // host-mapped pinned memory the last block writes into
  size_t mapped_cap = 0;
  ~NdtDriver();

  // second lane of a batched align (own buffers, own stream): see NdtDriver::align
  std::unique_ptr<NdtDriver> second;
  cudaStream_t second_stream = nullptr;
  cudaEvent_t ready = nullptr;
  bool pending_run = false;
  // launch: queue one evaluation round (h_params[0..count)) on stream s; collect: wait for it -> h_results[0..count)
  void launch(const float4* src, const uint32_t* d_offs, size_t max_pts, const NdtTarget& tgt, int search, int count, bool profile, cudaStream_t s);
  void collect(int count, bool profile, cudaStream_t s);
  // evaluate `count` requests (h_params[0..count)) -> h_results[0..count). Blocking.
  void evaluate(const float4* src, const uint32_t* d_offs, size_t max_pts, const NdtTarget& tgt, int search, int count, bool profile,
                cudaStream_t s);
  // full registration of n_scans scans (offs: host offsets, n_scans+1)
  int align(const float4* src, const size_t* offs, size_t n_scans, const NdtTarget& tgt, const pcr_params& prm, double* T,
            int32_t* converged, int32_t* iters, double* trans_prob, bool profile, cudaStream_t s);
};

int ndt_build_target(const float4* pts, size_t n, const pcr_params& prm, NdtTarget& tgt, KeySort& ks, BBoxWork& bw, cudaStream_t s);

}  // namespace pcr
