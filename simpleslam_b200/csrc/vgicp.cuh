// FastVGICP on the GPU: exact k-NN covariances (V1), Gaussian voxel map (V2), correspondence + Mahalanobis +
// linearize / compute_error kernels (V3/V4), LM / GN driver (V5), fitness (V6).
#pragma once
#include "common.cuh"
#include "voxel.cuh"
#include "knn.cuh"
#include "vgicp_logic.cuh"
#include "../../include/pcr_cuda.h"
#include <utility>
#include <vector>

namespace pcr {

// `prof` (nullable): CUDA-event time of the k-NN kernel is added to it.
struct KnnProfile {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float ms = 0.f;
  int launches = 0;
  long long queries = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;  // recorded, not yet read back
  ~KnnProfile();
  void begin(cudaStream_t s);
  void end(cudaStream_t s, size_t n);
  void collect();  // after a stream synchronise
  void reset() { ms = 0.f; launches = 0; queries = 0; }
};

struct __align__(16) VoxelRec {  // 80 B
  double mean[3];
  double cov[6];  // xx xy xz yy yz zz
  double n;       // number of points (as double: w = sqrt(n))
};

struct VgicpTarget {
  bool built = false;
  size_t n = 0;
  MortonGrid grid;            // over the raw target (kNN for covariances and fitness)
  DevBuf<double> covs;        // per target point, 6 doubles
  DevBuf<int32_t> knn;        // k neighbour indices per target point (scratch of the covariance build)
  KnnProfile* prof = nullptr; // set by the API layer while profiling is on
  // voxel map
  double resolution = 1.0;
  int cmin[3] = {0, 0, 0}, cdim[3] = {0, 0, 0};
  long long ncell = 0;
  size_t nvox = 0;
  DevBuf<VoxelRec> vox;
  DevBuf<int32_t> vox_key;    // linear key per voxel (ascending => (z,y,x) order)
  DevBuf<int32_t> table;      // dense cell -> voxel id or -1
};

struct VgicpEvalResult { double v[30]; };  // cost, H upper 21, b 6, count, pad

struct VgicpProgress {  // host-mapped pinned memory
  volatile int round;  // evaluation launch that has started
  volatile int done;   // the registration has finished (set by the tail that ends it)
  int pad[14];
};

struct VgicpDriver {
  DevBuf<VgicpState> d_states;
  PinBuf<VgicpState> h_states;
  VgicpProgress* progress = nullptr;
  DevBuf<VgicpEvalParams> d_params;
  DevBuf<VgicpEvalResult> d_results;
  DevBuf<double> partials;
  DevBuf<unsigned> tickets;
  DevBuf<uint32_t> offsets;
  DevBuf<double> src_covs;
  DevBuf<int32_t> knn_dbg;
  DevBuf<double> fit_partials;
  PinBuf<VgicpEvalParams> h_params;
  PinBuf<VgicpEvalResult> h_results;
  PinBuf<uint32_t> h_offsets;
  PinBuf<double> h_fit;
  MortonGrid src_grid;
  long long launches = 0;
  float hot_ms = 0.f;
  int hot_launches = 0;
  int n_linearize = 0, n_error = 0;
  int64_t last_corr = 0;
  long long total_corr = 0;
  double last_cost = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  KnnProfile* prof = nullptr;  // k-NN kernel timing (owned by the context)
  ~VgicpDriver();

  void evaluate(const float4* src, const double* covs, const uint32_t* d_offs, size_t max_pts, const VgicpTarget& tgt, int count,
                bool profile, cudaStream_t s);
  // single-scan registration; src_covs must have been computed (compute_source_covs)
  int align(const float4* src, size_t ns, const VgicpTarget& tgt, const pcr_params& prm, double* T, int32_t* converged, int32_t* iters,
            bool profile, cudaStream_t s);
  int compute_source_covs(const float4* src, size_t ns, int k, KeySort& ks, BBoxWork& bw, cudaStream_t s);
  // mean 1-NN squared distance of T*src in the target (float metric), FP64 accumulate
  int fitness(const float4* src, size_t ns, const VgicpTarget& tgt, const double* T, double max_range, double* score, cudaStream_t s);
};

// exact k-NN (float metric, (d2, idx) order) + PLANE-regularised covariance for every point of `pts` using `grid` built over it.
// covs: 6 doubles per point. knn_idx: device scratch/output, k ints per point.
void gicp_covariances(const float4* pts, size_t n, const MortonGrid& grid, int k, double* covs, int32_t* knn_idx, cudaStream_t s,
                      KnnProfile* prof = nullptr, bool sorted_idx = false);

int vgicp_build_target(const float4* pts, size_t n, const pcr_params& prm, VgicpTarget& tgt, KeySort& ks, BBoxWork& bw, cudaStream_t s);

}  // namespace pcr
