// pclomp NDT on the GPU: voxel-covariance target build (N1), neighbourhood lookup (N2/N6), derivative kernels (N3/N4)
// and the Newton + More-Thuente control flow (N5) as a device-side state machine (ndt_logic.cuh), batched over
// independent scans: a registration is a stream of evaluation rounds with no host round trip in between.
#pragma once
#include "common.cuh"
#include "voxel.cuh"
#include "ndt_logic.cuh"
#include "../../include/pcr_cuda.h"
#include <functional>

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

struct NdtEvalResult { double v[30]; };  // score, g[6], H upper-triangular 21 (row-major order r<=c), pairs, pad

struct NdtScanOut {  // per scan, written by the tail that finishes the scan (or by a single evaluation without state step)
  float final_T[16];
  int converged, nr_iterations, n_evals, n_hess;
  long long n_pairs;
  double score;
  double v[30];      // sums of the last evaluation (introspection entry points)
};

struct NdtProgress {  // host-mapped pinned memory: the only thing the host looks at while a registration runs
  volatile int round;     // round whose float kernel has started
  volatile int all_done;  // set by the tail that finishes the last scan
  volatile int finished;  // scans finished when the float kernel of `round` started
  int pad[13];
};

struct NdtCounters {  // device
  int finished;
  int work_launches;      // kernel launches that had at least one request
  long long point_evals;  // source points pushed through the evaluation kernels
  unsigned long long t_tail[10];  // NdtCfg::trace: %globaltimer at [4] kernel start, [0] last block found, [1] totals, [2] state step, [3] request filled,
                                  // [5] state stored; block 0 of the launch: [6] request list ready, [7] first item's points done, [8] its partial row written
};

// Scans may JOIN a running batch: a group of scans becomes eligible when its points have arrived on the device (an upload
// on another stream). ready(k) = group k can be admitted now (non-blocking); wait(k) = block until it can;
// event(k) = the CUDA event the registration stream has to wait for before it touches the group's points.
struct NdtArrivals {
  size_t n_groups = 0;
  const size_t* first = nullptr;   // [n_groups + 1] scan index boundaries of the groups
  std::function<bool(size_t)> ready;
  std::function<void(size_t)> wait;
  std::function<cudaEvent_t(size_t)> event;
};

struct NdtDriver {
  DevBuf<NdtScanState> states;
  DevBuf<NdtScanOut> outs;
  DevBuf<int32_t> stamps;  // [2][n]: round in which a scan wants a float evaluation / a double-path Hessian
  DevBuf<double> partials;
  DevBuf<unsigned> tickets;
  DevBuf<uint32_t> offsets;
  DevBuf<double> guesses;
  DevBuf<NdtCounters> counters;
  DevBuf<int> round_flags;  // per round: which kinds of evaluation are wanted (lets idle launches return at once)
  static constexpr int max_rounds_cap = 4096;
  double* gpart_ = nullptr;      // group rows of the two-level reduction (inside `partials`)
  unsigned* gtickets_ = nullptr;  // their tickets (inside `tickets`)
  int max_bpr_ = 1;         // blocks a single scan can use at most: ceil(points / block)
  PinBuf<NdtScanOut> h_outs;
  PinBuf<NdtScanState> h_state;
  PinBuf<uint32_t> h_offsets;
  PinBuf<double> h_guesses;
  PinBuf<NdtCounters> h_counters;
  NdtProgress* progress = nullptr;  // mapped
  long long launches = 0;
  float hot_ms = 0.f;
  int hot_launches = 0;
  int rounds = 0;
  int total_evals = 0, total_hess = 0;
  long long total_pairs = 0;
  long long point_evals = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ~NdtDriver();

  // one evaluation with explicit parameters (introspection: pcr_ndt_derivatives / pcr_ndt_hessian). Blocking.
  void evaluate_one(const float4* src, size_t ns, const NdtTarget& tgt, int search, const NdtEvalParams& ep, NdtEvalResult& out, cudaStream_t s);
  // full registration of n_scans scans (offs: host offsets, n_scans+1)
  int align(const float4* src, const size_t* offs, size_t n_scans, const NdtTarget& tgt, const pcr_params& prm, double* T,
            int32_t* converged, int32_t* iters, double* trans_prob, bool profile, cudaStream_t s, const NdtArrivals* arrivals = nullptr);
  static constexpr size_t max_streaming_scans = 2048;  // arrivals are supported for batches of up to this many scans

 private:
  void ensure_progress();
  void prepare(size_t n_scans, int grid_blocks, cudaStream_t s);
  void launch_round(const float4* src, const NdtTarget& tgt, int search, int n, int round, int step, const NdtCfg& cfg, int grid_blocks,
                    bool float_kernel, bool double_kernel, cudaStream_t s);
};

int ndt_build_target(const float4* pts, size_t n, const pcr_params& prm, NdtTarget& tgt, KeySort& ks, BBoxWork& bw, cudaStream_t s);

}  // namespace pcr
