// LOAM-style point-to-plane scan-to-map registration on the GPU.
#pragma once
#include "common.cuh"
#include "voxel.cuh"
#include "../../include/pcr_cuda.h"

namespace pcr {

struct LoamParams {
  double max_knn_d2, plane_thresh, point_thresh, pos_conv, rot_conv;
  int max_iters;
};

struct LoamState {  // one per scan, device resident
  double T[16];
  int done, converged, iters, n_last;
  unsigned ticket;
  int pad;
  long long cand_total;  // map points examined by the neighbour search, summed over iterations
  long long rows_total;  // x-rows of cells looked up in the start table (two 4-byte entries each), summed over iterations
  long long pt_evals;    // source points linearised, summed over iterations
};

// Target index of the LOAM search: uniform grid with a dense start table. Cell width = gate radius (one 27-cell ring
// covers the gate ball) for sparse maps, half of it (second ring on demand) when the map is dense — picked from the mean
// number of points per occupied cell. Returns a PCR error code.
int loam_build_target(const float4* pts, size_t n, double max_knn_d2, CellGrid& grid, KeySort& ks, BBoxWork& bw, cudaStream_t s);

struct LoamDriver {
  DevBuf<LoamState> states;
  DevBuf<double> partials;
  DevBuf<pcr_loam_iter_log> logs;
  DevBuf<uint32_t> offsets;
  DevBuf<int32_t> dbg_knn, dbg_status;
  DevBuf<float4> nb_buf;    // batch path: the five winners of every query as coordinates (five planes, w = original map index)
  DevBuf<int2> cnt_buf;     // batch path: candidates examined / rows walked per query
  // split mode: queries re-ordered along a Morton curve of their scan-frame position (once per call)
  DevBuf<unsigned long long> q_keys0, q_keys1;
  DevBuf<uint32_t> q_vals0, q_vals1;
  DevBuf<float4> q_sorted;
  DevBuf<unsigned char> q_tmp;
  PinBuf<LoamState> h_states;
  PinBuf<uint32_t> h_offsets;
  PinBuf<pcr_loam_iter_log> h_logs;
  int last_log_count = 0;
  long long launches = 0;
  long long cand_total = 0;
  long long rows_total = 0;
  long long pt_evals = 0;
  float hot_ms = 0.f;
  int hot_launches = 0;
  int last_lpq = 0, last_tile = 0;  // kernel shape of the last align / linearize call (introspection: pcr_loam_last_shape)
  bool last_split = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  ~LoamDriver();
  // src: device float4 points of all scans, concatenated; offs: host offsets [n_scans+1].
  // T: host, 16*n_scans doubles (column-major) in/out. Returns 0.
  int align(const float4* src, const size_t* offs, size_t n_scans, const CellGrid& grid, const LoamParams& prm, double* T,
            int32_t* converged, int32_t* iters_out, int64_t* n_last_out, bool profile, cudaStream_t s);
  // one linearisation at pose T for a single scan; outputs are host pointers (nullable)
  // Morton re-ordering of the queries of every scan: returns the sorted copy, *perm = sorted position -> original position
  const float4* sort_queries(const float4* src, const uint32_t* d_offs, size_t n_scans, size_t total_q, size_t max_pts, const uint32_t** perm,
                             cudaStream_t s);
  int linearize(const float4* src, size_t ns, const CellGrid& grid, const LoamParams& prm, const double* T, int32_t* knn_idx,
                int32_t* status, double* JtJ, double* JtE, int64_t* n_acc, cudaStream_t s);
};

}  // namespace pcr
