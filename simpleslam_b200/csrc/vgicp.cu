// FastVGICP on the GPU. Reference: third_parties/pclomp/src/fast_gicp_impl.hpp, fast_vgicp_impl.hpp,
// pclomp/fast_vgicp_voxel.hpp, lsq_registration_impl.hpp, so3/so3.hpp.
#include "vgicp.cuh"
#include "dev_linalg.cuh"
#include "host_math.hpp"
#include <cfloat>
#include <cstdlib>

namespace pcr {

constexpr int kVgBlock = 128;
constexpr int kVgNV = 29;  // cost, 21 H, 6 b, count
constexpr int kMaxK = kKnnMaxK;

// ================================================================================================================
// V1. FastGICP::calculate_covariances (fast_gicp_impl.hpp:241-298): k-NN (self included), cov = N N^T / k of the
// mean-centred neighbours (FP64), PLANE regularisation U diag(1,1,1e-3) V^T.
// Two kernels: warp-per-query k-NN writes the neighbour indices, then one thread per point does the 3x3 algebra.
// ================================================================================================================
// SORT: neighbours written in ascending (d2, index) order (what the parity tests compare); the covariance itself does not
// depend on the order beyond FP64 rounding, so the registration path skips the final sort.
template <bool SORT>
__global__ void __launch_bounds__(256)
gicp_knn_kernel(size_t n, MortonView grid, int k, int min_pop, int32_t* __restrict__ knn_idx) {
  // one warp per query; queries are taken in Morton order (the sorted copy), so neighbouring warps touch the same cells
  const int lane = threadIdx.x & 31;
  const size_t warps = size_t(gridDim.x) * (blockDim.x >> 5);
  for (size_t i = size_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
    const float4 q = __ldg(grid.pts + i);
    WarpKnn st = knn_warp_morton(grid, q.x, q.y, q.z, k, min_pop, lane);
    if (SORT) knn_sort_result(st, lane);  // ascending (d2, index): the order pcl::search::KdTree::nearestKSearch returns
    const size_t orig = size_t(__float_as_int(q.w));
    if (lane < k) knn_idx[orig * k + lane] = lane < st.cnt ? st.bi : -1;
  }
}

__global__ void __launch_bounds__(128)
gicp_cov_kernel(const float4* __restrict__ pts, size_t n, int k, const int32_t* __restrict__ knn_idx, double* __restrict__ covs) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t* nb = knn_idx + i * k;
  double mean[3] = {0, 0, 0};
  for (int j = 0; j < k; j++) {
    const int id = nb[j];
    if (id < 0) break;
    const float4 p = __ldg(pts + id);
    mean[0] += double(p.x); mean[1] += double(p.y); mean[2] += double(p.z);
  }
  const double dk = double(k);
  mean[0] /= dk; mean[1] /= dk; mean[2] /= dk;
  double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < k; j++) {
    const int id = nb[j];
    if (id < 0) break;
    const float4 p = __ldg(pts + id);
    const double d[3] = {double(p.x) - mean[0], double(p.y) - mean[1], double(p.z) - mean[2]};
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = r; c < 3; c++) cov[r][c] += d[r] * d[c];
  }
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = r; c < 3; c++) { cov[r][c] /= dk; cov[c][r] = cov[r][c]; }
  double w[3], V[3][3];
  eig_sym3(cov, w, V);  // ascending; the SVD's descending singular values get (1, 1, 1e-3)
  const double vals[3] = {1e-3, 1.0, 1.0};
  double* o = covs + i * 6;
  int t = 0;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = r; c < 3; c++) {
      double v = 0;
#pragma unroll
      for (int e = 2; e >= 0; e--) v += (V[r][e] * vals[e]) * V[c][e];
      o[t++] = v;
    }
}

KnnProfile::~KnnProfile() {
  for (auto& p : pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
}
void KnnProfile::begin(cudaStream_t s) {
  cudaEvent_t a = nullptr, b = nullptr;
  PCR_CUDA_CHECK(cudaEventCreate(&a));
  PCR_CUDA_CHECK(cudaEventCreate(&b));
  pending.emplace_back(a, b);
  PCR_CUDA_CHECK(cudaEventRecord(a, s));
}
void KnnProfile::end(cudaStream_t s, size_t n) {
  PCR_CUDA_CHECK(cudaEventRecord(pending.back().second, s));
  launches++;
  queries += (long long)n;
}
void KnnProfile::collect() {
  for (auto& p : pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.first, p.second) == cudaSuccess) ms += t;
    cudaEventDestroy(p.first); cudaEventDestroy(p.second);
  }
  pending.clear();
}

// knn_idx: device scratch of n*k ints (always needed)
void gicp_covariances(const float4* pts, size_t n, const MortonGrid& grid, int k, double* covs, int32_t* knn_idx, cudaStream_t s, KnnProfile* prof,
                      bool sorted_idx) {
  if (n == 0) return;
  const unsigned blocks = unsigned(std::min<size_t>((n + 7) / 8, size_t(kNumSMs) * 64));
  static const int min_pop_env = std::getenv("PCR_KNN_MINPOP") ? std::atoi(std::getenv("PCR_KNN_MINPOP")) : 0;  // tuning knob
  const int min_pop = min_pop_env > 0 ? min_pop_env : std::max(1, k / 2);  // sweep on C3 (profiles/README.md): flat from k/2.5 to 0.6 k
  if (prof) prof->begin(s);
  if (sorted_idx) gicp_knn_kernel<true><<<blocks, 256, 0, s>>>(n, view_of(grid), k, min_pop, knn_idx);
  else gicp_knn_kernel<false><<<blocks, 256, 0, s>>>(n, view_of(grid), k, min_pop, knn_idx);
  if (prof) prof->end(s, n);
  gicp_cov_kernel<<<unsigned((n + 127) / 128), 128, 0, s>>>(pts, n, k, knn_idx, covs);
}

// ================================================================================================================
// V2. GaussianVoxelMap::create_voxelmap, ADDITIVE (fast_vgicp_voxel.hpp:105-174): coord = floor(x / res - 0.5) in FP64
// ================================================================================================================
__device__ __forceinline__ int vg_coord(double x, double res) { return int(floor(__dsub_rn(__ddiv_rn(x, res), 0.5))); }

__global__ void __launch_bounds__(256)
vg_key_kernel(const float4* __restrict__ pts, size_t n, double res, int c0, int c1, int c2, int d0, int d1, uint32_t* __restrict__ keys) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + i);
  const int x = vg_coord(double(p.x), res) - c0, y = vg_coord(double(p.y), res) - c1, z = vg_coord(double(p.z), res) - c2;
  keys[i] = uint32_t(x) + uint32_t(d0) * (uint32_t(y) + uint32_t(d1) * uint32_t(z));
}

__global__ void __launch_bounds__(256)
vg_voxel_kernel(const float4* __restrict__ pts, const double* __restrict__ covs, const uint32_t* __restrict__ keys,
                const uint32_t* __restrict__ vals, const uint32_t* __restrict__ seg_start, size_t nseg, VoxelRec* __restrict__ vox,
                int32_t* __restrict__ vox_key, int32_t* __restrict__ table) {
  // one warp per voxel: lanes stride over the voxel's points (ascending original index), fixed-order warp reduction.
  // (The reference appends serially; the sums agree to FP64 rounding and are reproducible run to run.)
  const int lane = threadIdx.x & 31;
  const size_t v = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5;
  if (v >= nseg) return;
  const uint32_t b = seg_start[v], e = seg_start[v + 1];
  double m[3] = {0, 0, 0}, c[6] = {0, 0, 0, 0, 0, 0};
  // a voxel next to the sensor holds thousands of points: the indices of four trips are fetched together so that four
  // independent point / covariance gathers are in flight per lane (the summation order per lane is unchanged)
  for (uint32_t j0 = b + lane; j0 < e; j0 += 128) {
    uint32_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; u++) idx[u] = (j0 + 32 * u < e) ? vals[j0 + 32 * u] : 0xffffffffu;
    float4 p[4];
    double2 ca[4][3];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (idx[u] == 0xffffffffu) continue;
      p[u] = __ldg(pts + idx[u]);
      const double2* cp = reinterpret_cast<const double2*>(covs + size_t(idx[u]) * 6);
      ca[u][0] = __ldg(cp); ca[u][1] = __ldg(cp + 1); ca[u][2] = __ldg(cp + 2);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (idx[u] == 0xffffffffu) continue;
      m[0] += double(p[u].x); m[1] += double(p[u].y); m[2] += double(p[u].z);
      c[0] += ca[u][0].x; c[1] += ca[u][0].y; c[2] += ca[u][1].x; c[3] += ca[u][1].y; c[4] += ca[u][2].x; c[5] += ca[u][2].y;
    }
  }
#pragma unroll
  for (int t = 0; t < 3; t++) m[t] = warp_sum(m[t]);
#pragma unroll
  for (int t = 0; t < 6; t++) c[t] = warp_sum(c[t]);
  if (lane != 0) return;
  const double dn = double(e - b);
  VoxelRec r;
  for (int t = 0; t < 3; t++) r.mean[t] = m[t] / dn;
  for (int t = 0; t < 6; t++) r.cov[t] = c[t] / dn;
  r.n = dn;
  vox[v] = r;
  const uint32_t key = keys[b];
  vox_key[v] = int32_t(key);
  table[key] = int32_t(v);
}

int vgicp_build_target(const float4* pts, size_t n, const pcr_params& prm, VgicpTarget& tgt, KeySort& ks, BBoxWork& bw, cudaStream_t s) {
  tgt.built = false;
  tgt.n = n;
  tgt.nvox = 0;
  tgt.resolution = prm.vgicp_resolution;
  if (n == 0) { tgt.built = true; return 0; }
  if (prm.vgicp_k > kMaxK || prm.vgicp_k < 1) return PCR_ERR_INVALID;
  // nested kNN grid + per-point covariances
  int rc = build_morton_grid(pts, n, tgt.grid, bw, s);
  if (rc) return rc;
  tgt.covs.ensure(n * 6);
  tgt.knn.ensure(n * size_t(prm.vgicp_k));
  gicp_covariances(pts, n, tgt.grid, prm.vgicp_k, tgt.covs.p, tgt.knn.p, s, tgt.prof);
  // voxel map
  float mn[3], mx[3];
  bbox_blocking(pts, n, mn, mx, bw, s);
  long long ncell = 1;
  for (int a = 0; a < 3; a++) {
    tgt.cmin[a] = int(std::floor(double(mn[a]) / tgt.resolution - 0.5));
    const int cmax = int(std::floor(double(mx[a]) / tgt.resolution - 0.5));
    tgt.cdim[a] = cmax - tgt.cmin[a] + 1;
    ncell *= tgt.cdim[a];
    if (ncell > (1ll << 29)) return PCR_ERR_GRID_TOO_LARGE;
  }
  tgt.ncell = ncell;
  ks.k0.ensure(n);
  vg_key_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(pts, n, tgt.resolution, tgt.cmin[0], tgt.cmin[1], tgt.cmin[2], tgt.cdim[0],
                                                         tgt.cdim[1], ks.k0.p);
  ks.sort_keys_in_k0(n, ncell, s);
  ks.segment(s);
  tgt.nvox = ks.nseg;
  tgt.vox.ensure(tgt.nvox);
  tgt.vox_key.ensure(tgt.nvox);
  tgt.table.ensure(size_t(ncell));
  PCR_CUDA_CHECK(cudaMemsetAsync(tgt.table.p, 0xff, size_t(ncell) * sizeof(int32_t), s));
  vg_voxel_kernel<<<unsigned((tgt.nvox * 32 + 255) / 256), 256, 0, s>>>(pts, tgt.covs.p, ks.keys, ks.vals, ks.seg_start.p, tgt.nvox, tgt.vox.p,
                                                                  tgt.vox_key.p, tgt.table.p);
  PCR_CUDA_CHECK(cudaGetLastError());
  tgt.built = true;
  return 0;
}

// ================================================================================================================
// V3 + V4. update_correspondences (DIRECT1) + linearize / compute_error fused: one thread per source point.
// Correspondence and Mahalanobis matrix come from T0 (the linearisation point, fast_vgicp_impl.hpp:73-116);
// the error is evaluated at Ti (Ti = T0 for linearize, Ti = delta*T0 for compute_error, :183-204).
// ================================================================================================================
struct VgTargetView {
  const VoxelRec* vox;
  const int32_t* table;
  double res;
  int cmin[3], cdim[3];
};

__device__ __forceinline__ void xform_exact(const double* T, double x, double y, double z, double* o) {
#pragma unroll
  for (int r = 0; r < 3; r++)
    o[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[r], x), __dmul_rn(T[4 + r], y)), __dmul_rn(T[8 + r], z)), T[12 + r]);
}

__device__ __noinline__ void vgicp_state_step(VgicpState* st, const double* totals, VgicpCfg cfg) { vgicp_logic::on_result(*st, totals, cfg); }

// One evaluation (linearize, or the error of an LM trial) per launch and request. `states` (nullable): the request's parameters
// come from the scan's state machine and the last block advances it by the result (vgicp_logic.cuh): the host only queues
// launches. states == nullptr: explicit parameters (introspection entry point).
__global__ void __launch_bounds__(kVgBlock, 4)
vgicp_eval_kernel(const float4* __restrict__ src, const double* __restrict__ src_covs, const uint32_t* __restrict__ offs, VgTargetView tgt,
                  const VgicpEvalParams* __restrict__ params, VgicpEvalResult* __restrict__ results, double* __restrict__ partials,
                  unsigned* __restrict__ tickets, int max_blocks, VgicpState* __restrict__ states, VgicpCfg cfg, VgicpProgress* progress,
                  int round) {
  const int req = blockIdx.y;
  __shared__ VgicpEvalParams sp;
  __shared__ double sred[kVgNV * (kVgBlock / 32)];
  __shared__ double s_tot[kVgNV + 1];
  __shared__ int s_last, s_pend;
  if (states && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && progress) progress->round = round;
  {
    const int nwords = sizeof(VgicpEvalParams) / 4;
    const int* gp = reinterpret_cast<const int*>(states ? &states[req].next : params + req);
    int* spw = reinterpret_cast<int*>(&sp);
    for (int k = threadIdx.x; k < nwords; k += kVgBlock) spw[k] = gp[k];
    if (threadIdx.x == 0) s_pend = states ? states[req].pend : 1;
  }
  __syncthreads();
  if (!s_pend) return;  // finished earlier: a launch queued past the end of the registration
  const uint32_t begin = offs[sp.scan], end = offs[sp.scan + 1];
  const int nb = min(int((end - begin + kVgBlock - 1) / kVgBlock), max_blocks);
  if (int(blockIdx.x) >= nb) return;

  // points strided over one resident wave of blocks: the 29-value block reduction is paid once per thread, not per point
  double acc[kVgNV];
#pragma unroll
  for (int k = 0; k < kVgNV; k++) acc[k] = 0.0;
  for (uint32_t i = begin + blockIdx.x * kVgBlock + threadIdx.x; i < end; i += uint32_t(nb) * kVgBlock) {
    const float4 p = __ldg(src + i);
    const double px = double(p.x), py = double(p.y), pz = double(p.z);
    double t0[3];
    xform_exact(sp.T0, px, py, pz, t0);
    const int cx = vg_coord(t0[0], tgt.res) - tgt.cmin[0], cy = vg_coord(t0[1], tgt.res) - tgt.cmin[1],
              cz = vg_coord(t0[2], tgt.res) - tgt.cmin[2];
    if (cx >= 0 && cx < tgt.cdim[0] && cy >= 0 && cy < tgt.cdim[1] && cz >= 0 && cz < tgt.cdim[2]) {
      const long long key = cx + (long long)tgt.cdim[0] * (cy + (long long)tgt.cdim[1] * cz);
      const int vid = __ldg(tgt.table + key);
      if (vid >= 0) {
        const double2* vp = reinterpret_cast<const double2*>(tgt.vox + vid);
        const double2 v0 = __ldg(vp), v1 = __ldg(vp + 1), v2 = __ldg(vp + 2), v3 = __ldg(vp + 3), v4 = __ldg(vp + 4);
        const double mB[3] = {v0.x, v0.y, v1.x};
        const double cB[6] = {v1.y, v2.x, v2.y, v3.x, v3.y, v4.x};
        const double wgt = sqrt(v4.y);
        const double* ca = src_covs + size_t(i) * 6;
        const double A[3][3] = {{ca[0], ca[1], ca[2]}, {ca[1], ca[3], ca[4]}, {ca[2], ca[4], ca[5]}};
        // RCR = C_B + R C_A R^T (upper 3x3 of the reference's 4x4; the (3,3)=1 row/col decouples)
        double RC[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) RC[r][c] = (sp.T0[r] * A[0][c] + sp.T0[4 + r] * A[1][c]) + sp.T0[8 + r] * A[2][c];
        double S[3][3];
        S[0][0] = cB[0] + ((RC[0][0] * sp.T0[0] + RC[0][1] * sp.T0[4]) + RC[0][2] * sp.T0[8]);
        S[0][1] = cB[1] + ((RC[0][0] * sp.T0[1] + RC[0][1] * sp.T0[5]) + RC[0][2] * sp.T0[9]);
        S[0][2] = cB[2] + ((RC[0][0] * sp.T0[2] + RC[0][1] * sp.T0[6]) + RC[0][2] * sp.T0[10]);
        S[1][1] = cB[3] + ((RC[1][0] * sp.T0[1] + RC[1][1] * sp.T0[5]) + RC[1][2] * sp.T0[9]);
        S[1][2] = cB[4] + ((RC[1][0] * sp.T0[2] + RC[1][1] * sp.T0[6]) + RC[1][2] * sp.T0[10]);
        S[2][2] = cB[5] + ((RC[2][0] * sp.T0[2] + RC[2][1] * sp.T0[6]) + RC[2][2] * sp.T0[10]);
        S[1][0] = S[0][1]; S[2][0] = S[0][2]; S[2][1] = S[1][2];
        double M[3][3];
        inv3(S, M);
        double tA[3];
        xform_exact(sp.Ti, px, py, pz, tA);
        const double e[3] = {mB[0] - tA[0], mB[1] - tA[1], mB[2] - tA[2]};
        double Me[3];
#pragma unroll
        for (int r = 0; r < 3; r++) Me[r] = (M[r][0] * e[0] + M[r][1] * e[1]) + M[r][2] * e[2];
        acc[0] += wgt * ((e[0] * Me[0] + e[1] * Me[1]) + e[2] * Me[2]);
        acc[28] += 1.0;
        if (sp.want_hb) {
          // J = [ skew(T p) | -I ] : columns
          const double J[6][3] = {{0.0, tA[2], -tA[1]}, {-tA[2], 0.0, tA[0]}, {tA[1], -tA[0], 0.0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
          double MJ[6][3];
#pragma unroll
          for (int c = 0; c < 6; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) MJ[c][r] = (M[r][0] * J[c][0] + M[r][1] * J[c][1]) + M[r][2] * J[c][2];
          int k = 1;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = r; c < 6; c++) acc[k++] += wgt * ((J[r][0] * MJ[c][0] + J[r][1] * MJ[c][1]) + J[r][2] * MJ[c][2]);
#pragma unroll
          for (int r = 0; r < 6; r++) acc[22 + r] += wgt * ((J[r][0] * Me[0] + J[r][1] * Me[1]) + J[r][2] * Me[2]);
        }
      }
    }
  }
  double r = block_reduce_vec<kVgNV, kVgBlock>(acc, sred);
  double* my = partials + (size_t(req) * max_blocks + blockIdx.x) * kVgNV;
  if (threadIdx.x < kVgNV) my[threadIdx.x] = r;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(tickets + req, 1u);
    s_last = (t == unsigned(nb - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // totals in a fixed order: four interleaved slices of the block rows per component (a step of the loop reads four
  // consecutive rows, coalesced), two accumulators each — one thread per component walking all ~590 rows was a quarter of
  // the kernel's 37 us
  static_assert(4 * kVgNV <= kVgBlock, "four slices per component");
  if (threadIdx.x < 4 * kVgNV) {
    const int comp = threadIdx.x % kVgNV, slice = threadIdx.x / kVgNV;
    const double* pb = partials + size_t(req) * max_blocks * kVgNV + comp;
    double t0 = 0.0, t1 = 0.0;
    int b = slice;
    for (; b + 4 < nb; b += 8) {
      t0 += __ldcg(pb + size_t(b) * kVgNV);
      t1 += __ldcg(pb + size_t(b + 4) * kVgNV);
    }
    if (b < nb) t0 += __ldcg(pb + size_t(b) * kVgNV);
    sred[slice * kVgNV + comp] = t0 + t1;
  }
  __syncthreads();
  if (threadIdx.x < kVgNV) {
    const int c = threadIdx.x;
    const double tsum = (sred[c] + sred[kVgNV + c]) + (sred[2 * kVgNV + c] + sred[3 * kVgNV + c]);
    results[req].v[c] = tsum;
    s_tot[c] = tsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) tickets[req] = 0;
  if (!states) return;
  // every block of this request has read its parameters (they all took a ticket): the state may move on. It is staged in
  // shared memory so that the single thread that runs the scalar control flow does not chase global memory.
  __shared__ __align__(16) VgicpState s_state;
  {
    const int nwords = sizeof(VgicpState) / 16;
    const uint4* gp = reinterpret_cast<const uint4*>(states + req);
    uint4* sp4 = reinterpret_cast<uint4*>(&s_state);
    for (int k = threadIdx.x; k < nwords; k += kVgBlock) sp4[k] = gp[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) vgicp_state_step(&s_state, s_tot, cfg);
  __syncthreads();
  {
    const int nwords = sizeof(VgicpState) / 16;
    uint4* gp = reinterpret_cast<uint4*>(states + req);
    const uint4* sp4 = reinterpret_cast<const uint4*>(&s_state);
    for (int k = threadIdx.x; k < nwords; k += kVgBlock) gp[k] = sp4[k];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && s_state.pend == 0 && progress) { __threadfence_system(); progress->done = 1; }
}

static_assert(sizeof(VgicpState) % 16 == 0, "VgicpState is copied as uint4 words");

VgicpDriver::~VgicpDriver() {
  if (progress) cudaFreeHost(progress);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
}

void VgicpDriver::evaluate(const float4* src, const double* covs, const uint32_t* d_offs, size_t max_pts, const VgicpTarget& tgt, int count,
                           bool profile, cudaStream_t s) {
  if (count == 0) return;
  // 4 blocks of 128 threads are resident per SM (126 registers): one resident wave shared by the requests
  int max_blocks = std::min(int((max_pts + kVgBlock - 1) / kVgBlock), std::max(1, (kNumSMs * 4) / count));
  if (max_blocks < 1) max_blocks = 1;
  d_params.ensure(count);
  d_results.ensure(count);
  partials.ensure(size_t(count) * max_blocks * kVgNV);
  if (tickets.cap < size_t(count)) {
    tickets.ensure(count);
    PCR_CUDA_CHECK(cudaMemsetAsync(tickets.p, 0, tickets.cap * sizeof(unsigned), s));
  }
  PCR_CUDA_CHECK(cudaMemcpyAsync(d_params.p, h_params.p, size_t(count) * sizeof(VgicpEvalParams), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemsetAsync(d_results.p, 0, size_t(count) * sizeof(VgicpEvalResult), s));
  const bool run = tgt.nvox > 0 && max_pts > 0;
  if (run) {
    VgTargetView v;
    v.vox = tgt.vox.p; v.table = tgt.table.p; v.res = tgt.resolution;
    for (int a = 0; a < 3; a++) { v.cmin[a] = tgt.cmin[a]; v.cdim[a] = tgt.cdim[a]; }
    if (profile) {
      if (!ev0) { PCR_CUDA_CHECK(cudaEventCreate(&ev0)); PCR_CUDA_CHECK(cudaEventCreate(&ev1)); }
      PCR_CUDA_CHECK(cudaEventRecord(ev0, s));
    }
    vgicp_eval_kernel<<<dim3(max_blocks, count), kVgBlock, 0, s>>>(src, covs, d_offs, v, d_params.p, d_results.p, partials.p, tickets.p,
                                                                   max_blocks, nullptr, VgicpCfg{}, nullptr, 0);
    if (profile) PCR_CUDA_CHECK(cudaEventRecord(ev1, s));
    launches++;
    hot_launches++;
  }
  h_results.ensure(count);
  PCR_CUDA_CHECK(cudaMemcpyAsync(h_results.p, d_results.p, size_t(count) * sizeof(VgicpEvalResult), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  PCR_CUDA_CHECK(cudaGetLastError());
  if (profile && run) {
    float ms = 0.f;
    PCR_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    hot_ms += ms;
  }
}

int VgicpDriver::compute_source_covs(const float4* src, size_t ns, int k, KeySort& ks, BBoxWork& bw, cudaStream_t s) {
  if (ns == 0) return 0;
  if (k > kMaxK || k < 1) return PCR_ERR_INVALID;
  int rc = build_morton_grid(src, ns, src_grid, bw, s);
  if (rc) return rc;
  src_covs.ensure(ns * 6);
  knn_dbg.ensure(ns * size_t(k));
  gicp_covariances(src, ns, src_grid, k, src_covs.p, knn_dbg.p, s, prof);
  launches += 4;
  return 0;
}

// ================================================================================================================
// V5. LsqRegistration::computeTransformation / step_lm / step_gn / is_converged (lsq_registration_impl.hpp:53-172)
// ================================================================================================================
int VgicpDriver::align(const float4* src, size_t ns, const VgicpTarget& tgt, const pcr_params& prm, double* T, int32_t* converged,
                       int32_t* iters, bool profile, cudaStream_t s) {
  hot_ms = 0.f; hot_launches = 0; n_linearize = 0; n_error = 0; total_corr = 0;
  VgicpCfg cfg{};
  cfg.optimizer = prm.vgicp_optimizer == PCR_LSQ_GN ? 1 : 0;
  cfg.max_iters = prm.vgicp_max_iters;
  cfg.lm_max_iters = prm.vgicp_lm_max_iters;
  cfg.rot_eps = prm.vgicp_rot_eps; cfg.trans_eps = prm.vgicp_trans_eps; cfg.lm_init_lambda = prm.vgicp_lm_init_lambda;
  VgicpState* hs = h_states.ensure(1);
  memset(hs, 0, sizeof(VgicpState));
  vgicp_logic::start(*hs, T, 0, cfg);
  const bool run = tgt.nvox > 0 && ns > 0;
  if (!run) {
    // no voxel / no point: every evaluation is all zeros (no block would contribute) — the state machine digests them here
    const double zeros[kVgNV + 1] = {0};
    for (int guard = 0; hs->pend && guard < 100000; guard++) vgicp_logic::on_result(*hs, zeros, cfg);
  } else {
    if (!progress) PCR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&progress), sizeof(VgicpProgress), cudaHostAllocMapped));
    progress->round = -1;
    progress->done = 0;
    VgicpProgress* dprog = nullptr;
    PCR_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dprog), progress, 0));
    uint32_t* ho = h_offsets.ensure(2);
    ho[0] = 0; ho[1] = uint32_t(ns);
    offsets.ensure(2);
    d_states.ensure(1);
    d_results.ensure(1);
    PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    PCR_CUDA_CHECK(cudaMemcpyAsync(d_states.p, hs, sizeof(VgicpState), cudaMemcpyHostToDevice, s));
    // 4 blocks of 128 threads are resident per SM (126 registers): one resident wave
    const int max_blocks = std::max(1, std::min(int((ns + kVgBlock - 1) / kVgBlock), kNumSMs * 4));
    partials.ensure(size_t(max_blocks) * kVgNV);
    if (tickets.cap < 1) {
      tickets.ensure(1);
      PCR_CUDA_CHECK(cudaMemsetAsync(tickets.p, 0, tickets.cap * sizeof(unsigned), s));
    }
    VgTargetView v;
    v.vox = tgt.vox.p; v.table = tgt.table.p; v.res = tgt.resolution;
    for (int a = 0; a < 3; a++) { v.cmin[a] = tgt.cmin[a]; v.cdim[a] = tgt.cdim[a]; }
    if (profile) {
      if (!ev0) { PCR_CUDA_CHECK(cudaEventCreate(&ev0)); PCR_CUDA_CHECK(cudaEventCreate(&ev1)); }
      PCR_CUDA_CHECK(cudaEventRecord(ev0, s));
    }
    // The host only queues evaluation launches, a few ahead of the GPU (round counter and done flag in host-mapped memory);
    // launches queued past the end return on their first instructions.
    static const int lookahead = [] { const char* e = std::getenv("PCR_VGICP_LOOKAHEAD"); return e ? std::max(1, std::atoi(e)) : 3; }();
    const int max_rounds = std::max(1, prm.vgicp_max_iters) * (std::max(0, prm.vgicp_lm_max_iters) + 1) + 2;
    int r = 0;
    unsigned spins = 0;
    bool idle = false;
    while (r < max_rounds) {
      while (!progress->done && progress->round < r - lookahead) {
        if ((++spins & 0xfff) == 0) {
          const cudaError_t q = cudaStreamQuery(s);
          if (q == cudaSuccess) { idle = true; break; }
          if (q != cudaErrorNotReady) PCR_CUDA_CHECK(q);
        }
      }
      if (progress->done || idle) break;
      vgicp_eval_kernel<<<dim3(max_blocks, 1), kVgBlock, 0, s>>>(src, src_covs.p, offsets.p, v, nullptr, d_results.p, partials.p, tickets.p, max_blocks,
                                                                 d_states.p, cfg, dprog, r);
      launches++;
      r++;
    }
    if (profile) PCR_CUDA_CHECK(cudaEventRecord(ev1, s));
    PCR_CUDA_CHECK(cudaMemcpyAsync(hs, d_states.p, sizeof(VgicpState), cudaMemcpyDeviceToHost, s));
    PCR_CUDA_CHECK(cudaStreamSynchronize(s));
    PCR_CUDA_CHECK(cudaGetLastError());
    if (hs->pend) throw CudaError("VGICP: the evaluation launches ended before the registration finished");
    if (profile) {
      float ms = 0.f;
      PCR_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
      hot_ms += ms;
    }
  }
  n_linearize = hs->n_linearize; n_error = hs->n_error; total_corr = hs->total_corr;
  hot_launches = n_linearize + n_error;
  last_cost = hs->last_cost; last_corr = hs->last_corr;
  for (int i = 0; i < 16; i++) T[i] = double(static_cast<float>(hs->x0[i]));  // final_transformation_ = x0.cast<float>()
  if (converged) *converged = hs->converged;
  if (iters) *iters = hs->nr_iterations;
  return 0;
}

// ================================================================================================================
// V6. pcl::Registration::getFitnessScore: float transform, exact 1-NN (float metric), mean of d2 <= max_range (FP64)
// ================================================================================================================
__global__ void __launch_bounds__(256)
fitness_kernel(const float4* __restrict__ src, size_t ns, MortonView grid, const float* __restrict__ Tf, double max_range,
               double* __restrict__ partials) {
  // one warp per query; per-warp sums in FP64, fixed-order block reduction -> partials[block] = {sum d2, count}
  __shared__ double ssum[8], scnt[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t warps = size_t(gridDim.x) * 8;
  double sum = 0.0, cnt = 0.0;
  for (size_t i = size_t(blockIdx.x) * 8 + warp; i < ns; i += warps) {
    const float4 p = __ldg(src + i);
    float q[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
      q[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[r], p.x), __fmul_rn(Tf[4 + r], p.y)), __fmul_rn(Tf[8 + r], p.z)), Tf[12 + r]);
    const WarpKnn st = knn_warp_morton(grid, q[0], q[1], q[2], 1, 1, lane);
    if (st.cnt == 1 && double(st.td) <= max_range) { sum += double(st.td); cnt += 1.0; }
  }
  if (lane == 0) { ssum[warp] = sum; scnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) { a += ssum[w]; b += scnt[w]; }
    partials[size_t(blockIdx.x) * 2] = a;
    partials[size_t(blockIdx.x) * 2 + 1] = b;
  }
}

int VgicpDriver::fitness(const float4* src, size_t ns, const VgicpTarget& tgt, const double* T, double max_range, double* score,
                         cudaStream_t s) {
  *score = DBL_MAX;
  if (ns == 0 || tgt.n == 0 || !tgt.grid.built) return 0;
  const unsigned blocks = unsigned(std::min<size_t>((ns + 7) / 8, size_t(kNumSMs) * 16));
  fit_partials.ensure(size_t(blocks) * 2 + 16);
  float* dTf = reinterpret_cast<float*>(fit_partials.p + size_t(blocks) * 2);
  float hT[16];
  for (int i = 0; i < 16; i++) hT[i] = static_cast<float>(T[i]);
  PCR_CUDA_CHECK(cudaMemcpyAsync(dTf, hT, sizeof(hT), cudaMemcpyHostToDevice, s));
  fitness_kernel<<<blocks, 256, 0, s>>>(src, ns, view_of(tgt.grid), dTf, max_range, fit_partials.p);
  launches++;
  double* h = h_fit.ensure(size_t(blocks) * 2);
  PCR_CUDA_CHECK(cudaMemcpyAsync(h, fit_partials.p, size_t(blocks) * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  PCR_CUDA_CHECK(cudaGetLastError());
  double sum = 0, cnt = 0;
  for (unsigned b = 0; b < blocks; b++) { sum += h[b * 2]; cnt += h[b * 2 + 1]; }
  if (cnt > 0) *score = sum / cnt;
  return 0;
}

}  // namespace pcr
