// LOAM scan-to-map: one fused kernel per Gauss-Newton iteration.
//   transform (FP64, float-rounded) -> exact 5-NN in the 27-cell neighbourhood (FP64 metric, (d2, index) order)
//   -> 5x3 column-pivoted QR plane fit -> validity / weight gates -> residual + SE(3) Jacobian
//   -> warp-shuffle + block + last-block FP64 reduction of the 21 + 6 + 1 normal-equation terms
//   -> 6x6 LDLT solve, convergence test and exp-map pose update by the last block, all on the device.
// Restates PCR/src/LoamRegister.cpp:99-223 (reference CPU path: nanoflann kd-tree + OpenMP + omp critical).
#include "loam.cuh"
#include "dev_linalg.cuh"
#include <cfloat>

namespace pcr {

constexpr int kLoamBlock = 128;
constexpr int kNV = 29;  // 21 upper JtJ + 6 JtE + count + candidates examined

using GridView = CellGridView;

__device__ __forceinline__ bool knn_less(double d, int i, double d2, int i2) { return d < d2 || (d == d2 && i < i2); }

template <bool DEBUG>
__global__ void __launch_bounds__(kLoamBlock)
loam_iter_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, GridView grid, LoamParams prm,
                 LoamState* __restrict__ states, double* __restrict__ partials, int max_blocks,
                 pcr_loam_iter_log* __restrict__ logs, int apply_update, int32_t* __restrict__ dbg_knn,
                 int32_t* __restrict__ dbg_status) {
  const int scan = blockIdx.y;
  const uint32_t begin = offs[scan], end = offs[scan + 1];
  const uint32_t ns = end - begin;
  const int nb = int((ns + kLoamBlock - 1) / kLoamBlock);
  if (int(blockIdx.x) >= nb) return;
  LoamState* st = states + scan;
  if (st->done) return;

  __shared__ double sT[16];
  __shared__ double sred[kNV * (kLoamBlock / 32)];
  __shared__ double stot[kNV];
  __shared__ int s_last;
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
  __syncthreads();

  double acc[kNV];
#pragma unroll
  for (int k = 0; k < kNV; k++) acc[k] = 0.0;

  const uint32_t i = begin + blockIdx.x * kLoamBlock + threadIdx.x;
  if (i < end) {
    const float4 po = __ldg(src + i);
    // LoamRegister.cpp:128-130: ori = res * ori (double, ((R0 x + R1 y) + R2 z) + t*1), pointInMap = ori.cast<float>()
    const double ox = double(po.x), oy = double(po.y), oz = double(po.z);
    float pmf[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[r], ox), __dmul_rn(sT[4 + r], oy)), __dmul_rn(sT[8 + r], oz)), sT[12 + r]);
      pmf[r] = __double2float_rn(v);
    }
    const double q0 = double(pmf[0]), q1 = double(pmf[1]), q2 = double(pmf[2]);
    // ---- exact 5-NN over the 27-cell neighbourhood (cells are exactly 1.0 wide => covers the d2 < 1 ball) ----
    double bd[5];
    int bi[5];
    float bx[5], by[5], bz[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { bd[k] = DBL_MAX; bi[k] = 0x7fffffff; bx[k] = by[k] = bz[k] = 0.f; }
    int status = 0;
    int ncand = 0;
    // cell of the query (same key math as the build); clamp in float first so the int conversion cannot overflow
    const GridSpec& g = grid.g;
    float fc[3];
    fc[0] = __fsub_rn(floorf(__fmul_rn(pmf[0], g.inv_leaf[0])), float(g.min_b[0]));
    fc[1] = __fsub_rn(floorf(__fmul_rn(pmf[1], g.inv_leaf[1])), float(g.min_b[1]));
    fc[2] = __fsub_rn(floorf(__fmul_rn(pmf[2], g.inv_leaf[2])), float(g.min_b[2]));
    bool near = true;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (!(fc[a] >= -1.f && fc[a] <= float(g.div_b[a]))) near = false;
      c[a] = near ? int(fc[a]) : 0;
    }
    if (near) {
      const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, g.div_b[0] - 1);
      for (int dz = -1; dz <= 1; dz++) {
        const int z = c[2] + dz;
        if (z < 0 || z >= g.div_b[2]) continue;
        for (int dy = -1; dy <= 1; dy++) {
          const int y = c[1] + dy;
          if (y < 0 || y >= g.div_b[1]) continue;
          const long long rowbase = (long long)y * g.mul[1] + (long long)z * g.mul[2];
          for (int x = x0; x <= x1; x++) {
            const int2 rg = __ldg(grid.range + rowbase + x);
            ncand += rg.y - rg.x;
            for (int j = rg.x; j < rg.y; j++) {
              const float4 m = __ldg(grid.pts + j);
              const double dx = q0 - double(m.x), dyy = q1 - double(m.y), dzz = q2 - double(m.z);
              const double d2 = dx * dx + dyy * dyy + dzz * dzz;  // exact products; fused or not gives the same bits
              const int idx = __float_as_int(m.w);
              if (knn_less(d2, idx, bd[4], bi[4])) {
                bd[4] = d2; bi[4] = idx; bx[4] = m.x; by[4] = m.y; bz[4] = m.z;
#pragma unroll
                for (int k = 4; k > 0; k--) {
                  if (knn_less(bd[k], bi[k], bd[k - 1], bi[k - 1])) {
                    double td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                    int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                    float tf = bx[k]; bx[k] = bx[k - 1]; bx[k - 1] = tf;
                    tf = by[k]; by[k] = by[k - 1]; by[k - 1] = tf;
                    tf = bz[k]; bz[k] = bz[k - 1]; bz[k - 1] = tf;
                  }
                }
              }
            }
          }
        }
      }
    }
    // LoamRegister.cpp:59 gate: squared distance of the 5th neighbour < 1.0
    const bool gate = (bi[4] != 0x7fffffff) && (bd[4] < prm.max_knn_d2);
    if (DEBUG && dbg_knn) {
#pragma unroll
      for (int k = 0; k < 5; k++) dbg_knn[size_t(i) * 5 + k] = gate ? bi[k] : -1;
    }
    if (gate) {
      status = 1;
      double A[5][3], b[5], x[3];
#pragma unroll
      for (int k = 0; k < 5; k++) { A[k][0] = double(bx[k]); A[k][1] = double(by[k]); A[k][2] = double(bz[k]); b[k] = -1.0; }
      cpqr5x3_solve(A, b, x);
      const double xn = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      bool valid = true;
#pragma unroll
      for (int k = 0; k < 5; k++) {
        double v = x[0] * double(bx[k]) + x[1] * double(by[k]) + x[2] * double(bz[k]);
        if (fabs(v + 1.0) > prm.plane_thresh * xn) valid = false;
      }
      if (valid) {
        status = 2;
        const double dist = ((q0 * x[0] + q1 * x[1] + q2 * x[2]) + 1.0) / xn;
        // :147-148 float range term: sqrt(sqrt(x*x + y*y + z*z)) with float products, sums and roots
        const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(po.x, po.x), __fmul_rn(po.y, po.y)), __fmul_rn(po.z, po.z));
        const float rr = __fsqrt_rn(__fsqrt_rn(r2));
        const double s = 1.0 - 0.9 * fabs(dist) / double(rr);
        if (s > prm.point_thresh) {
          status = 3;
          const double E = s * dist;
          const double sn0 = s * (x[0] / xn), sn1 = s * (x[1] / xn), sn2 = s * (x[2] / xn);
          double J[6];
          J[0] = sn0; J[1] = sn1; J[2] = sn2;
          J[3] = sn1 * (-q2) + sn2 * q1;
          J[4] = sn0 * q2 + sn2 * (-q0);
          J[5] = sn0 * (-q1) + sn1 * q0;
          int k = 0;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = r; cc < 6; cc++) acc[k++] = J[r] * J[cc];
#pragma unroll
          for (int r = 0; r < 6; r++) acc[21 + r] = J[r] * E;
          acc[27] = 1.0;
        }
      }
    }
    acc[28] = double(ncand);
    if (DEBUG && dbg_status) dbg_status[i] = status;
  }

  // ---- block reduction (fixed order) and last-block epilogue ----
  double r = block_reduce_vec<kNV, kLoamBlock>(acc, sred);
  double* my = partials + (size_t(scan) * max_blocks + blockIdx.x) * kNV;
  if (threadIdx.x < kNV) my[threadIdx.x] = r;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&st->ticket, 1u);
    s_last = (t == unsigned(nb - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < kNV) {
    const double* base = partials + size_t(scan) * max_blocks * kNV + threadIdx.x;
    double tsum = 0.0;
    for (int b = 0; b < nb; b++) tsum += __ldcg(base + size_t(b) * kNV);
    stot[threadIdx.x] = tsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ticket = 0;
    double JtJ[36], JtE[6], nx[6], x[6];
    int k = 0;
    for (int rr = 0; rr < 6; rr++)
      for (int cc = rr; cc < 6; cc++) { JtJ[rr * 6 + cc] = stot[k]; JtJ[cc * 6 + rr] = stot[k]; k++; }
    for (int rr = 0; rr < 6; rr++) { JtE[rr] = stot[21 + rr]; nx[rr] = -JtE[rr]; x[rr] = 0.0; }
    const long long n = (long long)(stot[27] + 0.5);
    const int it = st->iters;
    pcr_loam_iter_log* lg = logs ? logs + size_t(scan) * prm.max_iters + it : nullptr;
    if (lg) {
      for (int q = 0; q < 16; q++) lg->T_before[q] = sT[q];
      for (int q = 0; q < 36; q++) lg->JtJ[q] = JtJ[q];
      for (int q = 0; q < 6; q++) lg->JtE[q] = JtE[q];
      lg->n = n; lg->converged = 0; lg->pad = 0;
    }
    st->iters = it + 1;
    st->n_last = int(n);
    st->cand_total += (long long)(stot[28] + 0.5);
    if (n < 6) {  // LoamRegister.cpp:173-176
      st->done = 1;
    } else {
      ldlt6_solve(JtJ, nx, x);
      const double np = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      const double nr = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
      if (np <= prm.pos_conv && nr <= prm.rot_conv) {  // :202-206 — converge BEFORE applying x
        st->converged = 1;
        st->done = 1;
        if (lg) lg->converged = 1;
      } else if (apply_update) {
        double E[16], Tn[16], Tc[16];
        for (int q = 0; q < 16; q++) Tc[q] = sT[q];
        se3_exp(x, E);
        mat4_mul(E, Tc, Tn);
        for (int q = 0; q < 16; q++) st->T[q] = Tn[q];
        if (it + 1 >= prm.max_iters) st->done = 1;
      }
    }
    if (lg) for (int q = 0; q < 6; q++) lg->x[q] = x[q];
    if (!apply_update) st->done = 1;
  }
}

// T2SE3 on every scan's pose (LoamRegister.cpp:220), also for non-converged / aborted scans.
__global__ void loam_finalize_kernel(LoamState* states, int n_scans) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  double T[16];
  for (int q = 0; q < 16; q++) T[q] = states[s].T[q];
  t2se3(T);
  for (int q = 0; q < 16; q++) states[s].T[q] = T[q];
}

LoamDriver::~LoamDriver() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
}

static GridView make_view(const CellGrid& grid) { return view_of(grid); }

int LoamDriver::align(const float4* src, const size_t* offs, size_t n_scans, const CellGrid& grid, const LoamParams& prm,
                      double* T, int32_t* converged, int32_t* iters_out, int64_t* n_last_out, bool profile, cudaStream_t s) {
  launches = 0; cand_total = 0; hot_ms = 0.f; hot_launches = 0;
  if (n_scans == 0) return 0;
  LoamState* hs = h_states.ensure(n_scans);
  uint32_t* ho = h_offsets.ensure(n_scans + 1);
  size_t max_pts = 0;
  for (size_t i = 0; i <= n_scans; i++) ho[i] = uint32_t(offs[i] - offs[0]);
  for (size_t i = 0; i < n_scans; i++) {
    max_pts = std::max(max_pts, size_t(offs[i + 1] - offs[i]));
    memset(&hs[i], 0, sizeof(LoamState));
    for (int q = 0; q < 16; q++) hs[i].T[q] = T[i * 16 + q];
    if (offs[i + 1] == offs[i]) hs[i].done = 1;  // empty scan: n = 0 < 6 -> not converged
  }
  int max_blocks = int((max_pts + kLoamBlock - 1) / kLoamBlock);
  if (max_blocks < 1) max_blocks = 1;
  states.ensure(n_scans);
  offsets.ensure(n_scans + 1);
  partials.ensure(n_scans * size_t(max_blocks) * kNV);
  logs.ensure(n_scans * size_t(prm.max_iters));
  PCR_CUDA_CHECK(cudaMemcpyAsync(states.p, hs, n_scans * sizeof(LoamState), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, (n_scans + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  GridView view = make_view(grid);
  dim3 gridDim(max_blocks, unsigned(n_scans));
  if (profile) {
    if (!ev0) { PCR_CUDA_CHECK(cudaEventCreate(&ev0)); PCR_CUDA_CHECK(cudaEventCreate(&ev1)); }
    PCR_CUDA_CHECK(cudaEventRecord(ev0, s));
  }
  if (grid.built && max_pts > 0) {
    for (int it = 0; it < prm.max_iters; it++) {
      loam_iter_kernel<false><<<gridDim, kLoamBlock, 0, s>>>(src, offsets.p, view, prm, states.p, partials.p, max_blocks, logs.p, 1,
                                                             nullptr, nullptr);
      launches++;
      hot_launches++;
    }
  }
  if (profile) PCR_CUDA_CHECK(cudaEventRecord(ev1, s));
  loam_finalize_kernel<<<unsigned((n_scans + 127) / 128), 128, 0, s>>>(states.p, int(n_scans));
  launches++;
  PCR_CUDA_CHECK(cudaMemcpyAsync(hs, states.p, n_scans * sizeof(LoamState), cudaMemcpyDeviceToHost, s));
  // logs of scan 0 for introspection
  pcr_loam_iter_log* hl = h_logs.ensure(size_t(prm.max_iters));
  PCR_CUDA_CHECK(cudaMemcpyAsync(hl, logs.p, size_t(prm.max_iters) * sizeof(pcr_loam_iter_log), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  PCR_CUDA_CHECK(cudaGetLastError());
  if (profile) PCR_CUDA_CHECK(cudaEventElapsedTime(&hot_ms, ev0, ev1));
  for (size_t i = 0; i < n_scans; i++) {
    for (int q = 0; q < 16; q++) T[i * 16 + q] = hs[i].T[q];
    if (converged) converged[i] = hs[i].converged;
    if (iters_out) iters_out[i] = hs[i].iters;
    if (n_last_out) n_last_out[i] = hs[i].n_last;
    cand_total += hs[i].cand_total;
  }
  last_log_count = hs[0].iters;
  return 0;
}

int LoamDriver::linearize(const float4* src, size_t ns, const CellGrid& grid, const LoamParams& prm, const double* T,
                          int32_t* knn_idx, int32_t* status, double* JtJ, double* JtE, int64_t* n_acc, cudaStream_t s) {
  LoamState* hs = h_states.ensure(1);
  uint32_t* ho = h_offsets.ensure(2);
  ho[0] = 0; ho[1] = uint32_t(ns);
  memset(hs, 0, sizeof(LoamState));
  for (int q = 0; q < 16; q++) hs->T[q] = T[q];
  int max_blocks = int((ns + kLoamBlock - 1) / kLoamBlock);
  if (max_blocks < 1) max_blocks = 1;
  states.ensure(1); offsets.ensure(2);
  partials.ensure(size_t(max_blocks) * kNV);
  logs.ensure(size_t(prm.max_iters));
  dbg_knn.ensure(ns * 5 + 1); dbg_status.ensure(ns + 1);
  PCR_CUDA_CHECK(cudaMemcpyAsync(states.p, hs, sizeof(LoamState), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemsetAsync(logs.p, 0, sizeof(pcr_loam_iter_log), s));
  if (ns > 0 && grid.built) {
    GridView view = make_view(grid);
    loam_iter_kernel<true><<<dim3(max_blocks, 1), kLoamBlock, 0, s>>>(src, offsets.p, view, prm, states.p, partials.p, max_blocks, logs.p,
                                                                     0, dbg_knn.p, dbg_status.p);
  }
  pcr_loam_iter_log* hl = h_logs.ensure(size_t(prm.max_iters));
  PCR_CUDA_CHECK(cudaMemcpyAsync(hl, logs.p, sizeof(pcr_loam_iter_log), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  PCR_CUDA_CHECK(cudaGetLastError());
  if (knn_idx && ns) PCR_CUDA_CHECK(cudaMemcpy(knn_idx, dbg_knn.p, ns * 5 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (status && ns) PCR_CUDA_CHECK(cudaMemcpy(status, dbg_status.p, ns * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (JtJ) memcpy(JtJ, hl->JtJ, sizeof(double) * 36);
  if (JtE) memcpy(JtE, hl->JtE, sizeof(double) * 6);
  if (n_acc) *n_acc = hl->n;
  return 0;
}

}  // namespace pcr
