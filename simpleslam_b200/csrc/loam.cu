// LOAM scan-to-map: one fused kernel per Gauss-Newton iteration. A warp owns a tile of 32 query points:
//   phase 1 (warp-cooperative, one query at a time): exact 5-NN in the 27-cell neighbourhood. Lanes 0..26 fetch the 27
//     cell ranges; the nine x-rows are contiguous runs of the cell-sorted map, flattened into one candidate list that
//     the 32 lanes read coalesced (float4 per lane, FP64 metric). Every lane keeps a sorted private top-5; five rounds
//     of hardware warp-min (redux.sync on the FP64 bit pattern, then on the original index: bit-exact (d2, index)
//     tie-break) merge them. Lane t keeps the five winners of query t.
//   phase 2 (one query per lane): gate, 5x3 column-pivoted QR plane fit, validity / weight gates, residual + SE(3)
//     Jacobian in FP64; the 21 + 6 + 1 normal-equation terms accumulate per thread in shared memory.
//   epilogue: fixed-order block + last-block FP64 reduction, 6x6 LDLT solve, convergence test and exp-map pose update
//     by the last block — the whole Gauss-Newton iteration stays on the device.
// Restates PCR/src/LoamRegister.cpp:99-223 (reference CPU path: nanoflann kd-tree + OpenMP + omp critical).
#include "loam.cuh"
#include "dev_linalg.cuh"
#include <cfloat>
#include <algorithm>

namespace pcr {

constexpr int kLoamBlock = 256;
constexpr int kLoamWarps = kLoamBlock / 32;
constexpr int kNV = 29;  // 21 upper JtJ + 6 JtE + count + candidates examined
constexpr size_t kLoamDynSmem = size_t(kNV) * kLoamBlock * sizeof(double);

using GridView = CellGridView;

struct Cand {  // one entry of a lane's private top-5
  unsigned hi, lo;  // bit pattern of the (non-negative) FP64 squared distance: unsigned order == numeric order
  int idx;          // original map index (tie-break)
  int j;            // position in the cell-sorted array (to re-fetch the coordinates)
};
__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) {
  return a.hi < b.hi || (a.hi == b.hi && (a.lo < b.lo || (a.lo == b.lo && a.idx < b.idx)));
}

template <bool DEBUG>
__global__ void __launch_bounds__(kLoamBlock, 2)
loam_iter_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, GridView grid, LoamParams prm,
                 LoamState* __restrict__ states, double* __restrict__ partials, int max_blocks,
                 pcr_loam_iter_log* __restrict__ logs, int apply_update, int tile, int32_t* __restrict__ dbg_knn,
                 int32_t* __restrict__ dbg_status) {
  // `tile` (power of two <= 32) = queries a warp owns per pass: 32 for throughput on large batches, smaller when there
  // are too few queries to fill the machine (a single scan), trading phase-2 lane utilisation for shorter latency chains.
  const int scan = blockIdx.y;
  const uint32_t begin = offs[scan], end = offs[scan + 1];
  const uint32_t ns = end - begin;
  const uint32_t per_block = uint32_t(kLoamWarps * tile);
  const int nb = min(int((ns + per_block - 1) / per_block), max_blocks);  // blocks working on this scan
  if (int(blockIdx.x) >= nb) return;
  LoamState* st = states + scan;
  if (st->done) return;

  __shared__ double sT[16];
  extern __shared__ double sacc[];  // [kNV][kLoamBlock] per-thread accumulators, column = thread (conflict-free)
  __shared__ double sred[kNV * kLoamWarps];
  __shared__ double stot[kNV];
  __shared__ int s_last;
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
#pragma unroll
  for (int k = 0; k < kNV; k++) sacc[k * kLoamBlock + threadIdx.x] = 0.0;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const GridSpec& g = grid.g;
  double* acc = sacc + threadIdx.x;
  constexpr unsigned FULL = 0xffffffffu;

  for (uint32_t tile0 = begin + (blockIdx.x * kLoamWarps + warp) * uint32_t(tile); tile0 < end; tile0 += uint32_t(nb) * per_block) {
    // ---- my own query (lane-private)
    const uint32_t i = tile0 + lane;
    const bool have = lane < tile && i < end;
    float4 po = make_float4(0.f, 0.f, 0.f, 0.f);
    float pmf[3] = {0.f, 0.f, 0.f};
    int c[3] = {0, 0, 0};
    bool near = false;
    if (have) {
      po = __ldg(src + i);
      // LoamRegister.cpp:128-130: ori = res * ori (double, ((R0 x + R1 y) + R2 z) + t*1), pointInMap = ori.cast<float>()
      const double ox = double(po.x), oy = double(po.y), oz = double(po.z);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[r], ox), __dmul_rn(sT[4 + r], oy)), __dmul_rn(sT[8 + r], oz)), sT[12 + r]);
        pmf[r] = __double2float_rn(v);
      }
      // cell of the query (same key math as the build); range-check in float so the int conversion cannot overflow
      near = true;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const float fc = __fsub_rn(floorf(__fmul_rn(pmf[a], g.inv_leaf[a])), float(g.min_b[a]));
        if (!(fc >= -1.f && fc <= float(g.div_b[a]))) near = false;
        c[a] = near ? int(fc) : 0;
      }
    }
    int wj[5] = {-1, -1, -1, -1, -1};  // cell-sorted positions of my query's five nearest neighbours
    int my_ncand = 0;

    // ---- phase 1: the warp searches the tile's queries one after the other
    const unsigned todo = __ballot_sync(FULL, have && near);
    for (unsigned rem = todo; rem; rem &= rem - 1) {
      const int t = __ffs(rem) - 1;
      const double q0 = double(__shfl_sync(FULL, pmf[0], t)), q1 = double(__shfl_sync(FULL, pmf[1], t)), q2 = double(__shfl_sync(FULL, pmf[2], t));
      const int cx = __shfl_sync(FULL, c[0], t), cy = __shfl_sync(FULL, c[1], t), cz = __shfl_sync(FULL, c[2], t);
      // lanes 0..26 fetch the 27 cell ranges: lane = row*3 + dx, row = (dz+1)*3 + (dy+1)
      int lo = 0x7fffffff, hi = 0;
      if (lane < 27) {
        const int row = lane / 3, x = cx + lane % 3 - 1;
        const int z = cz + row / 3 - 1, y = cy + row % 3 - 1;
        if (x >= 0 && x < g.div_b[0] && y >= 0 && y < g.div_b[1] && z >= 0 && z < g.div_b[2]) {
          const int2 rg = __ldg(grid.range + ((long long)x + (long long)y * g.mul[1] + (long long)z * g.mul[2]));
          if (rg.y > rg.x) { lo = rg.x; hi = rg.y; }
        }
      }
      // the three cells of a row are consecutive keys -> one contiguous run of the cell-sorted array
      lo = min(lo, min(__shfl_down_sync(FULL, lo, 1), __shfl_down_sync(FULL, lo, 2)));
      hi = max(hi, max(__shfl_down_sync(FULL, hi, 1), __shfl_down_sync(FULL, hi, 2)));
      int rlo[9], pre[10];
      pre[0] = 0;
#pragma unroll
      for (int row = 0; row < 9; row++) {
        const int l = __shfl_sync(FULL, lo, row * 3), h = __shfl_sync(FULL, hi, row * 3);
        rlo[row] = l;
        pre[row + 1] = pre[row] + max(h - l, 0);
      }
      const int total = pre[9];
      if (lane == t) my_ncand = total;
      // private sorted top-5
      Cand best[5];
#pragma unroll
      for (int k = 0; k < 5; k++) { best[k].hi = 0xffffffffu; best[k].lo = 0xffffffffu; best[k].idx = 0x7fffffff; best[k].j = -1; }
      for (int f = lane; f < total; f += 32) {
        int j = rlo[0] + f;
#pragma unroll
        for (int row = 1; row < 9; row++) j = (f >= pre[row]) ? rlo[row] + (f - pre[row]) : j;
        const float4 m = __ldg(grid.pts + j);
        const double dx = q0 - double(m.x), dyy = q1 - double(m.y), dzz = q2 - double(m.z);
        const double d2 = dx * dx + dyy * dyy + dzz * dzz;  // exact products; fused or not gives the same bits
        Cand cnd;
        cnd.hi = unsigned(__double2hiint(d2)); cnd.lo = unsigned(__double2loint(d2)); cnd.idx = __float_as_int(m.w); cnd.j = j;
        if (cand_less(cnd, best[4])) {
          best[4] = cnd;
#pragma unroll
          for (int k = 4; k > 0; k--) {
            if (cand_less(best[k], best[k - 1])) { Cand tmp = best[k]; best[k] = best[k - 1]; best[k - 1] = tmp; }
          }
        }
      }
      // merge: five rounds of hardware warp-min over the lanes' heads, (d2 bits, original index) lexicographic
#pragma unroll
      for (int r = 0; r < 5; r++) {
        const unsigned mh = __reduce_min_sync(FULL, best[0].hi);
        const bool e1 = best[0].hi == mh;
        const unsigned ml = __reduce_min_sync(FULL, e1 ? best[0].lo : 0xffffffffu);
        const bool e2 = e1 && best[0].lo == ml;
        const unsigned mi = __reduce_min_sync(FULL, e2 ? unsigned(best[0].idx) : 0x7fffffffu);
        const bool mine = e2 && unsigned(best[0].idx) == mi && best[0].j >= 0;
        const int jw = int(__reduce_min_sync(FULL, mine ? unsigned(best[0].j) : 0xffffffffu));  // -1 when fewer than r+1 exist
        if (lane == t) wj[r] = jw;
        if (mine) {
#pragma unroll
          for (int k = 0; k < 4; k++) best[k] = best[k + 1];
          best[4].hi = 0xffffffffu; best[4].lo = 0xffffffffu; best[4].idx = 0x7fffffff; best[4].j = -1;
        }
      }
    }

    // ---- phase 2: one query per lane
    if (have) {
      int status = 0;
      const double q0 = double(pmf[0]), q1 = double(pmf[1]), q2 = double(pmf[2]);
      float nx[5], ny[5], nz[5];
      int nidx[5];
      bool five = wj[4] >= 0;
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const float4 m = five ? __ldg(grid.pts + wj[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        nx[k] = m.x; ny[k] = m.y; nz[k] = m.z; nidx[k] = __float_as_int(m.w);
      }
      double d5 = DBL_MAX;
      if (five) {
        const double dx = q0 - double(nx[4]), dyy = q1 - double(ny[4]), dzz = q2 - double(nz[4]);
        d5 = dx * dx + dyy * dyy + dzz * dzz;
      }
      // LoamRegister.cpp:59 gate: squared distance of the 5th neighbour < 1.0
      const bool gate = five && (d5 < prm.max_knn_d2);
      if (DEBUG && dbg_knn) {
#pragma unroll
        for (int k = 0; k < 5; k++) dbg_knn[size_t(i) * 5 + k] = gate ? nidx[k] : -1;
      }
      if (gate) {
        status = 1;
        double A[5][3], b[5], x[3];
#pragma unroll
        for (int k = 0; k < 5; k++) { A[k][0] = double(nx[k]); A[k][1] = double(ny[k]); A[k][2] = double(nz[k]); b[k] = -1.0; }
        cpqr5x3_solve(A, b, x);
        const double xn = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        bool valid = true;
#pragma unroll
        for (int k = 0; k < 5; k++) {
          double v = x[0] * double(nx[k]) + x[1] * double(ny[k]) + x[2] * double(nz[k]);
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
              for (int cc = r; cc < 6; cc++) { acc[k * kLoamBlock] += J[r] * J[cc]; k++; }
#pragma unroll
            for (int r = 0; r < 6; r++) acc[(21 + r) * kLoamBlock] += J[r] * E;
            acc[27 * kLoamBlock] += 1.0;
          }
        }
      }
      acc[28 * kLoamBlock] += double(my_ncand);
      if (DEBUG && dbg_status) dbg_status[i] = status;
    }
  }

  // ---- block reduction straight out of shared memory, fixed order: warp w owns components w, w+8, ...
  __syncthreads();
  for (int k = warp; k < kNV; k += kLoamWarps) {
    const double* col = sacc + k * kLoamBlock;
    double v = 0.0;
#pragma unroll
    for (int jj = 0; jj < kLoamBlock / 32; jj++) v += col[lane + 32 * jj];
    v = warp_sum(v);
    if (lane == 0) sred[k] = v;
  }
  __syncthreads();
  double* my = partials + (size_t(scan) * max_blocks + blockIdx.x) * kNV;
  if (threadIdx.x < kNV) {
    my[threadIdx.x] = sred[threadIdx.x];
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&st->ticket, 1u);
    s_last = (t == unsigned(nb - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block: slice-parallel fixed-order sum over the nb block partials
  {
    const int comp = lane, slice = warp;
    double tsum = 0.0;
    if (comp < kNV) {
      const double* base = partials + size_t(scan) * max_blocks * kNV + comp;
      for (int b = slice; b < nb; b += kLoamWarps) tsum += __ldcg(base + size_t(b) * kNV);
    }
    __syncthreads();  // sred is reused
    if (comp < kNV) sred[slice * kNV + comp] = tsum;
    __syncthreads();
    if (threadIdx.x < kNV) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < kLoamWarps; w++) r += sred[w * kNV + threadIdx.x];
      stot[threadIdx.x] = r;
    }
  }
  __syncthreads();
  if (warp != 0) return;
  // ---- warp-parallel epilogue (warp 0 of the last block): normal equations -> LDLT solve -> convergence -> exp-map update
  __shared__ double sJ[36], sF[36], sy[6], sE[16];
  __shared__ int str[6];
  const long long n = (long long)(stot[27] + 0.5);
  const int it = st->iters;
  pcr_loam_iter_log* lg = logs ? logs + size_t(scan) * prm.max_iters + it : nullptr;
  for (int e = lane; e < 36; e += 32) {
    const int r = e / 6, cidx = e % 6, a = min(r, cidx), b = max(r, cidx);
    const double v = stot[a * 6 - a * (a - 1) / 2 + (b - a)];  // index of (a, b), a <= b, in the packed upper triangle
    sJ[e] = v;
    sF[e] = v;
    if (lg) lg->JtJ[e] = v;
  }
  if (lane < 6) {
    sy[lane] = -stot[21 + lane];
    if (lg) { lg->JtE[lane] = stot[21 + lane]; lg->x[lane] = 0.0; }
  }
  if (lg && lane < 16) lg->T_before[lane] = sT[lane];
  if (lg && lane == 0) { lg->n = n; lg->converged = 0; lg->pad = 0; }
  __syncwarp();
  bool done = false, conv = false, update = false;
  if (n < 6) {  // LoamRegister.cpp:173-176
    done = true;
  } else {
    ldlt6_solve_warp(sF, sy, str, lane);  // x -> sy
    const double x0 = sy[0], x1 = sy[1], x2 = sy[2], x3 = sy[3], x4 = sy[4], x5 = sy[5];
    if (lg && lane < 6) lg->x[lane] = sy[lane];
    const double np = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
    const double nr = sqrt(x3 * x3 + x4 * x4 + x5 * x5);
    if (np <= prm.pos_conv && nr <= prm.rot_conv) {  // :202-206 — converge BEFORE applying x
      conv = true;
      done = true;
    } else if (apply_update) {
      update = true;
      if (lane == 0) {
        const double xv[6] = {x0, x1, x2, x3, x4, x5};
        double E[16];
        se3_exp(xv, E);
#pragma unroll
        for (int q = 0; q < 16; q++) sE[q] = E[q];
      }
      __syncwarp();
      if (lane < 16) {  // T <- exp(x) * T   (column-major)
        const int r = lane & 3, cidx = lane >> 2;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) v += sE[q * 4 + r] * sT[cidx * 4 + q];
        st->T[lane] = v;
      }
      if (it + 1 >= prm.max_iters) done = true;
    }
  }
  if (lane == 0) {
    st->ticket = 0;
    st->iters = it + 1;
    st->n_last = int(n);
    st->cand_total += (long long)(stot[28] + 0.5);
    st->pt_evals += (long long)ns;
    if (conv) { st->converged = 1; if (lg) lg->converged = 1; }
    if (done || !apply_update) st->done = 1;
  }
  (void)update;
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

static void loam_opt_in_smem() {
  static bool done_dev[64] = {false};
  int dev = 0;
  PCR_CUDA_CHECK(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return;
  PCR_CUDA_CHECK(cudaFuncSetAttribute(loam_iter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLoamDynSmem)));
  PCR_CUDA_CHECK(cudaFuncSetAttribute(loam_iter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLoamDynSmem)));
  done = true;
}

LoamDriver::~LoamDriver() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
}

static GridView make_view(const CellGrid& grid) { return view_of(grid); }

int LoamDriver::align(const float4* src, const size_t* offs, size_t n_scans, const CellGrid& grid, const LoamParams& prm,
                      double* T, int32_t* converged, int32_t* iters_out, int64_t* n_last_out, bool profile, cudaStream_t s) {
  launches = 0; cand_total = 0; pt_evals = 0; hot_ms = 0.f; hot_launches = 0;
  if (n_scans == 0) return 0;
  loam_opt_in_smem();
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
  // a warp owns 32 queries per pass; at most one resident wave of blocks (2 x 256 threads per SM) shared by the scans
  const size_t total_q = offs[n_scans] - offs[0];
  int tile = 32;
  while (tile > 1 && total_q / size_t(tile) < size_t(kNumSMs) * 16) tile >>= 1;
  const size_t per_block = size_t(kLoamWarps) * tile;
  int max_blocks = int((max_pts + per_block - 1) / per_block);
  max_blocks = std::max(1, std::min(max_blocks, std::max(2, int((kNumSMs * 2 + n_scans - 1) / n_scans))));
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
      loam_iter_kernel<false><<<gridDim, kLoamBlock, kLoamDynSmem, s>>>(src, offsets.p, view, prm, states.p, partials.p, max_blocks, logs.p, 1,
                                                             tile, nullptr, nullptr);
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
    pt_evals += hs[i].pt_evals;
  }
  last_log_count = hs[0].iters;
  return 0;
}

int LoamDriver::linearize(const float4* src, size_t ns, const CellGrid& grid, const LoamParams& prm, const double* T,
                          int32_t* knn_idx, int32_t* status, double* JtJ, double* JtE, int64_t* n_acc, cudaStream_t s) {
  loam_opt_in_smem();
  LoamState* hs = h_states.ensure(1);
  uint32_t* ho = h_offsets.ensure(2);
  ho[0] = 0; ho[1] = uint32_t(ns);
  memset(hs, 0, sizeof(LoamState));
  for (int q = 0; q < 16; q++) hs->T[q] = T[q];
  int tile = 32;
  while (tile > 1 && ns / size_t(tile) < size_t(kNumSMs) * 16) tile >>= 1;
  const size_t per_block = size_t(kLoamWarps) * tile;
  int max_blocks = int((ns + per_block - 1) / per_block);
  max_blocks = std::max(1, std::min(max_blocks, kNumSMs * 2));
  states.ensure(1); offsets.ensure(2);
  partials.ensure(size_t(max_blocks) * kNV);
  logs.ensure(size_t(prm.max_iters));
  dbg_knn.ensure(ns * 5 + 1); dbg_status.ensure(ns + 1);
  PCR_CUDA_CHECK(cudaMemcpyAsync(states.p, hs, sizeof(LoamState), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemsetAsync(logs.p, 0, sizeof(pcr_loam_iter_log), s));
  if (ns > 0 && grid.built) {
    GridView view = make_view(grid);
    loam_iter_kernel<true><<<dim3(max_blocks, 1), kLoamBlock, kLoamDynSmem, s>>>(src, offsets.p, view, prm, states.p, partials.p, max_blocks, logs.p,
                                                                     0, tile, dbg_knn.p, dbg_status.p);
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
