// LOAM scan-to-map. Small problems (a single scan): one FUSED kernel per Gauss-Newton iteration, described here; large
// batches: a search kernel + a fit kernel per iteration (see "Batch path" below). Fused kernel: a warp owns a tile of up to
// 32 query points.
//   phase 1: exact 5-NN on a uniform grid whose cells are HALF the gate radius wide (gate: 5th neighbour closer than 1 m,
//     LoamRegister.cpp:59). LPQ lanes (1, 2, 4 or 8, picked from the problem size) co-operate on one query: the 3x3 x-rows
//     of the 27-cell ring are contiguous runs of the cell-sorted map (two loads of the dense `start` table per row), the
//     rows are dealt out to the lanes, every lane keeps a sorted private top-5 (FP64 metric on FP32 coordinates, bit
//     pattern compare), five rounds of hardware redux.sync min over (d2 bits, original index) merge them: bit-exact
//     (d2, index) tie-break. If the 5th distance is not provably inside the ring (d < (1 + margin) cells) the 5x5x5 ring is
//     searched instead; it covers the whole gate radius, so every ACCEPTED query has its exact 5-NN.
//   phase 2 (one query per lane): gate, 5x3 column-pivoted QR plane fit, validity / weight gates, residual + SE(3)
//     Jacobian in FP64; the 21 + 6 + 1 normal-equation terms accumulate per thread in shared memory.
//   epilogue: fixed-order block + last-block FP64 reduction, 6x6 LDLT solve, convergence test and exp-map pose update
//     by the last block — the whole Gauss-Newton iteration stays on the device.
// Restates PCR/src/LoamRegister.cpp:99-223 (reference CPU path: nanoflann kd-tree + OpenMP + omp critical).
#include "loam.cuh"
#include "dev_linalg.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <algorithm>
#include <cstdlib>

namespace pcr {

constexpr int kLoamBlock = 256;
constexpr int kLoamWarps = kLoamBlock / 32;
constexpr int kNV = 30;  // 21 upper JtJ + 6 JtE + count + candidates examined + x-rows looked up
constexpr size_t kLoamDynSmem = size_t(kNV) * kLoamBlock * sizeof(double);

using GridView = CellGridView;

struct Cand {  // one entry of a lane's private top-5
  unsigned long long key;  // bit pattern of the (non-negative) FP64 squared distance: unsigned order == numeric order
  int idx;                 // original map index (tie-break)
  int j;                   // position in the cell-sorted array (to re-fetch the coordinates)
};
__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) { return a.key < b.key || (a.key == b.key && a.idx < b.idx); }
__device__ __forceinline__ Cand cand_sel(bool p, const Cand& a, const Cand& b) {
  Cand r;
  r.key = p ? a.key : b.key; r.idx = p ? a.idx : b.idx; r.j = p ? a.j : b.j;
  return r;
}

// (dy, dz) of the x-rows of the 5x5 neighbourhood, nearest first: centre, 4 faces, 4 corners (ring 1), then the 16 rows of
// ring 2. With R rings the first (2R+1)^2 entries are used.
__constant__ signed char c_row_dy[25] = {0, 1, -1, 0, 0, 1, 1, -1, -1, 2, -2, 0, 0, 2, 2, -2, -2, 1, -1, 1, -1, 2, 2, -2, -2};
__constant__ signed char c_row_dz[25] = {0, 0, 0, 1, -1, 1, -1, 1, -1, 0, 0, 2, -2, 1, -1, 1, -1, 2, 2, -2, -2, 2, -2, 2, -2};

// Fused iteration (search + fit + solve) — small problems, one launch per iteration keeps the latency short.
// A5-A7 for one query whose five neighbours are known (ascending (d2, index) order): gate, 5x3 column-pivoted QR plane fit,
// validity / weight gates, residual E and SE(3) Jacobian row J (LoamRegister.cpp:29-72, 141-160), all FP64 with the
// reference's float range term. Returns the status (0 gate, 1 plane invalid, 2 weight, 3 accepted); J, E valid for 3.
__device__ __forceinline__ int loam_point_residual(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5], bool five, double d5,
                                                   double q0, double q1, double q2, const float4& po, const LoamParams& prm, double (&J)[6],
                                                   double& E) {
  // LoamRegister.cpp:59 gate: squared distance of the 5th neighbour < 1.0
  if (!(five && d5 < prm.max_knn_d2)) return 0;
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
  if (!valid) return 1;
  const double dist = ((q0 * x[0] + q1 * x[1] + q2 * x[2]) + 1.0) / xn;
  // :147-148 float range term: sqrt(sqrt(x*x + y*y + z*z)) with float products, sums and roots
  const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(po.x, po.x), __fmul_rn(po.y, po.y)), __fmul_rn(po.z, po.z));
  const float rr = __fsqrt_rn(__fsqrt_rn(r2));
  const double s = 1.0 - 0.9 * fabs(dist) / double(rr);
  if (!(s > prm.point_thresh)) return 2;
  E = s * dist;
  const double sn0 = s * (x[0] / xn), sn1 = s * (x[1] / xn), sn2 = s * (x[2] / xn);
  J[0] = sn0; J[1] = sn1; J[2] = sn2;
  J[3] = sn1 * (-q2) + sn2 * q1;
  J[4] = sn0 * q2 + sn2 * (-q0);
  J[5] = sn0 * (-q1) + sn1 * q0;
  return 3;
}

// A8 + A9 by one warp once the 30 sums of a scan are complete (stot): normal equations -> warp-parallel 6x6 LDLT ->
// convergence test -> T <- exp(x) T, iteration log (LoamRegister.cpp:171-220). sT = the pose this iteration linearised at.
struct SolveScratch { double sJ[36], sF[36], sy[6], sE[16]; int str[6]; };
__device__ __forceinline__ void loam_solve_step(const double* stot, const double* sT, LoamState* st, pcr_loam_iter_log* logs, int scan,
                                                const LoamParams& prm, int apply_update, uint32_t ns, int lane, SolveScratch& sc) {
  const long long n = (long long)(stot[27] + 0.5);
  const int it = st->iters;
  pcr_loam_iter_log* lg = logs ? logs + size_t(scan) * prm.max_iters + it : nullptr;
  for (int e = lane; e < 36; e += 32) {
    const int r = e / 6, cidx = e % 6, a = min(r, cidx), b = max(r, cidx);
    const double v = stot[a * 6 - a * (a - 1) / 2 + (b - a)];  // index of (a, b), a <= b, in the packed upper triangle
    sc.sJ[e] = v;
    sc.sF[e] = v;
    if (lg) lg->JtJ[e] = v;
  }
  if (lane < 6) {
    sc.sy[lane] = -stot[21 + lane];
    if (lg) { lg->JtE[lane] = stot[21 + lane]; lg->x[lane] = 0.0; }
  }
  if (lg && lane < 16) lg->T_before[lane] = sT[lane];
  if (lg && lane == 0) { lg->n = n; lg->converged = 0; lg->pad = 0; }
  __syncwarp();
  bool done = false, conv = false;
  if (n < 6) {  // LoamRegister.cpp:173-176
    done = true;
  } else {
    ldlt6_solve_warp(sc.sF, sc.sy, sc.str, lane);  // x -> sy
    const double x0 = sc.sy[0], x1 = sc.sy[1], x2 = sc.sy[2], x3 = sc.sy[3], x4 = sc.sy[4], x5 = sc.sy[5];
    if (lg && lane < 6) lg->x[lane] = sc.sy[lane];
    const double np = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
    const double nr = sqrt(x3 * x3 + x4 * x4 + x5 * x5);
    if (np <= prm.pos_conv && nr <= prm.rot_conv) {  // :202-206 — converge BEFORE applying x
      conv = true;
      done = true;
    } else if (apply_update) {
      if (lane == 0) {
        const double xv[6] = {x0, x1, x2, x3, x4, x5};
        double E[16];
        se3_exp(xv, E);
#pragma unroll
        for (int q = 0; q < 16; q++) sc.sE[q] = E[q];
      }
      __syncwarp();
      if (lane < 16) {  // T <- exp(x) * T   (column-major)
        const int r = lane & 3, cidx = lane >> 2;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) v += sc.sE[q * 4 + r] * sT[cidx * 4 + q];
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
    st->rows_total += (long long)(stot[29] + 0.5);
    st->pt_evals += (long long)ns;
    if (conv) { st->converged = 1; if (lg) lg->converged = 1; }
    if (done || !apply_update) st->done = 1;
  }
}

template <int LPQ, bool DEBUG>
__global__ void __launch_bounds__(kLoamBlock, 2)
loam_iter_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, GridView grid, LoamParams prm,
                 LoamState* __restrict__ states, double* __restrict__ partials, int max_blocks,
                 pcr_loam_iter_log* __restrict__ logs, int apply_update, int tile, double slack, int max_ring,
                 int32_t* __restrict__ dbg_knn, int32_t* __restrict__ dbg_status) {
  // LPQ = lanes co-operating on one query, G = 32 / LPQ queries searched concurrently by a warp.
  // `tile` (multiple of G, <= 32) = queries a warp owns per pass: 32 for throughput on large batches, smaller when there
  // are too few queries to fill the machine (a single scan), trading phase-2 lane utilisation for shorter latency chains.
  constexpr int G = 32 / LPQ;
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
  const int gl = lane % LPQ, gi = lane / LPQ;  // lane within its group, group within the warp
  const GridSpec& g = grid.g;
  double* acc = sacc + threadIdx.x;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned gmask = LPQ == 32 ? FULL : (((1u << LPQ) - 1u) << (gi * LPQ));
  const double leaf = double(g.leaf[0]);

  for (uint32_t tile0 = begin + (blockIdx.x * kLoamWarps + warp) * uint32_t(tile); tile0 < end; tile0 += uint32_t(nb) * per_block) {
    // ---- my own query (lane-private)
    const uint32_t i = tile0 + lane;
    const bool have = lane < tile && i < end;
    float4 po = make_float4(0.f, 0.f, 0.f, 0.f);
    float pmf[3] = {0.f, 0.f, 0.f};
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
        if (!(fc >= float(-max_ring) && fc <= float(g.div_b[a] - 1 + max_ring))) near = false;  // the rings would not touch the grid
      }
    }
    int wj[5] = {-1, -1, -1, -1, -1};  // cell-sorted positions of my query's five nearest neighbours
    int my_ncand = 0, my_nrows = 0;

    // ---- phase 1: G queries of the tile are searched per round, LPQ lanes each
    const unsigned todo = __ballot_sync(FULL, have && near);
    const int rounds = (tile + G - 1) / G;
    for (int r = 0; r < rounds; r++) {
      const unsigned round_bits = (G == 32 ? todo : ((todo >> (r * G)) & ((1u << G) - 1u)));
      if (round_bits == 0) continue;  // warp-uniform
      const int t = r * G + gi;        // the query (lane of the tile) my group works on
      const float qf0 = __shfl_sync(FULL, pmf[0], t), qf1 = __shfl_sync(FULL, pmf[1], t), qf2 = __shfl_sync(FULL, pmf[2], t);
      const double q0 = double(qf0), q1 = double(qf1), q2 = double(qf2);
      const bool act = (todo >> t) & 1u;
      int jw[5] = {-1, -1, -1, -1, -1};
      int ncand = 0, nrows = 0;
      if (act) {  // group-uniform
        // cell and in-cell position of the query (cells; same float key math as the build)
        const float sx = __fmul_rn(qf0, g.inv_leaf[0]), sy = __fmul_rn(qf1, g.inv_leaf[1]), sz = __fmul_rn(qf2, g.inv_leaf[2]);
        const float flx = floorf(sx), fly = floorf(sy), flz = floorf(sz);
        const int cx = int(__fsub_rn(flx, float(g.min_b[0]))), cy = int(__fsub_rn(fly, float(g.min_b[1]))), cz = int(__fsub_rn(flz, float(g.min_b[2])));
        const float fx = sx - flx, fy = sy - fly, fz = sz - flz;
        const float h2 = float(leaf * leaf) * 0.99999f;
        const float slk = float(slack);
        // private sorted top-5
        Cand best[5];
#pragma unroll
        for (int k = 0; k < 5; k++) { best[k].key = ~0ull; best[k].idx = 0x7fffffff; best[k].j = -1; }
        // FP32 pre-filter. thr is a float upper bound (1e-5 relative head-room; the float evaluation is good to 3e-7) of
        // min(gate, current 5th-best exact distance): a candidate that the exact FP64 (d2, index) compare would accept AND
        // that can matter for an accepted query (5th neighbour closer than the gate, LoamRegister.cpp:59) is never dropped.
        // Rows and x-cells are pruned against the same bound, so only the cells that can still hold a top-5 point are read.
        float thr = float(prm.max_knn_d2) * 1.00001f;
        auto exact = [&](float mx, float my, float mz, int midx, int j) {
          const double dx = q0 - double(mx), dyy = q1 - double(my), dzz = q2 - double(mz);
          const double d2 = dx * dx + dyy * dyy + dzz * dzz;  // exact products; fused or not gives the same bits
          Cand cnd;
          cnd.key = (unsigned long long)__double_as_longlong(d2); cnd.idx = midx; cnd.j = j;
          if (cand_less(cnd, best[4])) {  // branch-free sorted insertion
            const bool l0 = cand_less(best[0], cnd), l1 = cand_less(best[1], cnd), l2 = cand_less(best[2], cnd), l3 = cand_less(best[3], cnd);
            best[4] = cand_sel(l3, cnd, best[3]);
            best[3] = cand_sel(l3, best[3], cand_sel(l2, cnd, best[2]));
            best[2] = cand_sel(l2, best[2], cand_sel(l1, cnd, best[1]));
            best[1] = cand_sel(l1, best[1], cand_sel(l0, cnd, best[0]));
            best[0] = cand_sel(l0, best[0], cnd);
            if (best[4].j >= 0) thr = fminf(thr, __double2float_ru(__longlong_as_double((long long)best[4].key)) * 1.00001f);
          }
        };
        auto d2f_of = [&](const float4& m) {
          const float ax = qf0 - m.x, ay = qf1 - m.y, az = qf2 - m.z;
          return fmaf(az, az, fmaf(ay, ay, ax * ax));
        };
        // distance (cells) from the query to the cells at x-offset +-2: below it one x-cell either side is enough
        const float ax2 = fmaxf(1.f + fminf(fx, 1.f - fx) - slk, 0.f);
        const int NR = (2 * max_ring + 1) * (2 * max_ring + 1);
        // every row of ring 2 is at least this far (squared): once the bound is below it the second ring is skipped as a whole
        const float ring2 = fmaxf(1.f + fminf(fminf(fy, 1.f - fy), fminf(fz, 1.f - fz)) - slk, 0.f);
        const float ring2_min2 = ring2 * ring2 * h2;
#pragma unroll 1
        for (int k = gl; k < NR; k += LPQ) {
          if (k >= 9 && thr < ring2_min2) break;
          const int dy = c_row_dy[k], dz = c_row_dz[k];
          // squared distance from the query to the row's (y, z) slab, shrunk by the cell-assignment slack
          const float ay = fmaxf((dy == 0 ? 0.f : (dy > 0 ? float(dy) - fy : fy + float(-dy - 1))) - slk, 0.f);
          const float az = fmaxf((dz == 0 ? 0.f : (dz > 0 ? float(dz) - fz : fz + float(-dz - 1))) - slk, 0.f);
          const float row2 = (ay * ay + az * az) * h2;
          if (row2 > thr) continue;  // every point of this row is farther than the current bound
          const int y = cy + dy, z = cz + dz;
          if (y < 0 || y >= g.div_b[1] || z < 0 || z >= g.div_b[2]) continue;
          const int rx = (max_ring > 1 && thr < row2 + ax2 * ax2 * h2) ? 1 : max_ring;
          const int x0 = max(cx - rx, 0), x1 = min(cx + rx, g.div_b[0] - 1);
          if (x0 > x1) continue;
          // a row is ONE contiguous run of the cell-sorted map: two loads of the dense start table
          const long long key0 = (long long)x0 + (long long)y * g.mul[1] + (long long)z * g.mul[2];
          const int lo = __ldg(grid.start + key0);
          const int hi = __ldg(grid.start + key0 + (x1 - x0) + 1);
          ncand += hi - lo;
          nrows++;
#pragma unroll 1
          for (int j = lo; j < hi; j += 4) {  // four independent float4 loads in flight
            const int rem = hi - j;
            const float4 none = make_float4(0.f, 0.f, 0.f, 0.f);  // masked out below
            const float4 m0 = __ldg(grid.pts + j);
            const float4 m1 = rem > 1 ? __ldg(grid.pts + j + 1) : none;
            const float4 m2 = rem > 2 ? __ldg(grid.pts + j + 2) : none;
            const float4 m3 = rem > 3 ? __ldg(grid.pts + j + 3) : none;
            const float f0 = d2f_of(m0), f1 = d2f_of(m1), f2 = d2f_of(m2), f3 = d2f_of(m3);
            unsigned pass = (f0 <= thr ? 1u : 0u) | (rem > 1 && f1 <= thr ? 2u : 0u) | (rem > 2 && f2 <= thr ? 4u : 0u) |
                            (rem > 3 && f3 <= thr ? 8u : 0u);
            while (pass) {  // rare once five candidates are in
              const int u = __ffs(pass) - 1;
              pass &= pass - 1;
              const float4 m = u == 0 ? m0 : (u == 1 ? m1 : (u == 2 ? m2 : m3));
              exact(m.x, m.y, m.z, __float_as_int(m.w), j + u);
            }
          }
        }
        // merge: five rounds of hardware min over the group's heads, (d2 bits, original index) lexicographic
        if (LPQ == 1) {
#pragma unroll
          for (int q = 0; q < 5; q++) jw[q] = best[q].j;
        } else {
#pragma unroll
          for (int q = 0; q < 5; q++) {
            const unsigned bh = unsigned(best[0].key >> 32), bl = unsigned(best[0].key);
            const unsigned mh = __reduce_min_sync(gmask, bh);
            const bool e1 = bh == mh;
            const unsigned ml = __reduce_min_sync(gmask, e1 ? bl : 0xffffffffu);
            const bool e2 = e1 && bl == ml;
            const unsigned mi = __reduce_min_sync(gmask, e2 ? unsigned(best[0].idx) : 0x7fffffffu);
            const bool mine = e2 && unsigned(best[0].idx) == mi && best[0].j >= 0;
            jw[q] = int(__reduce_min_sync(gmask, mine ? unsigned(best[0].j) : 0xffffffffu));  // -1 when fewer than q+1 exist
            if (mine) {
#pragma unroll
              for (int k = 0; k < 4; k++) best[k] = best[k + 1];
              best[4].key = ~0ull; best[4].idx = 0x7fffffff; best[4].j = -1;
            }
          }
          ncand = int(__reduce_add_sync(gmask, unsigned(ncand)));  // candidates examined / rows looked up by the group
          nrows = int(__reduce_add_sync(gmask, unsigned(nrows)));
        }
      }
      // hand the winners to the lane that owns the query
#pragma unroll
      for (int q = 0; q < 5; q++) {
        const int v = __shfl_sync(FULL, jw[q], (lane % G) * LPQ);
        if (lane / G == r) wj[q] = v;
      }
      const int nc = __shfl_sync(FULL, ncand, (lane % G) * LPQ), nr = __shfl_sync(FULL, nrows, (lane % G) * LPQ);
      if (lane / G == r) { my_ncand = nc; my_nrows = nr; }
    }

    // ---- phase 2: one query per lane
    if (have) {
      const double q0 = double(pmf[0]), q1 = double(pmf[1]), q2 = double(pmf[2]);
      float nx[5], ny[5], nz[5];
      int nidx[5];
      const bool five = wj[4] >= 0;
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
      double J[6], E = 0.0;
      const int status = loam_point_residual(nx, ny, nz, five, d5, q0, q1, q2, po, prm, J, E);
      if (DEBUG && dbg_knn) {
#pragma unroll
        for (int k = 0; k < 5; k++) dbg_knn[size_t(i) * 5 + k] = status >= 1 ? nidx[k] : -1;
      }
      if (status == 3) {
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int cc = r; cc < 6; cc++) { acc[k * kLoamBlock] += J[r] * J[cc]; k++; }
#pragma unroll
        for (int r = 0; r < 6; r++) acc[(21 + r) * kLoamBlock] += J[r] * E;
        acc[27 * kLoamBlock] += 1.0;
      }
      acc[28 * kLoamBlock] += double(my_ncand);
      acc[29 * kLoamBlock] += double(my_nrows);
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
  __shared__ SolveScratch sc;
  loam_solve_step(stot, sT, st, logs, scan, prm, apply_update, ns, lane, sc);
}

// ================================================================================================================
// Batch path (>= ~150 k queries per iteration): every Gauss-Newton iteration is two kernels.
//
// loam_search_kernel — one lane per query, queries in Morton order of their scan-frame position (neighbouring lanes walk
// neighbouring rows of the map: similar trip counts, shared cache lines). Rows are visited nearest first and pruned against
// the current bound, candidates are read four float4 at a time. (Measured and rejected, profiles/README.md: issuing all
// rows' table entries up front and walking a per-lane list of non-empty rows from shared memory — fewer instructions,
// better lane utilisation, but 1.5x slower: the extra table sectors of rows the bound would have pruned cost more than
// the saved latency; warm-starting the bound from the previous iteration's winners — no change; a select-only top-5 update
// — 27 % more instructions, lanes per instruction 14 -> 15, 6 % slower; one flattened (row switch | four candidates) loop
// instead of the nested row / chunk loops — 24..47 % more instructions, 2..12 % slower. STAGE (PCR_LOAM_STAGE=1) keeps the
// cp.async double buffering of a lane's candidates selectable: parity-green, 13 % more instructions, 1..6 % slower.)
// Selection runs on the FP32 metric f (|f - e| <= 3e-7 e against the exact FP64 metric e of the reference, both taken on
// the same float coordinates): a lane keeps its five smallest (f, position) pairs and the smallest f that was looked at but
// is NOT among them (f_out). Candidates and rows are pruned against thr = 1.00001 * min(gate, current 5th f), so whatever
// is dropped unseen is farther than the final 5th by a margin of 1e-5 >> 2 * 3e-7. If f_out > 5th f * (1 + 9.5e-7) the five
// kept candidates are exactly the reference's five nearest (as a set; the fit kernel orders them by the exact (e, index)
// key). Otherwise — a near tie between the 5th and the 6th, e.g. on quantised clouds — the query is searched again by
// exact_search_one with the FP64 (e, index) order throughout.
// The five winners leave the kernel as COORDINATES (five float4 planes, w = original map index): they are cache-hot here,
// and the fit kernel then reads them coalesced instead of chasing five pointers per query.
//
// loam_fit_kernel — one query per thread, FP64: exact ordering of the five, gate, pivoted QR plane fit, residual, Jacobian.
// No block barrier and no per-thread accumulators: a warp reduces its 32 contributions through a transposed shared-memory
// tile (lane p owns one of the 28 sums), writes one partial per warp, and the warp that takes the last ticket of the scan
// sums the partials in fixed order and runs the solve (bit-reproducible run to run).
// ================================================================================================================
struct RowGeom {  // per-query constants of the row walk (cells; same float key math as the build)
  int cx, cy, cz;
  float fx, fy, fz, h2, slk, ax2, ring2_min2;
};
__device__ __forceinline__ RowGeom row_geom(const GridSpec& g, float qf0, float qf1, float qf2, double leaf, double slack) {
  RowGeom r;
  const float sx = __fmul_rn(qf0, g.inv_leaf[0]), sy = __fmul_rn(qf1, g.inv_leaf[1]), sz = __fmul_rn(qf2, g.inv_leaf[2]);
  const float flx = floorf(sx), fly = floorf(sy), flz = floorf(sz);
  r.cx = int(__fsub_rn(flx, float(g.min_b[0]))); r.cy = int(__fsub_rn(fly, float(g.min_b[1]))); r.cz = int(__fsub_rn(flz, float(g.min_b[2])));
  r.fx = sx - flx; r.fy = sy - fly; r.fz = sz - flz;
  r.h2 = float(leaf * leaf) * 0.99999f;
  r.slk = float(slack);
  // distance (cells) from the query to the cells at x-offset +-2: below it one x-cell either side is enough
  r.ax2 = fmaxf(1.f + fminf(r.fx, 1.f - r.fx) - r.slk, 0.f);
  // every row of ring 2 is at least this far (squared): once the bound is below it the second ring is skipped as a whole
  const float ring2 = fmaxf(1.f + fminf(fminf(r.fy, 1.f - r.fy), fminf(r.fz, 1.f - r.fz)) - r.slk, 0.f);
  r.ring2_min2 = ring2 * ring2 * r.h2;
  return r;
}
// x-run of row k of the neighbourhood that can still hold a point closer than thr: [lo, hi) in the cell-sorted map
__device__ __forceinline__ bool row_run(const GridView& grid, const RowGeom& q, int k, float thr, int max_ring, int& lo, int& hi) {
  const GridSpec& g = grid.g;
  const int dy = c_row_dy[k], dz = c_row_dz[k];
  // squared distance from the query to the row's (y, z) slab, shrunk by the cell-assignment slack
  const float ay = fmaxf((dy == 0 ? 0.f : (dy > 0 ? float(dy) - q.fy : q.fy + float(-dy - 1))) - q.slk, 0.f);
  const float az = fmaxf((dz == 0 ? 0.f : (dz > 0 ? float(dz) - q.fz : q.fz + float(-dz - 1))) - q.slk, 0.f);
  const float row2 = (ay * ay + az * az) * q.h2;
  if (row2 > thr) return false;  // every point of this row is farther than the current bound
  const int y = q.cy + dy, z = q.cz + dz;
  if (y < 0 || y >= g.div_b[1] || z < 0 || z >= g.div_b[2]) return false;
  const int rx = (max_ring > 1 && thr < row2 + q.ax2 * q.ax2 * q.h2) ? 1 : max_ring;
  const int x0 = max(q.cx - rx, 0), x1 = min(q.cx + rx, g.div_b[0] - 1);
  if (x0 > x1) return false;
  // a row is ONE contiguous run of the cell-sorted map: two loads of the dense start table
  const long long key0 = (long long)x0 + (long long)y * g.mul[1] + (long long)z * g.mul[2];
  lo = __ldg(grid.start + key0);
  hi = __ldg(grid.start + key0 + (x1 - x0) + 1);
  return true;
}

// exact path for one query (rare): FP64 (d2, index) order throughout, as the fused kernel does
__device__ __noinline__ void exact_search_one(const GridView& grid, float qf0, float qf1, float qf2, double leaf, double slack, int max_ring,
                                              float gate_thr, int (&wj)[5], int& ncand, int& nrows) {
  const RowGeom q = row_geom(grid.g, qf0, qf1, qf2, leaf, slack);
  const double q0 = double(qf0), q1 = double(qf1), q2 = double(qf2);
  Cand best[5];
#pragma unroll
  for (int k = 0; k < 5; k++) { best[k].key = ~0ull; best[k].idx = 0x7fffffff; best[k].j = -1; }
  float thr = gate_thr;
  const int NR = (2 * max_ring + 1) * (2 * max_ring + 1);
#pragma unroll 1
  for (int k = 0; k < NR; k++) {
    if (k >= 9 && thr < q.ring2_min2) break;
    int lo, hi;
    if (!row_run(grid, q, k, thr, max_ring, lo, hi)) continue;
    ncand += hi - lo;
    nrows++;
#pragma unroll 1
    for (int j = lo; j < hi; j++) {
      const float4 m = __ldg(grid.pts + j);
      const float ax = qf0 - m.x, ay = qf1 - m.y, az = qf2 - m.z;
      if (!(fmaf(az, az, fmaf(ay, ay, ax * ax)) <= thr)) continue;
      const double dx = q0 - double(m.x), dyy = q1 - double(m.y), dzz = q2 - double(m.z);
      Cand cnd;
      cnd.key = (unsigned long long)__double_as_longlong(dx * dx + dyy * dyy + dzz * dzz); cnd.idx = __float_as_int(m.w); cnd.j = j;
      if (cand_less(cnd, best[4])) {
        const bool l0 = cand_less(best[0], cnd), l1 = cand_less(best[1], cnd), l2 = cand_less(best[2], cnd), l3 = cand_less(best[3], cnd);
        best[4] = cand_sel(l3, cnd, best[3]);
        best[3] = cand_sel(l3, best[3], cand_sel(l2, cnd, best[2]));
        best[2] = cand_sel(l2, best[2], cand_sel(l1, cnd, best[1]));
        best[1] = cand_sel(l1, best[1], cand_sel(l0, cnd, best[0]));
        best[0] = cand_sel(l0, best[0], cnd);
        if (best[4].j >= 0) thr = fminf(thr, __double2float_ru(__longlong_as_double((long long)best[4].key)) * 1.00001f);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 5; k++) wj[k] = best[k].j;
}

// cp.async (LDGSTS) staging of a lane's candidates in a private shared-memory slot: the next four candidates of a row are
// copied while the current four are examined, at no register cost (the register version of the same double buffering
// needs 16 more registers and loses a resident block).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(unsigned(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool PREFETCH, bool STAGE>
__global__ void __launch_bounds__(kLoamBlock, 4)
loam_search_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, GridView grid, LoamParams prm,
                   const LoamState* __restrict__ states, double slack, int max_ring, float4* __restrict__ nb_out, int2* __restrict__ cnt_out,
                   size_t stride) {
  const int scan = blockIdx.y;
  const uint32_t begin = offs[scan], end = offs[scan + 1];
  const LoamState* st = states + scan;
  if (st->done) return;
  __shared__ double sT[16];
  __shared__ float4 s_stage[STAGE ? 2 * 4 * kLoamBlock : 1];  // [buffer][slot][thread]: conflict-free 16-byte accesses
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
  __syncthreads();
  const int tid = threadIdx.x;
  const GridSpec& g = grid.g;
  const double leaf = double(g.leaf[0]);
  const float gate_thr = float(prm.max_knn_d2) * 1.00001f;
  for (uint32_t i = begin + blockIdx.x * kLoamBlock + tid; i < end; i += gridDim.x * kLoamBlock) {
    const float4 po = __ldg(src + i);
    // LoamRegister.cpp:128-130: ori = res * ori (double, ((R0 x + R1 y) + R2 z) + t*1), pointInMap = ori.cast<float>()
    const double ox = double(po.x), oy = double(po.y), oz = double(po.z);
    float pmf[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[r], ox), __dmul_rn(sT[4 + r], oy)), __dmul_rn(sT[8 + r], oz)), sT[12 + r]);
      pmf[r] = __double2float_rn(v);
    }
    bool near = true;  // range-check in float so the int conversion cannot overflow
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float fc = __fsub_rn(floorf(__fmul_rn(pmf[a], g.inv_leaf[a])), float(g.min_b[a]));
      if (!(fc >= float(-max_ring) && fc <= float(g.div_b[a] - 1 + max_ring))) near = false;  // the rings would not touch the grid
    }
    int bj[5] = {-1, -1, -1, -1, -1};
    int ncand = 0, nrows = 0;
    if (near) {
      const float qf0 = pmf[0], qf1 = pmf[1], qf2 = pmf[2];
      const RowGeom q = row_geom(g, qf0, qf1, qf2, leaf, slack);
      float bf[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};  // five smallest f, ascending
      float f_out = INFINITY;  // smallest f that was examined but is not among the five
      float thr = gate_thr;
      auto consider = [&](const float4& m, int j) {
        const float ax = qf0 - m.x, ay = qf1 - m.y, az = qf2 - m.z;
        const float f = fmaf(az, az, fmaf(ay, ay, ax * ax));
        if (f <= thr) {
          if (f < bf[4]) {  // branch-free sorted insertion; the old 5th drops out
            f_out = fminf(f_out, bf[4]);
            const bool p3 = f < bf[3], p2 = f < bf[2], p1 = f < bf[1], p0 = f < bf[0];
            bf[4] = p3 ? bf[3] : f;                   bj[4] = p3 ? bj[3] : j;
            bf[3] = p3 ? (p2 ? bf[2] : f) : bf[3];    bj[3] = p3 ? (p2 ? bj[2] : j) : bj[3];
            bf[2] = p2 ? (p1 ? bf[1] : f) : bf[2];    bj[2] = p2 ? (p1 ? bj[1] : j) : bj[2];
            bf[1] = p1 ? (p0 ? bf[0] : f) : bf[1];    bj[1] = p1 ? (p0 ? bj[0] : j) : bj[1];
            bf[0] = p0 ? f : bf[0];                   bj[0] = p0 ? j : bj[0];
            thr = fminf(thr, bf[4] * 1.00001f);       // stays at the gate bound until five candidates are in (bf[4] = inf)
          } else {
            f_out = fminf(f_out, f);
          }
        }
      };
      const int NR = (2 * max_ring + 1) * (2 * max_ring + 1);
      // rows nearest first; each row is one contiguous run of the cell-sorted map. The table entries of the NEXT row are
      // requested before the candidates of the current one are scanned (two register sets, A and B, taking turns), so
      // that one of the two dependent loads per row overlaps with work. A row chosen with the looser, earlier bound is
      // tested against the current bound again before it is scanned.
      int k = 0;
      auto next_row = [&](int& lo, int& hi, float& row2) -> bool {
        for (; k < NR; k++) {
          if (k >= 9 && thr < q.ring2_min2) { k = NR; return false; }
          const int dy = c_row_dy[k], dz = c_row_dz[k];
          const float ay = fmaxf((dy == 0 ? 0.f : (dy > 0 ? float(dy) - q.fy : q.fy + float(-dy - 1))) - q.slk, 0.f);
          const float az = fmaxf((dz == 0 ? 0.f : (dz > 0 ? float(dz) - q.fz : q.fz + float(-dz - 1))) - q.slk, 0.f);
          row2 = (ay * ay + az * az) * q.h2;
          if (row2 > thr) continue;  // every point of this row is farther than the current bound
          const int y = q.cy + dy, z = q.cz + dz;
          if (y < 0 || y >= g.div_b[1] || z < 0 || z >= g.div_b[2]) continue;
          const int rx = (max_ring > 1 && thr < row2 + q.ax2 * q.ax2 * q.h2) ? 1 : max_ring;
          const int x0 = max(q.cx - rx, 0), x1 = min(q.cx + rx, g.div_b[0] - 1);
          if (x0 > x1) continue;
          const long long key0 = (long long)x0 + (long long)y * g.mul[1] + (long long)z * g.mul[2];
          lo = __ldg(grid.start + key0);
          hi = __ldg(grid.start + key0 + (x1 - x0) + 1);
          k++;
          return true;
        }
        return false;
      };
      // PREFETCH: once the first candidates of a row have been looked at, the table entries of the next row have
      // arrived; its first points are requested into L1 / L2 so that the next scan does not start with a DRAM miss.
      auto scan = [&](int lo, int hi, float row2, bool pf, int pf_lo, int pf_hi) {
        if (row2 > thr) {  // pruned by what was found since the row was chosen
          if (PREFETCH && pf && pf_hi > pf_lo) asm volatile("prefetch.global.L1 [%0];" ::"l"(grid.pts + pf_lo));
          return;
        }
        ncand += hi - lo;
        nrows++;
        bool first = true;
        if (STAGE) {
          auto issue = [&](int buf, int j, int rem) {
#pragma unroll
            for (int u = 0; u < 4; u++)
              if (u < rem) cp_async16(&s_stage[(buf * 4 + u) * kLoamBlock + tid], grid.pts + j + u);
            cp_async_commit();
          };
          int buf = 0;
          issue(0, lo, hi - lo);
#pragma unroll 1
          for (int j = lo; j < hi; j += 4) {
            const int rem = hi - j;
            if (rem > 4) { issue(buf ^ 1, j + 4, rem - 4); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            const float4 far = make_float4(INFINITY, 0.f, 0.f, 0.f);
            const float4 m0 = s_stage[(buf * 4 + 0) * kLoamBlock + tid];
            const float4 m1 = rem > 1 ? s_stage[(buf * 4 + 1) * kLoamBlock + tid] : far;
            const float4 m2 = rem > 2 ? s_stage[(buf * 4 + 2) * kLoamBlock + tid] : far;
            const float4 m3 = rem > 3 ? s_stage[(buf * 4 + 3) * kLoamBlock + tid] : far;
            consider(m0, j); consider(m1, j + 1); consider(m2, j + 2); consider(m3, j + 3);
            if (PREFETCH && first && pf && pf_hi > pf_lo) asm volatile("prefetch.global.L1 [%0];" ::"l"(grid.pts + pf_lo));
            first = false;
            buf ^= 1;
          }
          return;
        }
#pragma unroll 1
        for (int j = lo; j < hi; j += 4) {  // four independent float4 loads in flight
          const int rem = hi - j;
          const float4 far = make_float4(INFINITY, 0.f, 0.f, 0.f);  // f = inf: never considered
          const float4 m0 = __ldg(grid.pts + j);
          const float4 m1 = rem > 1 ? __ldg(grid.pts + j + 1) : far;
          const float4 m2 = rem > 2 ? __ldg(grid.pts + j + 2) : far;
          const float4 m3 = rem > 3 ? __ldg(grid.pts + j + 3) : far;
          consider(m0, j); consider(m1, j + 1); consider(m2, j + 2); consider(m3, j + 3);
          if (PREFETCH && first && pf && pf_hi > pf_lo) asm volatile("prefetch.global.L1 [%0];" ::"l"(grid.pts + pf_lo));
          first = false;
        }
      };
      int loA = 0, hiA = 0, loB = 0, hiB = 0;
      float r2A = 0.f, r2B = 0.f;
      bool haveA = next_row(loA, hiA, r2A);
#pragma unroll 1
      while (haveA) {
        const bool haveB = next_row(loB, hiB, r2B);
        scan(loA, hiA, r2A, haveB, loB, hiB);
        if (!haveB) break;
        haveA = next_row(loA, hiA, r2A);
        scan(loB, hiB, r2B, haveA, loA, hiA);
      }
      // a 6th candidate within 1e-6 (relative) of the 5th: the FP32 metric cannot tell which of them the reference keeps
      if (bj[4] >= 0 && !(f_out > bf[4] * 1.000001f)) {
        ncand = 0; nrows = 0;
        exact_search_one(grid, qf0, qf1, qf2, leaf, slack, max_ring, gate_thr, bj, ncand, nrows);
      }
    }
    // the winners' coordinates (cache-hot) + original map index; w = -1 marks an empty slot
#pragma unroll
    for (int k = 0; k < 5; k++) {
      float4 m = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      if (bj[k] >= 0) m = __ldg(grid.pts + bj[k]);
      nb_out[size_t(k) * stride + i] = m;
    }
    cnt_out[i] = make_int2(ncand, nrows);
  }
}

constexpr int kFitRow = 33;  // padded row of the transposed tile: lanes reading different rows hit different banks
struct FitWarpTile { double v[7][kFitRow]; };

template <bool DEBUG>
__global__ void __launch_bounds__(kLoamBlock, 3)
loam_fit_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, const float4* __restrict__ nb, const int2* __restrict__ cnt_in,
                size_t stride, LoamParams prm, LoamState* __restrict__ states, double* __restrict__ partials, int max_warps,
                pcr_loam_iter_log* __restrict__ logs, int apply_update, int32_t* __restrict__ dbg_knn, int32_t* __restrict__ dbg_status,
                const uint32_t* __restrict__ qperm) {
  const int scan = blockIdx.y;
  const uint32_t begin = offs[scan], end = offs[scan + 1];
  const uint32_t ns = end - begin;
  LoamState* st = states + scan;
  if (st->done) return;
  __shared__ double sT[16];
  __shared__ FitWarpTile s_tile[kLoamWarps];
  __shared__ double stot[kNV];
  __shared__ SolveScratch sc;
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
  __syncthreads();  // the only block barrier: warps are independent from here on
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w0 = begin + (blockIdx.x * kLoamWarps + warp) * 32u;  // first query of this warp
  if (w0 >= end) return;  // warps without a query take no ticket
  const int n_warps = int((ns + 31) / 32);
  const int gw = blockIdx.x * kLoamWarps + warp;  // warp index within the scan
  const uint32_t i = w0 + lane;
  const bool have = i < end;
  double J[6] = {0, 0, 0, 0, 0, 0}, E = 0.0;
  int status = -1;
  int2 cn = make_int2(0, 0);
  if (have) {
    const float4 po = __ldg(src + i);
    cn = __ldg(cnt_in + i);
    double qd[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      // the same transform as the search kernel (LoamRegister.cpp:128-130), then the float cast and back
      const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sT[r], double(po.x)), __dmul_rn(sT[4 + r], double(po.y))), __dmul_rn(sT[8 + r], double(po.z))), sT[12 + r]);
      qd[r] = double(__double2float_rn(v));
    }
    float nx[5], ny[5], nz[5];
    int nidx[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const float4 m = __ldg(nb + size_t(k) * stride + i);
      nx[k] = m.x; ny[k] = m.y; nz[k] = m.z; nidx[k] = __float_as_int(m.w);
    }
    const bool five = nidx[4] >= 0;
    double d5 = DBL_MAX;
    if (five) {
      // the search kernel selected the five on the FP32 metric (the SET is exact): order them by the exact FP64
      // (d2, original index) key, the order nanoflann returns them in
      unsigned long long ek[5];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const double dx = qd[0] - double(nx[k]), dyy = qd[1] - double(ny[k]), dzz = qd[2] - double(nz[k]);
        ek[k] = (unsigned long long)__double_as_longlong(dx * dx + dyy * dyy + dzz * dzz);
      }
      auto cswap = [&](int a, int b) {  // compare-exchange, a < b
        const bool sw = ek[b] < ek[a] || (ek[b] == ek[a] && nidx[b] < nidx[a]);
        const unsigned long long te = sw ? ek[a] : ek[b]; ek[a] = sw ? ek[b] : ek[a]; ek[b] = te;
        const float tx = sw ? nx[a] : nx[b]; nx[a] = sw ? nx[b] : nx[a]; nx[b] = tx;
        const float ty = sw ? ny[a] : ny[b]; ny[a] = sw ? ny[b] : ny[a]; ny[b] = ty;
        const float tz = sw ? nz[a] : nz[b]; nz[a] = sw ? nz[b] : nz[a]; nz[b] = tz;
        const int ti = sw ? nidx[a] : nidx[b]; nidx[a] = sw ? nidx[b] : nidx[a]; nidx[b] = ti;
      };
      cswap(0, 1); cswap(3, 4); cswap(2, 4); cswap(2, 3); cswap(0, 3); cswap(0, 2); cswap(1, 4); cswap(1, 3); cswap(1, 2);  // 9-exchange network
      d5 = __longlong_as_double((long long)ek[4]);
    }
    status = loam_point_residual(nx, ny, nz, five, d5, qd[0], qd[1], qd[2], po, prm, J, E);
    if (DEBUG) {
      const uint32_t iq = qperm ? __ldg(qperm + i) : i;  // original position of this query (queries are spatially re-ordered)
      if (dbg_knn) {
#pragma unroll
        for (int k = 0; k < 5; k++) dbg_knn[size_t(iq) * 5 + k] = status >= 1 ? nidx[k] : -1;
      }
      if (dbg_status) dbg_status[iq] = status;
    }
  }
  // ---- warp reduction through the transposed tile: lane p owns sum p (21 J J^T, 6 J E, count, candidates, rows)
  FitWarpTile& tl = s_tile[warp];
  const bool acc = status == 3;
#pragma unroll
  for (int r = 0; r < 6; r++) tl.v[r][lane] = acc ? J[r] : 0.0;
  tl.v[6][lane] = acc ? E : 0.0;
  __syncwarp();
  double sum = 0.0;
  if (lane < 27) {
    int ra, rb;
    if (lane < 21) {  // packed upper triangle index -> (row, col)
      int p = lane, r = 0;
      while (p >= 6 - r) { p -= 6 - r; r++; }
      ra = r; rb = r + p;
    } else { ra = lane - 21; rb = 6; }
    const double* a = tl.v[ra];
    const double* b = tl.v[rb];
#pragma unroll 8
    for (int l = 0; l < 32; l++) sum += a[l] * b[l];
  }
  const unsigned n_acc = __popc(__ballot_sync(0xffffffffu, acc));
  if (lane == 27) sum = double(n_acc);
  const unsigned c_cand = __reduce_add_sync(0xffffffffu, unsigned(cn.x)), c_rows = __reduce_add_sync(0xffffffffu, unsigned(cn.y));
  if (lane == 28) sum = double(c_cand);
  if (lane == 29) sum = double(c_rows);
  double* my = partials + (size_t(scan) * max_warps + gw) * kNV;
  if (lane < kNV) my[lane] = sum;
  __threadfence();
  __syncwarp();
  unsigned t = 0;
  if (lane == 0) t = atomicAdd(&st->ticket, 1u);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t != unsigned(n_warps - 1)) return;
  // ---- the last warp of the scan: fixed-order sum over the warp partials, then the solve
  __threadfence();
  if (lane < kNV) {
    const double* base = partials + size_t(scan) * max_warps * kNV + lane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four interleaved chains, combined in a fixed order
    int w = 0;
    for (; w + 4 <= n_warps; w += 4) {
      a0 += __ldcg(base + size_t(w) * kNV); a1 += __ldcg(base + size_t(w + 1) * kNV);
      a2 += __ldcg(base + size_t(w + 2) * kNV); a3 += __ldcg(base + size_t(w + 3) * kNV);
    }
    for (; w < n_warps; w++) a0 += __ldcg(base + size_t(w) * kNV);
    stot[lane] = (a0 + a1) + (a2 + a3);
  }
  __syncwarp();
  loam_solve_step(stot, sT, st, logs, scan, prm, apply_update, ns, lane, sc);
}

// ---- query re-ordering: Morton code of the scan-frame position (0.25 m cells, +-128 m), scan index on top -----------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void __launch_bounds__(256) loam_query_key_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs,
                                                             unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int scan = blockIdx.y;
  const uint32_t begin = offs[scan], end = offs[scan + 1];
  for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(src + i);
    const int qx = min(max(int(floorf(fminf(fmaxf(p.x, -1000.f), 1000.f) * 4.f)) + 512, 0), 1023);
    const int qy = min(max(int(floorf(fminf(fmaxf(p.y, -1000.f), 1000.f) * 4.f)) + 512, 0), 1023);
    const int qz = min(max(int(floorf(fminf(fmaxf(p.z, -1000.f), 1000.f) * 4.f)) + 512, 0), 1023);
    const uint32_t mc = spread10(uint32_t(qx)) | (spread10(uint32_t(qy)) << 1) | (spread10(uint32_t(qz)) << 2);
    keys[i] = ((unsigned long long)scan << 30) | mc;
    vals[i] = i;
  }
}
__global__ void __launch_bounds__(256) loam_query_gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ vals, size_t n,
                                                                float4* __restrict__ out) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = __ldg(src + vals[i]);
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

template <int LPQ, bool DEBUG>
static void opt_in_one() {
  PCR_CUDA_CHECK(cudaFuncSetAttribute(loam_iter_kernel<LPQ, DEBUG>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLoamDynSmem)));
}
static void loam_opt_in_smem() {
  static bool done_dev[64] = {false};
  int dev = 0;
  PCR_CUDA_CHECK(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return;
  opt_in_one<1, false>(); opt_in_one<2, false>(); opt_in_one<4, false>(); opt_in_one<8, false>();
  opt_in_one<1, true>(); opt_in_one<2, true>(); opt_in_one<4, true>(); opt_in_one<8, true>();
  done = true;
}

// lanes per query and queries per warp pass from the problem size: one lane per query once the queries alone fill the
// machine (batches), up to 8 lanes per query and short tiles for a single scan (latency)
static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}
static void pick_shape(size_t total_q, int& lpq, int& tile) {
  const size_t lanes_wave = size_t(kNumSMs) * 2 * kLoamBlock;
  lpq = 1;
  while (lpq < 8 && total_q * size_t(lpq) < 2 * lanes_wave) lpq <<= 1;
  const int forced = env_int("PCR_LOAM_LPQ", 0);  // tuning knob (1, 2, 4, 8)
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) lpq = forced;
  const int G = 32 / lpq;
  tile = 32;
  while (tile > G && (total_q / size_t(tile)) * 32 < lanes_wave) tile >>= 1;
  const int ftile = env_int("PCR_LOAM_TILE", 0);
  if (ftile >= G && ftile <= 32 && (ftile & (ftile - 1)) == 0) tile = ftile;
}

template <bool DEBUG>
static void launch_iter(int lpq, dim3 grid, cudaStream_t s, const float4* src, const uint32_t* offs, const GridView& view, const LoamParams& prm,
                        LoamState* states, double* partials, int max_blocks, pcr_loam_iter_log* logs, int apply, int tile, double slack,
                        int max_ring, int32_t* dbg_knn, int32_t* dbg_status) {
#define PCR_LOAM_LAUNCH(L)                                                                                                                      \
  loam_iter_kernel<L, DEBUG><<<grid, kLoamBlock, kLoamDynSmem, s>>>(src, offs, view, prm, states, partials, max_blocks, logs, apply, tile, slack, \
                                                                    max_ring, dbg_knn, dbg_status)
  switch (lpq) {
    case 1: PCR_LOAM_LAUNCH(1); break;
    case 2: PCR_LOAM_LAUNCH(2); break;
    case 4: PCR_LOAM_LAUNCH(4); break;
    default: PCR_LOAM_LAUNCH(8); break;
  }
#undef PCR_LOAM_LAUNCH
}

int loam_build_target(const float4* pts, size_t n, double max_knn_d2, CellGrid& grid, KeySort& ks, BBoxWork& bw, cudaStream_t s) {
  // gate radius r = sqrt(max_knn_d2) (LoamRegister.cpp:59 compares the SQUARED 5th distance with 1.0). With cells of
  // r / (R - 4 * slack) the R-ring cube around a query's cell contains every map point closer than r even after the float
  // rounding of the cell assignment (slack <= 2e-3 cells below ~2 km, grid_slack_cells()).
  // Which of the two cell sizes is used is decided BEFORE the index is built, from the number of gate-sized cells the map
  // occupies (bitmap + popcount, no sort): dense maps (more than PCR_LOAM_FINE_ABOVE points per occupied gate-sized cell)
  // get half-gate cells — most 5-NN balls then fit inside the 27 cells around the query, 8x fewer candidates — and the index
  // is built ONCE. If the half-gate table would not fit the dense-table budget the gate-sized grid is used instead.
  grid.built = false;
  grid.has_start = false;
  grid.n = n;
  grid.max_ring = 1;
  grid.occupied = 0;
  if (n == 0) return 0;
  const float r = std::sqrt(float(max_knn_d2));
  float bb[6];
  bbox_blocking(pts, n, bb, bb + 3, bw, s);
  if (bw.n_nonfinite) return kRetryNonFinite;
  const float coarse = r / (1.0f - 0.008f), fine = r / (2.0f - 0.008f);
  GridSpec gc;
  if (!make_grid_spec(bb, bb + 3, coarse, gc) || gc.ncell > (1ll << 29)) return -5;
  grid.occupied = count_occupied_cells(pts, n, gc, ks.tmp, ks.d_count, ks.h_count, s);
  const double per_cell = double(n) / double(std::max<size_t>(grid.occupied, 1));
  if (per_cell > double(env_int("PCR_LOAM_FINE_ABOVE", 6))) {
    GridSpec gf;
    if (make_grid_spec(bb, bb + 3, fine, gf) && gf.ncell <= (1ll << 29)) {
      const int rc = build_cell_grid(pts, n, fine, grid, ks, bw, s, true, bb);
      if (rc == 0) { grid.max_ring = 2; return 0; }
    }
  }
  const int rc = build_cell_grid(pts, n, coarse, grid, ks, bw, s, true, bb);
  grid.max_ring = 1;
  return rc;
}

LoamDriver::~LoamDriver() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
}

static GridView make_view(const CellGrid& grid) { return view_of(grid); }

static void launch_search(bool prefetch, bool stage, dim3 g, cudaStream_t s, const float4* q, const uint32_t* offs, const GridView& view, const LoamParams& prm,
                          const LoamState* states, double slack, int max_ring, float4* nb, int2* cnt, size_t stride) {
  if (stage && prefetch) loam_search_kernel<true, true><<<g, kLoamBlock, 0, s>>>(q, offs, view, prm, states, slack, max_ring, nb, cnt, stride);
  else if (stage) loam_search_kernel<false, true><<<g, kLoamBlock, 0, s>>>(q, offs, view, prm, states, slack, max_ring, nb, cnt, stride);
  else if (prefetch) loam_search_kernel<true, false><<<g, kLoamBlock, 0, s>>>(q, offs, view, prm, states, slack, max_ring, nb, cnt, stride);
  else loam_search_kernel<false, false><<<g, kLoamBlock, 0, s>>>(q, offs, view, prm, states, slack, max_ring, nb, cnt, stride);
}

const float4* LoamDriver::sort_queries(const float4* src, const uint32_t* d_offs, size_t n_scans, size_t total_q, size_t max_pts, const uint32_t** perm,
                                       cudaStream_t s) {
  q_keys0.ensure(total_q); q_keys1.ensure(total_q); q_vals0.ensure(total_q); q_vals1.ensure(total_q); q_sorted.ensure(total_q);
  const unsigned bx = unsigned(std::min<size_t>((max_pts + 255) / 256, 64));
  loam_query_key_kernel<<<dim3(bx, unsigned(n_scans)), 256, 0, s>>>(src, d_offs, q_keys0.p, q_vals0.p);
  int bits = 30;
  while ((size_t(1) << (bits - 30)) < n_scans) bits++;
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, q_keys0.p, q_keys1.p, q_vals0.p, q_vals1.p, int(total_q), 0, bits, s);
  q_tmp.ensure(bytes);
  PCR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(q_tmp.p, bytes, q_keys0.p, q_keys1.p, q_vals0.p, q_vals1.p, int(total_q), 0, bits, s));
  loam_query_gather_kernel<<<unsigned((total_q + 255) / 256), 256, 0, s>>>(src, q_vals1.p, total_q, q_sorted.p);
  launches += 3;
  *perm = q_vals1.p;
  return q_sorted.p;
}

int LoamDriver::align(const float4* src, const size_t* offs, size_t n_scans, const CellGrid& grid, const LoamParams& prm,
                      double* T, int32_t* converged, int32_t* iters_out, int64_t* n_last_out, bool profile, cudaStream_t s) {
  launches = 0; cand_total = 0; rows_total = 0; pt_evals = 0; hot_ms = 0.f; hot_launches = 0;
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
  int tile = 32, lpq = 1;
  pick_shape(total_q, lpq, tile);
  const double slack = grid_slack_cells(grid.g);
  const size_t per_block = size_t(kLoamWarps) * tile;
  int max_blocks = int((max_pts + per_block - 1) / per_block);
  max_blocks = std::max(1, std::min(max_blocks, int(size_t(kNumSMs) * 2 / n_scans)));  // all scans' blocks resident in one wave
  // large batches (one lane per query, full tiles): search and fit as two kernels per iteration ("Batch path" above)
  const bool split = lpq == 1 && tile == 32 && env_int("PCR_LOAM_SPLIT", 1) != 0;
  // L1 prefetch of the next row's first points: +1..2.5 % when the map lives in DRAM (18.6 M points), -1 % when it fits in L2
  const bool search_prefetch = env_int("PCR_LOAM_PREFETCH", grid.n > (size_t(4) << 20) ? 1 : 0) != 0;
  const bool search_stage = env_int("PCR_LOAM_STAGE", 0) != 0;
  // the fit kernel: one query per thread, one partial per warp
  const int fit_blocks = std::max(1, int((max_pts + kLoamBlock - 1) / kLoamBlock));
  const int max_warps = fit_blocks * kLoamWarps;
  states.ensure(n_scans);
  offsets.ensure(n_scans + 1);
  partials.ensure(n_scans * size_t(split ? max_warps : max_blocks) * kNV);
  logs.ensure(n_scans * size_t(prm.max_iters));
  PCR_CUDA_CHECK(cudaMemcpyAsync(states.p, hs, n_scans * sizeof(LoamState), cudaMemcpyHostToDevice, s));
  PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, (n_scans + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  GridView view = make_view(grid);
  dim3 gridDim(max_blocks, unsigned(n_scans));
  if (profile) {
    if (!ev0) { PCR_CUDA_CHECK(cudaEventCreate(&ev0)); PCR_CUDA_CHECK(cudaEventCreate(&ev1)); }
    PCR_CUDA_CHECK(cudaEventRecord(ev0, s));
  }
  last_lpq = lpq; last_tile = tile; last_split = split;
  if (grid.built && grid.has_start && max_pts > 0) {
    const float4* q = src;
    const uint32_t* perm = nullptr;
    if (split) {
      nb_buf.ensure(5 * total_q);
      cnt_buf.ensure(total_q);
      q = sort_queries(src, offsets.p, n_scans, total_q, max_pts, &perm, s);
    }
    const size_t search_pb = size_t(kLoamWarps) * 32;
    const dim3 sgrid(unsigned((max_pts + search_pb - 1) / search_pb), unsigned(n_scans));
    const dim3 fgrid(static_cast<unsigned>(fit_blocks), static_cast<unsigned>(n_scans));
    for (int it = 0; it < prm.max_iters; it++) {
      if (split) {
        launch_search(search_prefetch, search_stage, sgrid, s, q, offsets.p, view, prm, states.p, slack, grid.max_ring, nb_buf.p, cnt_buf.p, total_q);
        loam_fit_kernel<false><<<fgrid, kLoamBlock, 0, s>>>(q, offsets.p, nb_buf.p, cnt_buf.p, total_q, prm, states.p, partials.p, max_warps, logs.p, 1,
                                                            nullptr, nullptr, perm);
        launches += 2;
      } else {
        launch_iter<false>(lpq, gridDim, s, src, offsets.p, view, prm, states.p, partials.p, max_blocks, logs.p, 1, tile, slack, grid.max_ring, nullptr, nullptr);
        launches++;
      }
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
  if (split && env_int("PCR_LOAM_HIST", 0)) {
    // diagnostics of the lane divergence in loam_search_kernel: rows / candidates per query of the LAST iteration, and how
    // much of a warp's lock-step work its lanes really need (sum over the 32 queries / 32 x the largest)
    std::vector<int2> hc(total_q);
    PCR_CUDA_CHECK(cudaMemcpy(hc.data(), cnt_buf.p, total_q * sizeof(int2), cudaMemcpyDeviceToHost));
    long long hist_r[27] = {0}, hist_c[12] = {0};
    double sum_r = 0, max_r = 0, sum_c = 0, max_c = 0, sum_ch = 0, max_ch = 0;
    for (size_t w = 0; w + 32 <= total_q; w += 32) {
      int mr = 0, mc = 0, mch = 0;
      for (int l = 0; l < 32; l++) {
        const int c = hc[w + l].x, r = hc[w + l].y, ch = (c + 3) / 4 + r;  // four candidates per step + one step per row
        hist_r[std::min(r, 26)]++;
        hist_c[std::min(c / 16, 11)]++;
        sum_r += r; sum_c += c; sum_ch += ch;
        mr = std::max(mr, r); mc = std::max(mc, c); mch = std::max(mch, ch);
      }
      max_r += 32.0 * mr; max_c += 32.0 * mc; max_ch += 32.0 * mch;
    }
    std::fprintf(stderr, "[pcr loam hist] queries %zu | rows/query:", total_q);
    for (int r = 0; r < 27; r++) if (hist_r[r]) std::fprintf(stderr, " %d:%.1f%%", r, 100.0 * double(hist_r[r]) / double(total_q));
    std::fprintf(stderr, " | candidates/query (x16):");
    for (int c = 0; c < 12; c++) if (hist_c[c]) std::fprintf(stderr, " %d:%.1f%%", c * 16, 100.0 * double(hist_c[c]) / double(total_q));
    std::fprintf(stderr, " | warp lock-step efficiency: rows %.2f candidates %.2f steps (rows + candidates / 4) %.2f\n", sum_r / std::max(max_r, 1.0),
                 sum_c / std::max(max_c, 1.0), sum_ch / std::max(max_ch, 1.0));
  }
  for (size_t i = 0; i < n_scans; i++) {
    for (int q = 0; q < 16; q++) T[i * 16 + q] = hs[i].T[q];
    if (converged) converged[i] = hs[i].converged;
    if (iters_out) iters_out[i] = hs[i].iters;
    if (n_last_out) n_last_out[i] = hs[i].n_last;
    cand_total += hs[i].cand_total;
    rows_total += hs[i].rows_total;
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
  int tile = 32, lpq = 1;
  pick_shape(ns, lpq, tile);
  // the same kernel selection as align(): with one lane per query and full tiles (large batches, or forced through the
  // PCR_LOAM_LPQ / PCR_LOAM_TILE knobs) the linearisation runs as the search kernel followed by the fit kernel
  const bool split = lpq == 1 && tile == 32 && env_int("PCR_LOAM_SPLIT", 1) != 0;
  const double slack = grid_slack_cells(grid.g);
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
  last_split = false;
  if (ns > 0 && grid.built && grid.has_start) {
    GridView view = make_view(grid);
    if (split) {
      nb_buf.ensure(5 * ns);
      cnt_buf.ensure(ns);
      const uint32_t* perm = nullptr;
      const float4* q = sort_queries(src, offsets.p, 1, ns, ns, &perm, s);
      const size_t search_pb = size_t(kLoamWarps) * 32;
      const dim3 sgrid(unsigned((ns + search_pb - 1) / search_pb), 1);
      const int fit_blocks = std::max(1, int((ns + kLoamBlock - 1) / kLoamBlock));
      partials.ensure(size_t(fit_blocks) * kLoamWarps * kNV);
      launch_search(env_int("PCR_LOAM_PREFETCH", grid.n > (size_t(4) << 20) ? 1 : 0) != 0, env_int("PCR_LOAM_STAGE", 0) != 0, sgrid, s, q, offsets.p, view, prm,
                    states.p, slack, grid.max_ring, nb_buf.p, cnt_buf.p, ns);
      loam_fit_kernel<true><<<dim3(fit_blocks, 1), kLoamBlock, 0, s>>>(q, offsets.p, nb_buf.p, cnt_buf.p, ns, prm, states.p, partials.p, fit_blocks * kLoamWarps,
                                                                       logs.p, 0, dbg_knn.p, dbg_status.p, perm);
      last_split = true;
    } else {
      launch_iter<true>(lpq, dim3(max_blocks, 1), s, src, offsets.p, view, prm, states.p, partials.p, max_blocks, logs.p, 0, tile, slack,
                        grid.max_ring, dbg_knn.p, dbg_status.p);
    }
  }
  last_lpq = lpq; last_tile = tile;
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
