// pclomp NDT on the GPU. Reference: third_parties/pclomp/src/ndt_omp_impl.hpp, voxel_grid_covariance_omp_impl.hpp.
#include "ndt.cuh"
#include "dev_linalg.cuh"
#include "host_math.hpp"
#include <cstdlib>
#include <cfloat>
#include <algorithm>
#include <cstdlib>

namespace pcr {

constexpr int kNdtBlock = 128;
constexpr int kNdtNV = 29;  // score, g[6], H upper 21, (point, leaf) pairs

// ================================================================================================================
// N1. target voxel grid: per-leaf FP64 moments (ascending original index, like the reference's serial pass),
// covariance, eigenvalue inflation, inverse — VoxelGridCovariance::applyFilter (voxel_grid_covariance_omp_impl.hpp:209-367)
// ================================================================================================================
__global__ void __launch_bounds__(128)
ndt_leaf_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                const uint32_t* __restrict__ seg_start, size_t nseg, int min_points, double eig_mult,
                NdtLeafRec* __restrict__ recs, int32_t* __restrict__ okeys, int32_t* __restrict__ onpts,
                double* __restrict__ omean, double* __restrict__ ocov, double* __restrict__ oicov, int32_t* __restrict__ table,
                float4* __restrict__ centroids) {
  size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (l >= nseg) return;
  const uint32_t b = seg_start[l], e = seg_start[l + 1];
  const int n = int(e - b);
  const uint32_t key = keys[b];
  double sx = 0, sy = 0, sz = 0;
  // Leaf() starts cov_ at Identity (pclomp/voxel_grid_covariance_omp.h:107) and `cov_ += pt pt^T` (impl :237)
  double cxx = 1, cxy = 0, cxz = 0, cyy = 1, cyz = 0, czz = 1;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (uint32_t j = b; j < e; j++) {
    const float4 p = __ldg(pts + j);  // `pts` is gathered into key order: a leaf is one contiguous run
    const double x = p.x, y = p.y, z = p.z;
    sx += x; sy += y; sz += z;
    cxx += x * x; cxy += x * y; cxz += x * z; cyy += y * y; cyz += y * z; czz += z * z;  // exact products
    fx = __fadd_rn(fx, p.x); fy = __fadd_rn(fy, p.y); fz = __fadd_rn(fz, p.z);
  }
  const double dn = double(n);
  double mean[3] = {sx / dn, sy / dn, sz / dn};
  const double sum[3] = {sx, sy, sz};
  double cov[3][3] = {{cxx, cxy, cxz}, {cxy, cyy, cyz}, {cxz, cyz, czz}};
  double icov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  int npts = n;
  bool in_cloud = false;
  if (n >= min_points) {
    in_cloud = true;
    double c2[3][3];
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) c2[r][q] = (cov[r][q] - 2 * (sum[r] * mean[q])) / dn + mean[r] * mean[q];  // :329
    const double f = (dn - 1.0) / dn;
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) cov[r][q] = c2[r][q] * f;  // :330
    double w[3], V[3][3];
    eig_sym3(cov, w, V);
    if (w[0] < 0 || w[1] < 0 || w[2] <= 0) {
      npts = -1;  // :337-341
    } else {
      const double mn = eig_mult * w[2];
      if (w[0] < mn) {
        w[0] = mn;
        if (w[1] < mn) w[1] = mn;
        double Vi[3][3];
        inv3(V, Vi);
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) {
            double v = 0;
            for (int k = 0; k < 3; k++) v += (V[r][k] * w[k]) * Vi[k][q];
            cov[r][q] = v;  // :355 cov = evecs * diag * evecs^-1
          }
      }
      inv3(cov, icov);
      double mx = -DBL_MAX, mi = DBL_MAX;
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) { mx = fmax(mx, icov[r][q]); mi = fmin(mi, icov[r][q]); }
      if (mx == INFINITY || mi == -INFINITY) npts = -1;  // :360-364
    }
  } else {
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) cov[r][q] = (r == q) ? 1.0 : 0.0;
  }
  NdtLeafRec rec;
  for (int r = 0; r < 3; r++) {
    rec.mean[r] = mean[r];
    omean[l * 3 + r] = mean[r];
    for (int q = 0; q < 3; q++) {
      rec.icov[r * 3 + q] = float(icov[r][q]);
      ocov[l * 9 + r * 3 + q] = cov[r][q];
      oicov[l * 9 + r * 3 + q] = icov[r][q];
    }
  }
  rec.npts = npts;
  recs[l] = rec;
  okeys[l] = int32_t(key);
  onpts[l] = npts;
  // dense cell -> leaf table: l for usable leaves; -2 - l for leaves that entered the centroid cloud (n >= min_points) but
  // were rejected by the eigenvalue / inverse checks — only the KDTREE radius search still sees those (its centroid
  // kd-tree is built before the checks, voxel_grid_covariance_omp_impl.hpp:297-326 vs :337-364)
  if (npts >= min_points) table[key] = int32_t(l);
  else if (in_cloud) table[key] = -2 - int32_t(l);
  const float fn = float(n);
  centroids[l] = make_float4(__fdiv_rn(fx, fn), __fdiv_rn(fy, fn), __fdiv_rn(fz, fn), in_cloud ? 1.f : 0.f);
}

static void gauss_params(double resolution, double outlier_ratio, double& d1, double& d2, double& d3) {
  // ndt_omp_impl.hpp:86-93
  const double c1 = 10 * (1 - outlier_ratio);
  const double c2 = outlier_ratio / std::pow(resolution, 3);
  d3 = -std::log(c2);
  d1 = -std::log(c1 + c2) - d3;
  d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / d1);
}

int ndt_build_target(const float4* pts, size_t n, const pcr_params& prm, NdtTarget& tgt, KeySort& ks, BBoxWork& bw, cudaStream_t s) {
  tgt.built = false;
  tgt.overflow = false;
  tgt.nleaves = 0;
  tgt.resolution = prm.ndt_resolution;
  gauss_params(double(prm.ndt_resolution), prm.ndt_outlier_ratio, tgt.d1, tgt.d2, tgt.d3);
  if (n == 0) { tgt.overflow = true; tgt.built = true; return 0; }
  float mn[3], mx[3];
  bbox_blocking(pts, n, mn, mx, bw, s);
  if (bw.n_nonfinite) return kRetryNonFinite;
  if (!make_grid_spec(mn, mx, prm.ndt_resolution, tgt.g)) {  // :79-84 leaf size too small -> no leaves
    tgt.overflow = true;
    tgt.built = true;
    return 0;
  }
  if (tgt.g.ncell > (1ll << 29)) return PCR_ERR_GRID_TOO_LARGE;
  ks.sort(pts, n, tgt.g, s);
  ks.segment(s);
  const size_t L = ks.nseg;
  tgt.nleaves = L;
  tgt.recs.ensure(L); tgt.keys.ensure(L); tgt.npts.ensure(L);
  tgt.mean.ensure(L * 3); tgt.cov.ensure(L * 9); tgt.icov.ensure(L * 9);
  tgt.centroids.ensure(L);
  tgt.table.ensure(size_t(tgt.g.ncell));
  PCR_CUDA_CHECK(cudaMemsetAsync(tgt.table.p, 0xff, size_t(tgt.g.ncell) * sizeof(int32_t), s));
  gather_sorted(pts, ks, s);
  ndt_leaf_kernel<<<unsigned((L + 127) / 128), 128, 0, s>>>(ks.sorted_pts.p, ks.keys, ks.vals, ks.seg_start.p, L, prm.ndt_min_points, prm.ndt_eig_mult,
                                                           tgt.recs.p, tgt.keys.p, tgt.npts.p, tgt.mean.p, tgt.cov.p, tgt.icov.p,
                                                           tgt.table.p, tgt.centroids.p);
  PCR_CUDA_CHECK(cudaGetLastError());
  tgt.built = true;
  return 0;
}

// ================================================================================================================
// N2 + N3 + N4. Source points are strided over the threads of a request's blocks. Per point: float transform,
// DIRECT{7,1,26} voxel lookups (all table loads issued up front), then per valid leaf the FP32 score / gradient /
// Hessian terms of updateDerivatives, accumulated in FP64 registers (leaf records software-prefetched one ahead).
// Deterministic warp-shuffle + block + last-block reduction; the last block of a request then advances the scan's
// Newton / More-Thuente state machine on the device (ndt_round_kernel below).
// ================================================================================================================
struct NdtTargetView {
  const NdtLeafRec* recs;
  const double* mean;
  const double* icov;
  const int32_t* table;
  const float4* centroids;
  GridSpec g;
  float d2f;
  float radius2;
  double d1, d2;
};

template <int SEARCH>
struct NbTraits;
template <> struct NbTraits<PCR_NDT_DIRECT7> { static constexpr int N = 7; };
template <> struct NbTraits<PCR_NDT_DIRECT1> { static constexpr int N = 1; };
template <> struct NbTraits<PCR_NDT_DIRECT26> { static constexpr int N = 26; };
template <> struct NbTraits<PCR_NDT_KDTREE> { static constexpr int N = 27; };

template <int SEARCH>
__device__ __forceinline__ void nb_offset(int ni, int& ox, int& oy, int& oz) {
  if (SEARCH == PCR_NDT_KDTREE) {  // all 27 cells: a centroid within `resolution` of the point lies in one of them
    ox = ni / 9 - 1; oy = (ni / 3) % 3 - 1; oz = ni % 3 - 1;
  } else if (SEARCH == PCR_NDT_DIRECT26) {  // the 26 non-centre cells, x-major (pcl::getAllNeighborCellIndices)
    const int k = ni >= 13 ? ni + 1 : ni;
    ox = k / 9 - 1; oy = (k / 3) % 3 - 1; oz = k % 3 - 1;
  } else {  // centre, +x, -x, +y, -y, +z, -z (voxel_grid_covariance_omp_impl.hpp:423-430)
    ox = (ni == 1) - (ni == 2); oy = (ni == 3) - (ni == 4); oz = (ni == 5) - (ni == 6);
  }
}

// Per-thread FP64 accumulators live in shared memory, which frees ~58 registers per thread and lets more warps hide the
// table / leaf-record latency. Components 2j and 2j + 1 of a thread sit next to each other (a double2 at [j][tid]:
// conflict-free 128-bit accesses), so one LDS.128 / STS.128 pair serves two read-modify-writes: 59 instead of 87
// instructions for the 29 accumulations of a (point, leaf) pair. Components of a pair are added in ascending order; the
// even one waits in a register for the odd one. Every component still sees its own values in the same order: same bits.
constexpr int kNdtPairs = (kNdtNV + 1) / 2;
struct SmemAcc {
  double2* p;  // &sacc2[tid]
  double lo;
  __device__ __forceinline__ void add(int k, double v) {  // k is a compile-time constant after unrolling
    if ((k & 1) == 0) {
      lo = v;
    } else {
      double2 t = p[(k >> 1) * kNdtBlock];
      t.x += lo;
      t.y += v;
      p[(k >> 1) * kNdtBlock] = t;
    }
  }
  // an even component whose odd partner is not added this time
  __device__ __forceinline__ void flush_even(int k) { reinterpret_cast<double*>(p + (k >> 1) * kNdtBlock)[0] += lo; }
};
// component k of thread t in the double view of the accumulator array
__device__ __forceinline__ int ndt_acc_index(int k, int t) { return (((k >> 1) * kNdtBlock + t) << 1) + (k & 1); }

struct LeafRegs { float4 a, b, c, d; };
__device__ __forceinline__ LeafRegs load_leaf(const NdtLeafRec* recs, int id) {
  const float4* rp = reinterpret_cast<const float4*>(recs + id);
  LeafRegs r;
  r.a = __ldg(rp); r.b = __ldg(rp + 1); r.c = __ldg(rp + 2); r.d = __ldg(rp + 3);
  return r;
}

// float path: computePointDerivatives (:399-440) + updateDerivatives (:485-537) for one (point, leaf) pair
__device__ __forceinline__ void ndt_pair_f32(const LeafRegs& L, const float (&pt)[3], const float (&xj)[8], const float (&xh)[15], bool hess,
                                             float d2f, double d1, SmemAcc acc) {
  const double m0 = __hiloint2double(__float_as_int(L.a.y), __float_as_int(L.a.x));
  const double m1 = __hiloint2double(__float_as_int(L.a.w), __float_as_int(L.a.z));
  const double m2 = __hiloint2double(__float_as_int(L.b.y), __float_as_int(L.b.x));
  const float C[3][3] = {{L.b.z, L.b.w, L.c.x}, {L.c.y, L.c.z, L.c.w}, {L.d.x, L.d.y, L.d.z}};
  const float xt[3] = {float(double(pt[0]) - m0), float(double(pt[1]) - m1), float(double(pt[2]) - m2)};
  float xC[3];
#pragma unroll
  for (int c = 0; c < 3; c++) xC[c] = (xt[0] * C[0][c] + xt[1] * C[1][c]) + xt[2] * C[2][c];
  const float q = (xt[0] * xC[0] + xt[1] * xC[1]) + xt[2] * xC[2];
  float e = expf(-d2f * q * 0.5f);
  const float score_inc = float(-d1 * double(e));
  e = d2f * e;
  if (e > 1.f || e < 0.f || e != e) return;  // :506-507 contributes nothing, not even score
  e = float(double(e) * d1);
  float CJ[3][6];  // C * J, J = [I | Jang]
#pragma unroll
  for (int r = 0; r < 3; r++) {
    CJ[r][0] = C[r][0]; CJ[r][1] = C[r][1]; CJ[r][2] = C[r][2];
    CJ[r][3] = C[r][1] * xj[0] + C[r][2] * xj[1];
    CJ[r][4] = (C[r][0] * xj[2] + C[r][1] * xj[3]) + C[r][2] * xj[4];
    CJ[r][5] = (C[r][0] * xj[5] + C[r][1] * xj[6]) + C[r][2] * xj[7];
  }
  float xCJ[6];
#pragma unroll
  for (int c = 0; c < 6; c++) xCJ[c] = (xt[0] * CJ[0][c] + xt[1] * CJ[1][c]) + xt[2] * CJ[2][c];
  acc.add(0, double(score_inc));
#pragma unroll
  for (int c = 0; c < 6; c++) acc.add(1 + c, double(e * xCJ[c]));  // component 6 waits for 7 (or is flushed below)
  if (hess) {
    float JCJ[6][6];  // J^T C J (rows 0..2 are CJ itself)
#pragma unroll
    for (int c = 0; c < 6; c++) {
      JCJ[0][c] = CJ[0][c]; JCJ[1][c] = CJ[1][c]; JCJ[2][c] = CJ[2][c];
      JCJ[3][c] = xj[0] * CJ[1][c] + xj[1] * CJ[2][c];
      JCJ[4][c] = (xj[2] * CJ[0][c] + xj[3] * CJ[1][c]) + xj[4] * CJ[2][c];
      JCJ[5][c] = (xj[5] * CJ[0][c] + xj[6] * CJ[1][c]) + xj[7] * CJ[2][c];
    }
    // x^T C H_E blocks (only i, j >= 3 are nonzero): a b c / b d e / c e f
    const float ha = xC[1] * xh[0] + xC[2] * xh[1];
    const float hb = xC[1] * xh[2] + xC[2] * xh[3];
    const float hc = xC[1] * xh[4] + xC[2] * xh[5];
    const float hd = (xC[0] * xh[6] + xC[1] * xh[7]) + xC[2] * xh[8];
    const float he = (xC[0] * xh[9] + xC[1] * xh[10]) + xC[2] * xh[11];
    const float hf = (xC[0] * xh[12] + xC[1] * xh[13]) + xC[2] * xh[14];
    const float xH[3][3] = {{ha, hb, hc}, {hb, hd, he}, {hc, he, hf}};
    int k = 7;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = r; c < 6; c++) {
        const float hx = (r >= 3) ? xH[r - 3][c - 3] : 0.f;
        acc.add(k++, double(e * (-d2f * xCJ[r] * xCJ[c] + hx + JCJ[c][r])));
      }
  } else {
    acc.flush_even(6);
  }
  acc.add(28, 1.0);
  acc.flush_even(28);
}

// double path: computeHessian / updateHessian (:541-645) for one pair
__device__ __forceinline__ void ndt_pair_f64(const NdtTargetView& tgt, int id, const float (&pt)[3], const double (&xj)[8],
                                             const double (&xh)[15], SmemAcc acc) {
  const double Jc[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, xj[0], xj[1]}, {xj[2], xj[3], xj[4]}, {xj[5], xj[6], xj[7]}};
  double C[3][3], xt[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    xt[r] = double(pt[r]) - __ldg(tgt.mean + size_t(id) * 3 + r);
#pragma unroll
    for (int c = 0; c < 3; c++) C[r][c] = __ldg(tgt.icov + size_t(id) * 9 + r * 3 + c);
  }
  double Cx[3];
#pragma unroll
  for (int r = 0; r < 3; r++) Cx[r] = C[r][0] * xt[0] + C[r][1] * xt[1] + C[r][2] * xt[2];
  double e = tgt.d2 * exp(-tgt.d2 * (xt[0] * Cx[0] + xt[1] * Cx[1] + xt[2] * Cx[2]) / 2);
  if (e > 1 || e < 0 || e != e) return;
  e *= tgt.d1;
  acc.lo = 0.0;  // component 6 (gradient, untouched by computeHessian) gets + 0.0 with its partner 7
  double CJ[6][3], xCJ[6];  // C * J_i and x^T C J_i
#pragma unroll
  for (int c = 0; c < 6; c++) {
#pragma unroll
    for (int r = 0; r < 3; r++) CJ[c][r] = C[r][0] * Jc[c][0] + C[r][1] * Jc[c][1] + C[r][2] * Jc[c][2];
    xCJ[c] = xt[0] * CJ[c][0] + xt[1] * CJ[c][1] + xt[2] * CJ[c][2];
  }
  auto xCh = [&](double h0, double h1, double h2) {
    return xt[0] * (C[0][0] * h0 + C[0][1] * h1 + C[0][2] * h2) + xt[1] * (C[1][0] * h0 + C[1][1] * h1 + C[1][2] * h2) +
           xt[2] * (C[2][0] * h0 + C[2][1] * h1 + C[2][2] * h2);
  };
  const double ha = xCh(0, xh[0], xh[1]), hb = xCh(0, xh[2], xh[3]), hc = xCh(0, xh[4], xh[5]);
  const double hd = xCh(xh[6], xh[7], xh[8]), he = xCh(xh[9], xh[10], xh[11]), hf = xCh(xh[12], xh[13], xh[14]);
  const double xH[3][3] = {{ha, hb, hc}, {hb, hd, he}, {hc, he, hf}};
  int k = 7;
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = r; c < 6; c++) {
      const double hx = (r >= 3) ? xH[r - 3][c - 3] : 0.0;
      const double jd = Jc[c][0] * CJ[r][0] + Jc[c][1] * CJ[r][1] + Jc[c][2] * CJ[r][2];
      acc.add(k++, e * (-tgt.d2 * xCJ[r] * xCJ[c] + hx + jd));
    }
  acc.add(28, 1.0);
  acc.flush_even(28);
}

// ================================================================================================================
// One evaluation ROUND for every scan that is waiting for this kind of evaluation (float computeDerivatives or
// double-path computeHessian). Flat 1-D grid = one resident wave. Every block
//   1. compacts the list of scans whose stamp says "evaluate me in this round" (deterministic order, no atomics),
//   2. takes its share of the (request, sub-block) items: the wave is re-partitioned among the ACTIVE scans every round,
//   3. evaluates its points (FP32 pair math, FP64 shared-memory accumulators), writes its partial sums,
//   4. and, if it is the last block of its request, sums the partials in fixed order and advances the scan's Newton /
//      More-Thuente state machine (ndt_logic.cuh) by this result: the next evaluation's transform and angle tables are
//      written to the scan's state and its stamp is set for the kernel that has to serve it (the double-path kernel of
//      the same round, or the float kernel of the next round).
// Stamps are only ever written for LATER kernels, so every block of a launch sees the same request list.
// ================================================================================================================
constexpr int kNdtMaxBatch = 2048;  // scans per driver chunk (request list lives in shared memory)

static_assert(sizeof(NdtScanState) % 16 == 0, "NdtScanState is copied as uint4 words");

__device__ __noinline__ void ndt_state_step(NdtScanState* st, const double* totals, NdtCfg cfg) { ndt_logic::on_result(*st, totals, cfg); }

// ndt_logic::fill_request by the lanes of one warp: the six double and six float sines / cosines of the next evaluation's
// angles are evaluated by twelve lanes at once, then one lane assembles the angular derivative tables and another the float
// pose matrix — the same host_math.hpp functions on the same values as the serial version, a few microseconds shorter per
// evaluation round (the tail of the last block is on the critical path of every round).
__device__ __forceinline__ void ndt_fill_request_warp(NdtScanState* st, double* trig /* shared, 12 doubles */, int lane) {
  if (st->pend == NDT_PEND_NONE) return;  // warp-uniform (shared state)
  if (lane < 3) {
    double c, s;
    hm::ndt_angle_trig(st->eval_p[3 + lane], c, s);
    trig[lane] = c; trig[3 + lane] = s;
  } else if (lane < 6 && st->fill_matrix) {
    const float ang = static_cast<float>(st->eval_p[lane]);  // lane 3..5 -> angle index 3..5
    trig[6 + 2 * (lane - 3)] = double(hm::sin_f32(ang));
    trig[7 + 2 * (lane - 3)] = double(hm::cos_f32(ang));
  }
  __syncwarp();
  if (lane == 0) {
    hm::ndt_angle_tables_trig(trig[0], trig[1], trig[2], trig[3], trig[4], trig[5], st->next.j_ang, st->next.h_ang, st->next.j_ang_d, st->next.h_ang_d);
  } else if (lane == 1 && st->fill_matrix) {
    const float sc[6] = {float(trig[6]), float(trig[7]), float(trig[8]), float(trig[9]), float(trig[10]), float(trig[11])};
    hm::ndt_pose_matrix_sc_f32(st->eval_p, sc, st->final_T);
  }
  __syncwarp();
  if (lane < 16) st->next.Tf[lane] = st->final_T[lane];
  __syncwarp();
}

#ifndef PCR_NDT_FLOAT_BLOCKS
#define PCR_NDT_FLOAT_BLOCKS 5  // resident blocks per SM of the float-path kernel (96 registers); 6 = 80 registers, measured slower
#endif
constexpr int kNdtGroup = 32;  // blocks per first-level group of a request's two-level reduction

// totals of n_rows rows of kNdtNV partial sums: four interleaved slices per component (coalesced: a step of the loop reads
// four consecutive rows), combined in a fixed order. The accumulator columns in `sacc` are free when this runs.
__device__ __forceinline__ void ndt_reduce_rows(const double* __restrict__ rows, int n_rows, double* sacc, double* s_tot, int tid) {
  if (tid < 4 * kNdtNV) {
    const int comp = tid % kNdtNV, slice = tid / kNdtNV;
    const double* pb = rows + comp;
    double t0 = 0.0, t1 = 0.0;
    int b = slice;
    for (; b + 4 < n_rows; b += 8) {
      t0 += __ldcg(pb + size_t(b) * kNdtNV);
      t1 += __ldcg(pb + size_t(b + 4) * kNdtNV);
    }
    if (b < n_rows) t0 += __ldcg(pb + size_t(b) * kNdtNV);
    sacc[slice * kNdtNV + comp] = t0 + t1;
  }
  __syncthreads();
  if (tid < kNdtNV) s_tot[tid] = (sacc[tid] + sacc[kNdtNV + tid]) + (sacc[2 * kNdtNV + tid] + sacc[3 * kNdtNV + tid]);
  __syncthreads();
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int SEARCH, bool DOUBLE_PATH>
__global__ void __launch_bounds__(kNdtBlock, DOUBLE_PATH ? 2 : PCR_NDT_FLOAT_BLOCKS)
ndt_round_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, NdtTargetView tgt, NdtScanState* __restrict__ states,
                 NdtScanOut* __restrict__ outs, int32_t* __restrict__ float_round, int32_t* __restrict__ hess_round, int n_scans, int round,
                 int step, NdtCfg cfg, double* __restrict__ partials, unsigned* __restrict__ tickets, NdtProgress* progress,
                 NdtCounters* __restrict__ counters, int* __restrict__ round_flags, int max_bpr, double* __restrict__ gpart,
                 unsigned* __restrict__ gtickets) {
  constexpr int NNB = NbTraits<SEARCH>::N;
  // programmatic dependent launch: let the next kernel of the stream be scheduled, then wait until the previous grid has
  // completed and its writes are visible (both are no-ops for a launch without the attribute)
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // round_flags[r]: bit 0 = some scan wants a float evaluation in round r, bit 1 = a double-path Hessian. Written only by
  // tails of EARLIER kernels (or, for bit 1 of this round, by the float kernel that has completed): a launch without
  // work returns at once (most double-path launches, and the rounds queued past the end of the registration)
  if (blockIdx.x == 0 && threadIdx.x == 0 && !DOUBLE_PATH && progress) {
    progress->finished = counters->finished;  // complete: every earlier kernel of the stream has ended
    progress->round = round;
  }
  if (step && !(round_flags[round] & (DOUBLE_PATH ? 2 : 1))) return;
  if (cfg.trace && blockIdx.x == 0 && threadIdx.x == 0) counters->t_tail[4] = globaltimer_ns();
  __shared__ __align__(16) NdtScanState s_state;
  __shared__ __align__(16) double sacc[2 * kNdtPairs * kNdtBlock];  // double2 [kNdtPairs][kNdtBlock]
  __shared__ double s_tot[kNdtNV + 1];
  __shared__ double s_trig[12];
  __shared__ unsigned short s_list[kNdtMaxBatch];
  __shared__ int s_warp_cnt[kNdtBlock / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- 1. request list of this launch
  const int32_t* stamps = DOUBLE_PATH ? hess_round : float_round;
  const int per = (n_scans + kNdtBlock - 1) / kNdtBlock;
  const int lo = min(tid * per, n_scans), hi = min(lo + per, n_scans);
  int cnt = 0;
  for (int k = lo; k < hi; k++) cnt += (stamps[k] == round) ? 1 : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp_cnt[warp] = incl;
  __syncthreads();
  int base = 0, n_active = 0;
#pragma unroll
  for (int w = 0; w < kNdtBlock / 32; w++) {
    if (w < warp) base += s_warp_cnt[w];
    n_active += s_warp_cnt[w];
  }
  int pos = base + incl - cnt;
  for (int k = lo; k < hi; k++)
    if (stamps[k] == round) s_list[pos++] = (unsigned short)k;
  if (blockIdx.x == 0 && tid == 0 && n_active > 0) atomicAdd(&counters->work_launches, 1);
  if (n_active == 0) return;
  __syncthreads();

  // ---- 2. my items. Every active scan is cut into `bpr` equal items; item t goes to block t mod gridDim. bpr is the split
  // (up to ~6 items per block) that leaves the fewest blocks idle in the last pass: n_active * bpr / gridDim just below an
  // integer. With bpr = gridDim / n_active alone a batch of ~500 scans would keep a third of the wave idle, and 1024 scans
  // would make a quarter of the blocks do twice the work of the others.
  int bpr = 1;
  {
    const int G = int(gridDim.x);
    const int cap = max(1, min(max_bpr, (6 * G + n_active - 1) / n_active));
    float best = 0.f;
    for (int m = 1; m <= 6; m++) {  // candidates: the largest split that still fits m passes of the wave
      const int b = min(cap, (m * G) / n_active);
      if (b < 1) continue;
      const float load = float(n_active) * float(b) / float(G);
      const float eff = load / ceilf(load);
      if (eff > best + 0.03f) { best = eff; bpr = b; }  // more, smaller items only when they balance clearly better
    }
  }
  const int n_items = n_active * bpr;
  if (cfg.trace && blockIdx.x == 0 && tid == 0) counters->t_tail[6] = globaltimer_ns();
  const GridSpec& g = tgt.g;
  SmemAcc acc{reinterpret_cast<double2*>(sacc) + tid, 0.0};
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int req = item / bpr, sub = item - req * bpr;
    const int scan = s_list[req];
    const uint32_t begin = offs[scan], end = offs[scan + 1];
    const int nb = max(1, min(int((end - begin + kNdtBlock - 1) / kNdtBlock), bpr));  // blocks working on this request
    if (sub >= nb) continue;  // block-uniform
    {
      const int nwords = sizeof(NdtScanState) / 16;
      const uint4* gp = reinterpret_cast<const uint4*>(states + scan);
      uint4* sp4 = reinterpret_cast<uint4*>(&s_state);
      for (int k = tid; k < nwords; k += kNdtBlock) sp4[k] = gp[k];
    }
#pragma unroll
    for (int j = 0; j < kNdtPairs; j++) reinterpret_cast<double2*>(sacc)[j * kNdtBlock + tid] = make_double2(0.0, 0.0);
    __syncthreads();
    const NdtEvalParams& sp = s_state.next;
    const bool hess = sp.compute_hessian != 0;

    for (uint32_t i = begin + sub * kNdtBlock + tid; i < end; i += uint32_t(nb) * kNdtBlock) {
      const float4 po = __ldg(src + i);
      // pcl::transformPointCloud (float): ((m00 x + m01 y) + m02 z) + m03
      float pt[3];
#pragma unroll
      for (int r = 0; r < 3; r++)
        pt[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(sp.Tf[r], po.x), __fmul_rn(sp.Tf[4 + r], po.y)), __fmul_rn(sp.Tf[8 + r], po.z)),
                          sp.Tf[12 + r]);
      // voxel of the transformed point: floor(p / leaf)  (voxel_grid_covariance_omp_impl.hpp:379-381)
      float fi[3];
      bool ok = true;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        fi[a] = floorf(__fdiv_rn(pt[a], g.leaf[a]));
        if (!(fabsf(fi[a]) < 1.0e9f)) ok = false;
      }
      if (!ok) continue;
      const int ijk[3] = {int(fi[0]), int(fi[1]), int(fi[2])};
      // all neighbourhood table lookups in flight at once
      int ids[NNB];
#pragma unroll
      for (int ni = 0; ni < NNB; ni++) {
        int ox, oy, oz;
        nb_offset<SEARCH>(ni, ox, oy, oz);
        const int cx = ijk[0] + ox, cy = ijk[1] + oy, cz = ijk[2] + oz;
        const bool inb = cx >= g.min_b[0] && cx <= g.max_b[0] && cy >= g.min_b[1] && cy <= g.max_b[1] && cz >= g.min_b[2] && cz <= g.max_b[2];
        const long long key = (long long)(cx - g.min_b[0]) * g.mul[0] + (long long)(cy - g.min_b[1]) * g.mul[1] +
                              (long long)(cz - g.min_b[2]) * g.mul[2];
        ids[ni] = inb ? __ldg(tgt.table + key) : -1;
      }
      if (SEARCH == PCR_NDT_KDTREE) {
        // N6: VoxelGridCovariance::radiusSearch — leaves of the centroid cloud whose float centroid is within `resolution`
        // of the point: FLANN L2_Simple<float> metric (x->y->z float accumulate), strict d2 < r^2
#pragma unroll
        for (int ni = 0; ni < NNB; ni++) {
          int id = ids[ni];
          if (id <= -2) id = -2 - id;
          if (id >= 0) {
            const float4 c = __ldg(tgt.centroids + id);
            const float dx = __fsub_rn(pt[0], c.x), dy = __fsub_rn(pt[1], c.y), dz = __fsub_rn(pt[2], c.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (!(d2 < tgt.radius2)) id = -1;
          }
          ids[ni] = id;
        }
      }
      unsigned mask = 0;
#pragma unroll
      for (int ni = 0; ni < NNB; ni++) mask |= (ids[ni] >= 0) ? (1u << ni) : 0u;
      if (!mask) continue;
      auto pick = [&](int k) {
        int v = ids[0];
#pragma unroll
        for (int ni = 1; ni < NNB; ni++) v = (k == ni) ? ids[ni] : v;
        return v;
      };
      if (!DOUBLE_PATH) {
        float xj[8], xh[15];
#pragma unroll
        for (int r = 0; r < 8; r++) xj[r] = (sp.j_ang[r][0] * po.x + sp.j_ang[r][1] * po.y) + sp.j_ang[r][2] * po.z;
#pragma unroll
        for (int r = 0; r < 15; r++) xh[r] = hess ? (sp.h_ang[r][0] * po.x + sp.h_ang[r][1] * po.y) + sp.h_ang[r][2] * po.z : 0.f;
        // neighbours in reference order, next leaf record prefetched while the current one is evaluated
        int k = __ffs(mask) - 1;
        mask &= mask - 1;
        LeafRegs cur = load_leaf(tgt.recs, pick(k));
        while (true) {
          LeafRegs nxt = cur;
          const bool more = mask != 0;
          if (more) {
            k = __ffs(mask) - 1;
            mask &= mask - 1;
            nxt = load_leaf(tgt.recs, pick(k));
          }
          ndt_pair_f32(cur, pt, xj, xh, hess, tgt.d2f, tgt.d1, acc);
          if (!more) break;
          cur = nxt;
        }
      } else {
        const double x[3] = {double(po.x), double(po.y), double(po.z)};
        double xj[8], xh[15];
#pragma unroll
        for (int r = 0; r < 8; r++) xj[r] = x[0] * sp.j_ang_d[r][0] + x[1] * sp.j_ang_d[r][1] + x[2] * sp.j_ang_d[r][2];
#pragma unroll
        for (int r = 0; r < 15; r++) xh[r] = x[0] * sp.h_ang_d[r][0] + x[1] * sp.h_ang_d[r][1] + x[2] * sp.h_ang_d[r][2];
        while (mask) {
          const int k = __ffs(mask) - 1;
          mask &= mask - 1;
          ndt_pair_f64(tgt, pick(k), pt, xj, xh, acc);
        }
      }
    }

    // ---- 3. fixed-order block reduction straight out of shared memory: warp w owns components w, w+4, ...
    __syncthreads();
    if (cfg.trace && blockIdx.x == 0 && tid == 0 && item == int(blockIdx.x)) counters->t_tail[7] = globaltimer_ns();
    for (int k = warp; k < kNdtNV; k += kNdtBlock / 32) {
      double v = ((sacc[ndt_acc_index(k, lane)] + sacc[ndt_acc_index(k, lane + 32)]) + sacc[ndt_acc_index(k, lane + 64)]) +
                 sacc[ndt_acc_index(k, lane + 96)];
      v = warp_sum(v);
      if (lane == 0) partials[size_t(item) * kNdtNV + k] = v;
    }
    if (lane == 0) __threadfence();  // one fence per writer warp, after all of its partial sums
    __syncthreads();
    if (cfg.trace && blockIdx.x == 0 && tid == 0 && item == int(blockIdx.x)) counters->t_tail[8] = globaltimer_ns();
    // ---- 4. block partials -> totals in a fixed order, by whichever block finishes last. A single scan owns the whole wave
    // (740 partial rows, 170 KB): one block reading them all took as long as the evaluation itself, so requests with more
    // than kNdtGroup blocks are reduced in two levels - the last block of every group of kNdtGroup consecutive blocks adds
    // up its group's rows, the last of those adds up the group rows.
    const bool trace = cfg.trace && tid == 0;
    const int n_groups = (nb + kNdtGroup - 1) / kNdtGroup;
    const double* rows = partials + size_t(req) * bpr * kNdtNV;
    int n_rows = nb;
    if (n_groups > 1) {
      const int gid = sub / kNdtGroup, gsize = min(kNdtGroup, nb - gid * kNdtGroup);
      const size_t gslot = size_t(req) * ((bpr + kNdtGroup - 1) / kNdtGroup);
      if (tid == 0) {
        const unsigned t = atomicAdd(gtickets + gslot + gid, 1u);
        s_last = (t == unsigned(gsize - 1));
      }
      __syncthreads();
      if (!s_last) continue;  // block-uniform
      __threadfence();
      ndt_reduce_rows(rows + size_t(gid) * kNdtGroup * kNdtNV, gsize, sacc, s_tot, tid);
      if (tid < kNdtNV) {
        gpart[(gslot + gid) * kNdtNV + tid] = s_tot[tid];
        __threadfence();
      }
      if (tid == 0) gtickets[gslot + gid] = 0;
      __syncthreads();
      rows = gpart + gslot * kNdtNV;
      n_rows = n_groups;
    }
    if (tid == 0) {
      const unsigned t = atomicAdd(tickets + scan, 1u);
      s_last = (t == unsigned(n_rows - 1));
    }
    __syncthreads();
    if (!s_last) continue;  // block-uniform
    __threadfence();
    if (trace) counters->t_tail[0] = globaltimer_ns();
    ndt_reduce_rows(rows, n_rows, sacc, s_tot, tid);
    if (trace) counters->t_tail[1] = globaltimer_ns();
    if (tid < 30) outs[scan].v[tid] = tid < kNdtNV ? s_tot[tid] : 0.0;
    if (tid == 0) {
      tickets[scan] = 0;
      atomicAdd(reinterpret_cast<unsigned long long*>(&counters->point_evals), (unsigned long long)(end - begin));
    }
    if (step && warp == 0) {
      if (lane == 0) ndt_state_step(&s_state, s_tot, cfg);  // decides the next evaluation ...
      __syncwarp();
      if (trace) counters->t_tail[2] = globaltimer_ns();
      ndt_fill_request_warp(&s_state, s_trig, lane);          // ... whose transform and derivative tables the warp fills in
    }
    __syncthreads();
    if (trace) counters->t_tail[3] = globaltimer_ns();
    if (step) {
      const int nwords = sizeof(NdtScanState) / 16;
      uint4* gp = reinterpret_cast<uint4*>(states + scan);
      const uint4* sp4 = reinterpret_cast<const uint4*>(&s_state);
      for (int k = tid; k < nwords; k += kNdtBlock) gp[k] = sp4[k];
      if (tid == 0) {
        if (s_state.pend == NDT_PEND_FLOAT) { float_round[scan] = round + 1; atomicOr(round_flags + round + 1, 1); }
        else if (s_state.pend == NDT_PEND_DOUBLE) { hess_round[scan] = round; atomicOr(round_flags + round, 2); }  // served by the double-path kernel of this round
      }
      if (s_state.pend == NDT_PEND_NONE) {
        NdtScanOut* o = outs + scan;
        if (tid < 16) o->final_T[tid] = s_state.final_T[tid];
        if (tid == 0) {
          o->converged = s_state.converged; o->nr_iterations = s_state.nr_iterations;
          o->n_evals = s_state.n_evals; o->n_hess = s_state.n_hess; o->n_pairs = s_state.n_pairs; o->score = s_state.score;
          const int f = atomicAdd(&counters->finished, 1) + 1;
          if (f == n_scans && progress) { __threadfence_system(); progress->all_done = 1; }
        }
      }
    }
    if (trace) counters->t_tail[5] = globaltimer_ns();
    __syncthreads();  // s_state / s_tot are reused by the next item
  }
}

// start of a registration: one WARP per scan. Lane 0 runs ndt_logic::start (guess -> p, first request), the warp fills in
// the request's transform and derivative tables exactly as the evaluation kernels' tails do (ndt_fill_request_warp).
// Empty scans and an empty target evaluate to all-zero sums (no block would contribute), which the state machine digests
// right here.
constexpr int kNdtInitWarps = 4;
__global__ void __launch_bounds__(kNdtInitWarps * 32)
ndt_init_kernel(const double* __restrict__ guesses, const uint32_t* __restrict__ offs, NdtScanState* __restrict__ states,
                NdtScanOut* __restrict__ outs, int32_t* __restrict__ float_round, int32_t* __restrict__ hess_round, int first_scan,
                int count, int n_scans, int first_round, int no_target, NdtCfg cfg, NdtProgress* progress,
                NdtCounters* __restrict__ counters, int* __restrict__ round_flags) {
  __shared__ __align__(16) NdtScanState s_st[kNdtInitWarps];
  __shared__ double s_trig[kNdtInitWarps][12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * kNdtInitWarps + warp;
  if (k >= count) return;  // warp-uniform
  const int s = first_scan + k;
  NdtScanState* st = &s_st[warp];
  if (lane == 0) {
    ndt_logic::start(*st, guesses + size_t(s) * 16, s);
    hess_round[s] = -1;
    if (no_target || offs[s + 1] == offs[s]) {
      double zeros[kNdtNV];
      for (int q = 0; q < kNdtNV; q++) zeros[q] = 0.0;
      while (st->pend != NDT_PEND_NONE) ndt_logic::on_result(*st, zeros, cfg);
    }
  }
  __syncwarp();
  ndt_fill_request_warp(st, s_trig[warp], lane);
  {
    const int nwords = sizeof(NdtScanState) / 16;
    uint4* gp = reinterpret_cast<uint4*>(states + s);
    const uint4* sp4 = reinterpret_cast<const uint4*>(st);
    for (int q = lane; q < nwords; q += 32) gp[q] = sp4[q];
  }
  if (st->pend == NDT_PEND_NONE) {
    NdtScanOut* o = outs + s;
    if (lane < 16) o->final_T[lane] = st->final_T[lane];
    if (lane == 0) {
      float_round[s] = -1;
      o->converged = st->converged; o->nr_iterations = st->nr_iterations; o->n_evals = st->n_evals; o->n_hess = st->n_hess;
      o->n_pairs = st->n_pairs; o->score = st->score;
      const int f = atomicAdd(&counters->finished, 1) + 1;
      if (f == n_scans && progress) { __threadfence_system(); progress->all_done = 1; }
    }
  } else if (lane == 0) {
    float_round[s] = first_round;  // joins the batch in the next round to be launched
    atomicOr(round_flags + first_round, 1);
  }
}

NdtDriver::~NdtDriver() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (progress) cudaFreeHost(progress);
}

void NdtDriver::ensure_progress() {
  if (!progress) PCR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&progress), sizeof(NdtProgress), cudaHostAllocMapped));
}

void NdtDriver::prepare(size_t n_scans, int grid_blocks, cudaStream_t s) {
  ensure_progress();
  states.ensure(n_scans); outs.ensure(n_scans); stamps.ensure(2 * n_scans);
  // rows: items <= 6 * grid + n_active (+ slack); behind them the group rows of the two-level reductions (a request has
  // groups only when it is cut into more than kNdtGroup items: at most items / kNdtGroup + one per such request of them)
  const size_t item_rows = size_t(7) * size_t(std::max(grid_blocks, 1)) + 2 * n_scans + 64;
  const size_t group_rows = size_t(std::max(grid_blocks, 1)) + 64;
  partials.ensure((item_rows + group_rows) * kNdtNV);
  gpart_ = partials.p + item_rows * kNdtNV;
  counters.ensure(1);
  if (tickets.cap < n_scans + group_rows) {  // all zero between kernels: the last block to arrive resets its ticket
    tickets.ensure(n_scans + group_rows);
    PCR_CUDA_CHECK(cudaMemsetAsync(tickets.p, 0, tickets.cap * sizeof(unsigned), s));
  }
  gtickets_ = tickets.p + n_scans;
  PCR_CUDA_CHECK(cudaMemsetAsync(counters.p, 0, sizeof(NdtCounters), s));
  round_flags.ensure(size_t(max_rounds_cap) + 2);
  PCR_CUDA_CHECK(cudaMemsetAsync(round_flags.p, 0, (size_t(max_rounds_cap) + 2) * sizeof(int), s));
  PCR_CUDA_CHECK(cudaMemsetAsync(stamps.p, 0xff, 2 * n_scans * sizeof(int32_t), s));  // -1: not (yet) part of the batch
  progress->round = -1;
  progress->all_done = 0;
  progress->finished = 0;
}

static NdtTargetView make_view(const NdtTarget& tgt) {
  NdtTargetView v;
  v.recs = tgt.recs.p; v.mean = tgt.mean.p; v.icov = tgt.icov.p; v.table = tgt.table.p; v.g = tgt.g;
  v.centroids = tgt.centroids.p;
  v.d1 = tgt.d1; v.d2 = tgt.d2; v.d2f = float(tgt.d2);
  v.radius2 = float(double(tgt.resolution) * double(tgt.resolution));  // KdTreeFLANN::radiusSearch: (float)(radius * radius)
  return v;
}

// Round kernels are launched with programmatic stream serialisation: the next kernel of the stream may be scheduled while
// this one drains (every block releases its dependents at once and waits for the complete previous grid — griddepcontrol —
// before it reads anything), which takes the launch latency out of the chain float kernel -> double kernel -> next round.
template <int SEARCH>
static void launch_round_t(bool dbl, int grid, cudaStream_t s, const float4* src, const uint32_t* offs, const NdtTargetView& v, NdtScanState* states,
                           NdtScanOut* outs, int32_t* fr, int32_t* hr, int n, int round, int step, const NdtCfg& cfg, double* partials,
                           unsigned* tickets, NdtProgress* prog, NdtCounters* cnt, int* flags, int max_bpr, double* gpart, unsigned* gtickets) {
  // Measured (profiles/README.md): a single scan / a batch of 8 gain 3-7 %, the 1024-scan job loses 2 % -> small batches only.
  static const int pdl_env = [] { const char* e = std::getenv("PCR_NDT_PDL"); return e ? (std::atoi(e) != 0 ? 1 : 0) : -1; }();
  const bool pdl = pdl_env >= 0 ? pdl_env != 0 : n < 64;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(unsigned(dbl ? std::min(grid, kNumSMs * 2) : grid));  // double path: 212 registers, two blocks per SM, one resident wave
  lc.blockDim = dim3(kNdtBlock);
  lc.dynamicSmemBytes = 0;
  lc.stream = s;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = &attr;
  lc.numAttrs = pdl ? 1 : 0;
  if (dbl)
    PCR_CUDA_CHECK(cudaLaunchKernelEx(&lc, ndt_round_kernel<SEARCH, true>, src, offs, v, states, outs, fr, hr, n, round, step, cfg, partials, tickets, prog, cnt, flags, max_bpr, gpart, gtickets));
  else
    PCR_CUDA_CHECK(cudaLaunchKernelEx(&lc, ndt_round_kernel<SEARCH, false>, src, offs, v, states, outs, fr, hr, n, round, step, cfg, partials, tickets, prog, cnt, flags, max_bpr, gpart, gtickets));
}

void NdtDriver::launch_round(const float4* src, const NdtTarget& tgt, int search, int n, int round, int step, const NdtCfg& cfg, int grid_blocks,
                             bool float_kernel, bool double_kernel, cudaStream_t s) {
  const NdtTargetView v = make_view(tgt);
  NdtProgress* dprog = nullptr;
  PCR_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dprog), progress, 0));
  int32_t* fr = stamps.p;
  int32_t* hr = stamps.p + n;
  for (int part = 0; part < 2; part++) {
    if (part == 0 ? !float_kernel : !double_kernel) continue;
    const bool dbl = part == 1;
    switch (search) {
      case PCR_NDT_DIRECT1: launch_round_t<PCR_NDT_DIRECT1>(dbl, grid_blocks, s, src, offsets.p, v, states.p, outs.p, fr, hr, n, round, step, cfg, partials.p, tickets.p, dprog, counters.p, round_flags.p, max_bpr_, gpart_, gtickets_); break;
      case PCR_NDT_DIRECT26: launch_round_t<PCR_NDT_DIRECT26>(dbl, grid_blocks, s, src, offsets.p, v, states.p, outs.p, fr, hr, n, round, step, cfg, partials.p, tickets.p, dprog, counters.p, round_flags.p, max_bpr_, gpart_, gtickets_); break;
      case PCR_NDT_KDTREE: launch_round_t<PCR_NDT_KDTREE>(dbl, grid_blocks, s, src, offsets.p, v, states.p, outs.p, fr, hr, n, round, step, cfg, partials.p, tickets.p, dprog, counters.p, round_flags.p, max_bpr_, gpart_, gtickets_); break;
      default: launch_round_t<PCR_NDT_DIRECT7>(dbl, grid_blocks, s, src, offsets.p, v, states.p, outs.p, fr, hr, n, round, step, cfg, partials.p, tickets.p, dprog, counters.p, round_flags.p, max_bpr_, gpart_, gtickets_); break;
    }
    launches++;
  }
}

static int wave_blocks(size_t n_scans, size_t max_pts) {
  // one resident wave (PCR_NDT_FLOAT_BLOCKS blocks of 128 threads per SM), never more than one block per 128 points
  const size_t full = (max_pts + kNdtBlock - 1) / kNdtBlock;
  return int(std::max<size_t>(1, std::min<size_t>(size_t(kNumSMs) * PCR_NDT_FLOAT_BLOCKS, full * n_scans)));
}

void NdtDriver::evaluate_one(const float4* src, size_t ns, const NdtTarget& tgt, int search, const NdtEvalParams& ep, NdtEvalResult& out,
                             cudaStream_t s) {
  for (int k = 0; k < 30; k++) out.v[k] = 0.0;
  if (tgt.overflow || tgt.nleaves == 0 || ns == 0) return;  // no block would contribute
  const int grid = wave_blocks(1, ns);
  max_bpr_ = int((ns + kNdtBlock - 1) / kNdtBlock);
  prepare(1, grid, s);
  uint32_t* ho = h_offsets.ensure(2);
  ho[0] = 0; ho[1] = uint32_t(ns);
  offsets.ensure(2);
  PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  NdtScanState* hs = h_state.ensure(1);
  memset(hs, 0, sizeof(NdtScanState));
  hs->next = ep;
  hs->next.scan = 0;
  PCR_CUDA_CHECK(cudaMemcpyAsync(states.p, hs, sizeof(NdtScanState), cudaMemcpyHostToDevice, s));
  // stamps: the scan wants exactly this kind of evaluation in round 0
  int32_t* hst = reinterpret_cast<int32_t*>(h_counters.ensure(1));
  hst[0] = ep.kind == 0 ? 0 : -1;
  hst[1] = ep.kind == 0 ? -1 : 0;
  PCR_CUDA_CHECK(cudaMemcpyAsync(stamps.p, hst, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  NdtCfg cfg{};
  launch_round(src, tgt, search, 1, 0, 0, cfg, grid, ep.kind == 0, ep.kind != 0, s);
  NdtScanOut* ho_out = h_outs.ensure(1);
  PCR_CUDA_CHECK(cudaMemcpyAsync(ho_out, outs.p, sizeof(NdtScanOut), cudaMemcpyDeviceToHost, s));
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  PCR_CUDA_CHECK(cudaGetLastError());
  for (int k = 0; k < 30; k++) out.v[k] = ho_out->v[k];
}

// ================================================================================================================
// N5. Registration driver: the host only queues evaluation rounds. Each round = the float-path kernel followed by the
// double-path kernel (which serves the computeHessian requests raised by the float kernel's tails in the same round).
// How many rounds a batch needs is decided on the device; the host keeps a small number of rounds queued ahead of the
// GPU (it reads the round counter and the all-done flag from host-mapped memory, never synchronising the stream) and
// stops queueing when the last scan has finished. Kernels of rounds queued past that point find no request and return.
// ================================================================================================================
int NdtDriver::align(const float4* src, const size_t* offs, size_t n_scans, const NdtTarget& tgt, const pcr_params& prm, double* T,
                     int32_t* converged, int32_t* iters, double* trans_prob, bool profile, cudaStream_t s, const NdtArrivals* arrivals) {
  launches = 0; hot_ms = 0.f; hot_launches = 0; total_evals = 0; total_hess = 0; total_pairs = 0; point_evals = 0; rounds = 0;
  if (n_scans == 0) return 0;
  if (arrivals && n_scans > max_streaming_scans) throw CudaError("NDT: streaming admission supports at most 2048 scans per batch");
  NdtCfg cfg{};
  cfg.step_size = prm.ndt_step_size;
  cfg.trans_eps = prm.ndt_trans_eps;
  cfg.max_iters = prm.ndt_max_iters;
  static const int tail_trace = [] { const char* v = std::getenv("PCR_NDT_TAIL_TRACE"); return v ? std::atoi(v) : 0; }();
  cfg.trace = tail_trace;
  const int no_target = (tgt.overflow || tgt.nleaves == 0) ? 1 : 0;
  static const int lookahead = [] { const char* v = std::getenv("PCR_NDT_LOOKAHEAD"); return v ? std::max(1, std::atoi(v)) : 3; }();
  // every outer iteration costs at most 1 + kMaxStepIterations float evaluations (+ 1 double-path Hessian in the same round);
  // with arrivals the rounds go on for as long as scans keep joining
  const int max_rounds = arrivals ? max_rounds_cap - 2 : std::min((prm.ndt_max_iters + 3) * (ndt_logic::kMaxStepIterations + 2) + 2, max_rounds_cap);
  if (profile && !ev0) { PCR_CUDA_CHECK(cudaEventCreate(&ev0)); PCR_CUDA_CHECK(cudaEventCreate(&ev1)); }

  for (size_t c0 = 0; c0 < n_scans; c0 += kNdtMaxBatch) {  // chunks of scans (request list size); normally a single chunk
    const size_t n = std::min<size_t>(kNdtMaxBatch, n_scans - c0);
    size_t max_pts = 0;
    for (size_t i = 0; i < n; i++) max_pts = std::max(max_pts, size_t(offs[c0 + i + 1] - offs[c0 + i]));
    const int grid = wave_blocks(n, std::max<size_t>(max_pts, 1));
    max_bpr_ = std::max(1, int((max_pts + kNdtBlock - 1) / kNdtBlock));
    prepare(n, grid, s);
    uint32_t* ho = h_offsets.ensure(n + 1);
    double* hg = h_guesses.ensure(n * 16);
    for (size_t i = 0; i <= n; i++) ho[i] = uint32_t(offs[c0 + i] - offs[0]);
    memcpy(hg, T + c0 * 16, n * 16 * sizeof(double));
    offsets.ensure(n + 1);
    guesses.ensure(n * 16);
    PCR_CUDA_CHECK(cudaMemcpyAsync(offsets.p, ho, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    PCR_CUDA_CHECK(cudaMemcpyAsync(guesses.p, hg, n * 16 * sizeof(double), cudaMemcpyHostToDevice, s));
    NdtProgress* dprog = nullptr;
    PCR_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dprog), progress, 0));
    int r = 0;
    size_t next_group = 0, admitted = 0;
    const size_t n_groups = arrivals ? arrivals->n_groups : 1;
    auto admit = [&](size_t k) {  // start the state machines of group k; they join in round r
      const size_t a = arrivals ? arrivals->first[k] : 0, b = arrivals ? arrivals->first[k + 1] : n;
      if (arrivals) PCR_CUDA_CHECK(cudaStreamWaitEvent(s, arrivals->event(k), 0));
      if (b > a)
        ndt_init_kernel<<<unsigned((b - a + kNdtInitWarps - 1) / kNdtInitWarps), kNdtInitWarps * 32, 0, s>>>(guesses.p, offsets.p, states.p, outs.p, stamps.p, stamps.p + n, int(a), int(b - a),
                                                                  int(n), r, no_target, cfg, dprog, counters.p, round_flags.p);
      admitted = b;
      launches++;
    };
    if (profile) PCR_CUDA_CHECK(cudaEventRecord(ev0, s));
    if (arrivals) arrivals->wait(0);
    admit(next_group++);
    unsigned spins = 0;
    bool stream_idle = false;
    while (r < max_rounds) {
      // scans whose points have arrived meanwhile join the batch; if every admitted scan has finished, wait for the next group
      while (next_group < n_groups) {
        if (!arrivals->ready(next_group)) {
          if (size_t(progress->finished) < admitted) break;
          arrivals->wait(next_group);
        }
        admit(next_group++);
      }
      // throttle: at most `lookahead` rounds queued ahead of the round the GPU is working on
      while (!progress->all_done && progress->round < r - lookahead) {
        if ((++spins & 0xfff) == 0) {  // safety net: the stream drained (or failed) without the expected progress
          const cudaError_t q = cudaStreamQuery(s);
          if (q == cudaSuccess) { stream_idle = true; break; }
          if (q != cudaErrorNotReady) PCR_CUDA_CHECK(q);
        }
      }
      if (progress->all_done) break;
      if (stream_idle) {
        if (next_group >= n_groups) break;
        stream_idle = false;  // the queue ran dry while groups are still outstanding: go on
      }
      launch_round(src, tgt, prm.ndt_search, int(n), r, 1, cfg, grid, true, true, s);
      r++;
    }
    if (profile) PCR_CUDA_CHECK(cudaEventRecord(ev1, s));
    NdtScanOut* hout = h_outs.ensure(n);
    NdtCounters* hc = h_counters.ensure(1);
    PCR_CUDA_CHECK(cudaMemcpyAsync(hout, outs.p, n * sizeof(NdtScanOut), cudaMemcpyDeviceToHost, s));
    PCR_CUDA_CHECK(cudaMemcpyAsync(hc, counters.p, sizeof(NdtCounters), cudaMemcpyDeviceToHost, s));
    PCR_CUDA_CHECK(cudaStreamSynchronize(s));
    PCR_CUDA_CHECK(cudaGetLastError());
    if (hc->finished != int(n)) throw CudaError("NDT: the evaluation rounds ended before every scan finished");
    if (cfg.trace) {  // the LAST request tail that ran: ns from its kernel's start
      const unsigned long long* t = hc->t_tail;
      std::fprintf(stderr, "[pcr ndt tail] kernel start -> last block %.1f us, totals +%.1f, state step +%.1f, request filled +%.1f, state stored +%.1f"
                   " | block 0: request list +%.1f, its points +%.1f, partial row written +%.1f\n",
                   1e-3 * double((long long)(t[0] - t[4])), 1e-3 * double((long long)(t[1] - t[0])), 1e-3 * double((long long)(t[2] - t[1])),
                   1e-3 * double((long long)(t[3] - t[2])), 1e-3 * double((long long)(t[5] - t[3])), 1e-3 * double((long long)(t[6] - t[4])),
                   1e-3 * double((long long)(t[7] - t[6])), 1e-3 * double((long long)(t[8] - t[7])));
    }
    if (profile) {
      float ms = 0.f;
      PCR_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
      hot_ms += ms;
    }
    hot_launches += hc->work_launches;
    point_evals += hc->point_evals;
    rounds += r;
    for (size_t i = 0; i < n; i++) {
      const NdtScanOut& o = hout[i];
      for (int q = 0; q < 16; q++) T[(c0 + i) * 16 + q] = double(o.final_T[q]);  // NdtRegister.cpp:28
      if (converged) converged[c0 + i] = o.converged ? 1 : 0;
      if (iters) iters[c0 + i] = o.nr_iterations;
      const size_t ns = offs[c0 + i + 1] - offs[c0 + i];
      if (trans_prob) trans_prob[c0 + i] = o.score / double(ns);
      total_evals += o.n_evals;
      total_hess += o.n_hess;
      total_pairs += o.n_pairs;
    }
  }
  return 0;
}

}  // namespace pcr
