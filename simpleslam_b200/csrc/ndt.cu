// pclomp NDT on the GPU. Reference: third_parties/pclomp/src/ndt_omp_impl.hpp, voxel_grid_covariance_omp_impl.hpp.
#include "ndt.cuh"
#include "dev_linalg.cuh"
#include "host_math.hpp"
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
// Deterministic warp-shuffle + block + last-block reduction; the last block writes the 29 sums straight into
// host-mapped pinned memory, so one kernel launch + one stream sync is the whole evaluation.
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

// Per-thread FP64 accumulators live in shared memory (column `tid` of a [kNdtNV][kNdtBlock] array: conflict-free), which
// frees ~58 registers per thread and lets more warps hide the table / leaf-record latency.
struct SmemAcc {
  double* p;
  __device__ __forceinline__ void add(int k, double v) { p[k * kNdtBlock] += v; }
};

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
  acc.add(28, 1.0);
#pragma unroll
  for (int c = 0; c < 6; c++) acc.add(1 + c, double(e * xCJ[c]));
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
  }
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
  acc.add(28, 1.0);
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
}

template <int SEARCH, bool DOUBLE_PATH>
__global__ void __launch_bounds__(kNdtBlock, DOUBLE_PATH ? 2 : 5)
ndt_eval_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ offs, NdtTargetView tgt,
                const NdtEvalParams* __restrict__ params, NdtEvalResult* __restrict__ results, double* __restrict__ partials,
                unsigned* __restrict__ tickets, int max_blocks, int req_base) {
  constexpr int NNB = NbTraits<SEARCH>::N;
  const int req = req_base + blockIdx.y;
  __shared__ NdtEvalParams sp;
  __shared__ double sacc[kNdtNV * kNdtBlock];
  __shared__ int s_last;
  {
    const int nwords = sizeof(NdtEvalParams) / 4;
    const int* gp = reinterpret_cast<const int*>(params + req);
    int* spw = reinterpret_cast<int*>(&sp);
    for (int k = threadIdx.x; k < nwords; k += kNdtBlock) spw[k] = gp[k];
  }
  __syncthreads();
  const uint32_t begin = offs[sp.scan], end = offs[sp.scan + 1];
  const int nb = min(int((end - begin + kNdtBlock - 1) / kNdtBlock), max_blocks);  // blocks working on this request
  if (int(blockIdx.x) >= nb) return;

  SmemAcc acc{sacc + threadIdx.x};
#pragma unroll
  for (int k = 0; k < kNdtNV; k++) sacc[k * kNdtBlock + threadIdx.x] = 0.0;
  const GridSpec& g = tgt.g;
  const bool hess = sp.compute_hessian != 0;

  for (uint32_t i = begin + blockIdx.x * kNdtBlock + threadIdx.x; i < end; i += uint32_t(nb) * kNdtBlock) {
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
    // ---- phase 1: all neighbourhood table lookups in flight at once
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
      // ---- phase 2: neighbours in reference order, next leaf record prefetched while the current one is evaluated
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

  // fixed-order block reduction straight out of shared memory: warp w owns components w, w+4, ...
  __syncthreads();
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < kNdtNV; k += kNdtBlock / 32) {
      const double* col = sacc + k * kNdtBlock;
      double v = ((col[lane] + col[lane + 32]) + col[lane + 64]) + col[lane + 96];
      v = warp_sum(v);
      if (lane == 0) partials[(size_t(req) * max_blocks + blockIdx.x) * kNdtNV + k] = v;
    }
    if (lane == 0) __threadfence();  // one fence per writer warp, after all of its partial sums
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(tickets + req, 1u);
    s_last = (t == unsigned(nb - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < kNdtNV) {
    const double* base = partials + size_t(req) * max_blocks * kNdtNV + threadIdx.x;
    double tsum = 0.0;
    for (int b = 0; b < nb; b++) tsum += __ldcg(base + size_t(b) * kNdtNV);
    results[req].v[threadIdx.x] = tsum;  // host-mapped pinned memory
  }
  if (threadIdx.x == 0) tickets[req] = 0;
}

NdtDriver::~NdtDriver() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (ready) cudaEventDestroy(ready);
  if (second_stream) { cudaStreamSynchronize(second_stream); cudaStreamDestroy(second_stream); }
  if (mapped_results) cudaFreeHost(mapped_results);
}

template <int SEARCH>
static void launch_eval(bool double_path, dim3 grid, cudaStream_t s, const float4* src, const uint32_t* offs, const NdtTargetView& v,
                        const NdtEvalParams* params, NdtEvalResult* results, double* partials, unsigned* tickets, int max_blocks, int base) {
  if (double_path)
    ndt_eval_kernel<SEARCH, true><<<grid, kNdtBlock, 0, s>>>(src, offs, v, params, results, partials, tickets, max_blocks, base);
  else
    ndt_eval_kernel<SEARCH, false><<<grid, kNdtBlock, 0, s>>>(src, offs, v, params, results, partials, tickets, max_blocks, base);
}

// Requests h_params[0..count) must be ordered: all float-path (kind 0) requests first, then the double-path (kind 1) ones.
void NdtDriver::evaluate(const float4* src, const uint32_t* d_offs, size_t max_pts, const NdtTarget& tgt, int search, int count,
                         bool profile, cudaStream_t s) {
  launch(src, d_offs, max_pts, tgt, search, count, profile, s);
  collect(count, profile, s);
}

void NdtDriver::launch(const float4* src, const uint32_t* d_offs, size_t max_pts, const NdtTarget& tgt, int search, int count, bool profile,
                       cudaStream_t s) {
  pending_run = false;
  if (count == 0) return;
  int n0 = 0;
  while (n0 < count && h_params.p[n0].kind == 0) n0++;
  // enough blocks to fill the machine a few times over, never more than one block per 128 points
  const int full_blocks = int((max_pts + kNdtBlock - 1) / kNdtBlock);
  // 5 blocks of 128 threads are resident per SM: one resident wave, points strided over it
  // (floor, not ceil: every block does the same strided share of its request, so one block too many means a second wave)
  int max_blocks = std::max(1, std::min(full_blocks, (kNumSMs * 5) / count));
  d_params.ensure(count);
  partials.ensure(size_t(count) * max_blocks * kNdtNV);
  if (tickets.cap < size_t(count)) {
    tickets.ensure(count);
    PCR_CUDA_CHECK(cudaMemsetAsync(tickets.p, 0, tickets.cap * sizeof(unsigned), s));
  }
  if (mapped_cap < size_t(count)) {
    if (mapped_results) cudaFreeHost(mapped_results);
    mapped_cap = size_t(count) + 64;
    PCR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&mapped_results), mapped_cap * sizeof(NdtEvalResult), cudaHostAllocMapped));
  }
  memset(mapped_results, 0, size_t(count) * sizeof(NdtEvalResult));  // requests over empty scans launch no block
  NdtEvalResult* dev_results = nullptr;
  PCR_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev_results), mapped_results, 0));
  PCR_CUDA_CHECK(cudaMemcpyAsync(d_params.p, h_params.p, size_t(count) * sizeof(NdtEvalParams), cudaMemcpyHostToDevice, s));
  const bool run = !tgt.overflow && tgt.nleaves > 0 && max_pts > 0;
  if (run) {
    for (int q = 0; q < count; q++) point_evals += h_offsets.p[h_params.p[q].scan + 1] - h_offsets.p[h_params.p[q].scan];
    NdtTargetView v;
    v.recs = tgt.recs.p; v.mean = tgt.mean.p; v.icov = tgt.icov.p; v.table = tgt.table.p; v.g = tgt.g;
    v.centroids = tgt.centroids.p;
    v.d1 = tgt.d1; v.d2 = tgt.d2; v.d2f = float(tgt.d2);
    v.radius2 = float(double(tgt.resolution) * double(tgt.resolution));  // KdTreeFLANN::radiusSearch: (float)(radius * radius)
    if (profile) {
      if (!ev0) { PCR_CUDA_CHECK(cudaEventCreate(&ev0)); PCR_CUDA_CHECK(cudaEventCreate(&ev1)); }
      PCR_CUDA_CHECK(cudaEventRecord(ev0, s));
    }
    for (int part = 0; part < 2; part++) {
      const int base = part == 0 ? 0 : n0, n = part == 0 ? n0 : count - n0;
      if (n == 0) continue;
      dim3 grid(max_blocks, n);
      switch (search) {
        case PCR_NDT_DIRECT1: launch_eval<PCR_NDT_DIRECT1>(part == 1, grid, s, src, d_offs, v, d_params.p, dev_results, partials.p, tickets.p, max_blocks, base); break;
        case PCR_NDT_DIRECT26: launch_eval<PCR_NDT_DIRECT26>(part == 1, grid, s, src, d_offs, v, d_params.p, dev_results, partials.p, tickets.p, max_blocks, base); break;
        case PCR_NDT_KDTREE: launch_eval<PCR_NDT_KDTREE>(part == 1, grid, s, src, d_offs, v, d_params.p, dev_results, partials.p, tickets.p, max_blocks, base); break;
        default: launch_eval<PCR_NDT_DIRECT7>(part == 1, grid, s, src, d_offs, v, d_params.p, dev_results, partials.p, tickets.p, max_blocks, base); break;
      }
      launches++;
      hot_launches++;
    }
    if (profile) PCR_CUDA_CHECK(cudaEventRecord(ev1, s));
  }
  pending_run = run;
}

void NdtDriver::collect(int count, bool profile, cudaStream_t s) {
  if (count == 0) return;
  const bool run = pending_run;
  PCR_CUDA_CHECK(cudaStreamSynchronize(s));
  PCR_CUDA_CHECK(cudaGetLastError());
  h_results.ensure(count);
  memcpy(h_results.p, mapped_results, size_t(count) * sizeof(NdtEvalResult));
  if (profile && run) {
    float ms = 0.f;
    PCR_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    hot_ms += ms;
  }
}

// ================================================================================================================
// N5. Newton + More-Thuente driver (ndt_omp_impl.hpp:81-171, 773-932) as a per-scan state machine, so that a batch
// of independent scans advances in lock-step with ONE kernel launch per round.
// ================================================================================================================
namespace {
struct ScanState {
  enum Phase { INIT_EVAL, LS_FIRST, LS_LOOP, LS_HESSIAN, FINISHED } phase = INIT_EVAL;
  double p[6];
  double score = 0;
  double g[6];
  double H[36];
  float final_T[16];
  float eval_T[16];
  double eval_p[6];
  int eval_kind = 0, eval_hess = 1;
  int nr_iterations = 0;
  bool converged = false;
  // line search
  double step_dir[6], phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, x_t[6], phi_t, d_phi_t, psi_t, d_psi_t;
  bool interval_converged, open_interval;
  int step_iterations;
  int n_evals = 0, n_hess = 0;
  long long n_pairs = 0;
};

struct NdtLogic {
  const pcr_params& prm;
  static constexpr double mu = 1.e-4, nu = 0.9;
  static constexpr int max_step_iterations = 10;

  void request_deriv(ScanState& st, const float* T, const double* p, bool hess) {
    std::memcpy(st.eval_T, T, sizeof(float) * 16);
    std::memcpy(st.eval_p, p, sizeof(double) * 6);
    st.eval_kind = 0;
    st.eval_hess = hess ? 1 : 0;
  }
  void start(ScanState& st, const double* Tguess) {
    float guess[16];
    bool ident = true;
    for (int i = 0; i < 16; i++) {
      guess[i] = static_cast<float>(Tguess[i]);  // NdtRegister.cpp:27 res.matrix().cast<float>()
      if (guess[i] != ((i % 5 == 0) ? 1.f : 0.f)) ident = false;
    }
    for (int i = 0; i < 16; i++) st.final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
    if (!ident) std::memcpy(st.final_T, guess, sizeof(guess));  // ndt_omp_impl.hpp:95-101
    float Rm[9], eul[3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Rm[r * 3 + c] = st.final_T[c * 4 + r];
    hm::euler_xyz_f32(Rm, eul);  // :103-111
    st.p[0] = st.final_T[12]; st.p[1] = st.final_T[13]; st.p[2] = st.final_T[14];
    st.p[3] = eul[0]; st.p[4] = eul[1]; st.p[5] = eul[2];
    st.nr_iterations = 0;
    st.converged = false;
    st.phase = ScanState::INIT_EVAL;
    request_deriv(st, st.final_T, st.p, true);
  }
  void take(ScanState& st, const NdtEvalResult& r, bool with_hessian) {
    st.score = r.v[0];
    for (int i = 0; i < 6; i++) st.g[i] = r.v[1 + i];
    int k = 7;
    for (int a = 0; a < 6; a++)
      for (int b = a; b < 6; b++) { st.H[a * 6 + b] = with_hessian ? r.v[k] : 0.0; st.H[b * 6 + a] = st.H[a * 6 + b]; k++; }
  }
  void set_trial(ScanState& st) {
    st.a_t = std::min(st.a_t, prm.ndt_step_size);
    st.a_t = std::max(st.a_t, prm.ndt_trans_eps / 2);
    for (int i = 0; i < 6; i++) st.x_t[i] = st.p[i] + st.step_dir[i] * st.a_t;
    hm::ndt_pose_matrix_f32(st.x_t, st.final_T);
  }
  void begin_outer(ScanState& st) {
    double b[6], delta_p[6];
    for (int i = 0; i < 6; i++) b[i] = -st.g[i];
    hm::svd6_solve(st.H, b, delta_p);  // :127-129
    double nrm = 0;
    for (int i = 0; i < 6; i++) nrm += delta_p[i] * delta_p[i];
    nrm = std::sqrt(nrm);
    if (nrm == 0 || nrm != nrm) {  // :134-139
      st.converged = nrm == nrm;
      st.phase = ScanState::FINISHED;
      return;
    }
    for (int i = 0; i < 6; i++) st.step_dir[i] = delta_p[i] / nrm;
    // computeStepLengthMT :773-
    st.phi_0 = -st.score;
    double d = 0;
    for (int i = 0; i < 6; i++) d += st.g[i] * st.step_dir[i];
    st.d_phi_0 = -d;
    if (st.d_phi_0 >= 0) {
      if (st.d_phi_0 == 0) { end_outer(st, 0.0); return; }
      st.d_phi_0 *= -1;
      for (int i = 0; i < 6; i++) st.step_dir[i] *= -1;
    }
    st.step_iterations = 0;
    st.a_l = 0; st.a_u = 0;
    st.f_l = hm::mt_psi(st.a_l, st.phi_0, st.phi_0, st.d_phi_0, mu);
    st.g_l = hm::mt_dpsi(st.d_phi_0, st.d_phi_0, mu);
    st.f_u = hm::mt_psi(st.a_u, st.phi_0, st.phi_0, st.d_phi_0, mu);
    st.g_u = hm::mt_dpsi(st.d_phi_0, st.d_phi_0, mu);
    st.interval_converged = (prm.ndt_step_size - prm.ndt_trans_eps / 2) < 0;
    st.open_interval = true;
    st.a_t = nrm;
    set_trial(st);
    st.phase = ScanState::LS_FIRST;
    request_deriv(st, st.final_T, st.x_t, true);
  }
  void after_eval(ScanState& st) {
    st.phi_t = -st.score;
    double d = 0;
    for (int i = 0; i < 6; i++) d += st.g[i] * st.step_dir[i];
    st.d_phi_t = -d;
    st.psi_t = hm::mt_psi(st.a_t, st.phi_t, st.phi_0, st.d_phi_0, mu);
    st.d_psi_t = hm::mt_dpsi(st.d_phi_t, st.d_phi_0, mu);
  }
  void ls_continue(ScanState& st) {
    if (!st.interval_converged && st.step_iterations < max_step_iterations && !(st.psi_t <= 0 && st.d_phi_t <= -nu * st.d_phi_0)) {
      if (st.open_interval) st.a_t = hm::mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.psi_t, st.d_psi_t);
      else st.a_t = hm::mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.phi_t, st.d_phi_t);
      set_trial(st);
      st.phase = ScanState::LS_LOOP;
      request_deriv(st, st.final_T, st.x_t, false);
      return;
    }
    if (st.step_iterations) {  // :928-929 computeHessian
      st.phase = ScanState::LS_HESSIAN;
      std::memcpy(st.eval_T, st.final_T, sizeof(float) * 16);
      std::memcpy(st.eval_p, st.x_t, sizeof(double) * 6);
      st.eval_kind = 1;
      st.eval_hess = 1;
      return;
    }
    end_outer(st, st.a_t);
  }
  void end_outer(ScanState& st, double a) {
    for (int i = 0; i < 6; i++) st.p[i] += st.step_dir[i] * a;
    if (st.nr_iterations > prm.ndt_max_iters || (st.nr_iterations && (std::fabs(a) < prm.ndt_trans_eps))) st.converged = true;  // :158-162
    st.nr_iterations++;
    if (st.converged) { st.phase = ScanState::FINISHED; return; }
    begin_outer(st);
  }
  void on_result(ScanState& st, const NdtEvalResult& r) {
    st.n_pairs += (long long)(r.v[28] + 0.5);
    switch (st.phase) {
      case ScanState::INIT_EVAL:
        st.n_evals++;
        take(st, r, true);
        begin_outer(st);
        break;
      case ScanState::LS_FIRST:
        st.n_evals++;
        take(st, r, true);
        after_eval(st);
        ls_continue(st);
        break;
      case ScanState::LS_LOOP: {
        st.n_evals++;
        take(st, r, false);
        after_eval(st);
        if (st.open_interval && (st.psi_t <= 0 && st.d_psi_t >= 0)) {
          st.open_interval = false;
          st.f_l = st.f_l + st.phi_0 - mu * st.d_phi_0 * st.a_l;
          st.g_l = st.g_l + mu * st.d_phi_0;
          st.f_u = st.f_u + st.phi_0 - mu * st.d_phi_0 * st.a_u;
          st.g_u = st.g_u + mu * st.d_phi_0;
        }
        if (st.open_interval)
          st.interval_converged = hm::mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.psi_t, st.d_psi_t);
        else
          st.interval_converged = hm::mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.phi_t, st.d_phi_t);
        st.step_iterations++;
        ls_continue(st);
        break;
      }
      case ScanState::LS_HESSIAN: {
        st.n_hess++;
        int k = 7;
        for (int a = 0; a < 6; a++)
          for (int b = a; b < 6; b++) { st.H[a * 6 + b] = r.v[k]; st.H[b * 6 + a] = r.v[k]; k++; }
        end_outer(st, st.a_t);
        break;
      }
      default: break;
    }
  }
};
}  // namespace

static void fill_params(NdtEvalParams& ep, const float* T, const double* p, int kind, int hess, int scan) {
  std::memcpy(ep.Tf, T, sizeof(float) * 16);
  hm::ndt_angle_tables(p, ep.j_ang, ep.h_ang, ep.j_ang_d, ep.h_ang_d);
  ep.compute_hessian = hess;
  ep.kind = kind;
  ep.scan = scan;
  ep.pad = 0;
}

// One lane = one NdtDriver + one stream driving the state machines of a subset of the scans in lock-step.
namespace {
struct Lane {
  NdtDriver* drv;
  cudaStream_t stream;
  std::vector<int> scans;    // scans of the batch owned by this lane
  std::vector<int> active;   // scans with a request in flight (order of h_params)
  bool in_flight = false;
};

// builds the lane's next round of requests; returns false when all of its scans are finished
bool lane_launch(Lane& L, std::vector<ScanState>& st, const float4* src, size_t max_pts, const NdtTarget& tgt, const pcr_params& prm, bool profile) {
  L.active.clear();
  for (int i : L.scans)
    if (st[size_t(i)].phase != ScanState::FINISHED) L.active.push_back(i);
  L.in_flight = false;
  if (L.active.empty()) return false;
  std::stable_partition(L.active.begin(), L.active.end(), [&](int i) { return st[size_t(i)].eval_kind == 0; });
  L.drv->h_params.ensure(L.active.size());
  for (size_t k = 0; k < L.active.size(); k++) {
    ScanState& ss = st[size_t(L.active[k])];
    fill_params(L.drv->h_params.p[k], ss.eval_T, ss.eval_p, ss.eval_kind, ss.eval_hess, L.active[k]);
  }
  L.drv->launch(src, L.drv->offsets.p, max_pts, tgt, prm.ndt_search, int(L.active.size()), profile, L.stream);
  L.in_flight = true;
  return true;
}
}  // namespace

int NdtDriver::align(const float4* src, const size_t* offs, size_t n_scans, const NdtTarget& tgt, const pcr_params& prm, double* T,
                     int32_t* converged, int32_t* iters, double* trans_prob, bool profile, cudaStream_t s) {
  launches = 0; hot_ms = 0.f; hot_launches = 0; total_evals = 0; total_hess = 0; total_pairs = 0; point_evals = 0;
  if (n_scans == 0) return 0;
  size_t max_pts = 0;
  for (size_t i = 0; i < n_scans; i++) max_pts = std::max(max_pts, size_t(offs[i + 1] - offs[i]));
  std::vector<ScanState> st(n_scans);
  NdtLogic logic{prm};
  for (size_t i = 0; i < n_scans; i++) logic.start(st[i], T + i * 16);

  // Batches run as two lanes on two streams: while one lane's kernel runs, the host digests the other lane's results
  // (Newton direction, More-Thuente bookkeeping, angle tables) and queues its next round, so the GPU does not idle
  // between evaluation rounds. A single scan uses one lane.
  const char* lanes_env = std::getenv("PCR_NDT_LANES");  // tuning / measurement knob: 1 = single lane (clean per-kernel timing)
  const bool two = n_scans >= 4 && !(lanes_env && std::atoi(lanes_env) == 1);
  if (two && !second) {
    second.reset(new NdtDriver());
    PCR_CUDA_CHECK(cudaStreamCreateWithFlags(&second_stream, cudaStreamNonBlocking));
    PCR_CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  }
  Lane lanes[2];
  const int nl = two ? 2 : 1;
  lanes[0].drv = this; lanes[0].stream = s;
  if (two) {
    lanes[1].drv = second.get(); lanes[1].stream = second_stream;
    second->launches = 0; second->hot_ms = 0.f; second->hot_launches = 0; second->point_evals = 0;
    PCR_CUDA_CHECK(cudaEventRecord(ready, s));                      // src was packed on the main stream
    PCR_CUDA_CHECK(cudaStreamWaitEvent(second_stream, ready, 0));
  }
  for (size_t i = 0; i < n_scans; i++) lanes[two ? (i * 2 / n_scans) : 0].scans.push_back(int(i));
  for (int l = 0; l < nl; l++) {
    NdtDriver* d = lanes[l].drv;
    uint32_t* ho = d->h_offsets.ensure(n_scans + 1);
    for (size_t i = 0; i <= n_scans; i++) ho[i] = uint32_t(offs[i] - offs[0]);
    d->offsets.ensure(n_scans + 1);
    PCR_CUDA_CHECK(cudaMemcpyAsync(d->offsets.p, ho, (n_scans + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, lanes[l].stream));
  }
  for (int l = 0; l < nl; l++) lane_launch(lanes[l], st, src, max_pts, tgt, prm, profile);
  for (;;) {
    bool any = false;
    for (int l = 0; l < nl; l++) {
      Lane& L = lanes[l];
      if (!L.in_flight) continue;
      any = true;
      L.drv->collect(int(L.active.size()), profile, L.stream);
      for (size_t k = 0; k < L.active.size(); k++) logic.on_result(st[size_t(L.active[k])], L.drv->h_results.p[k]);
      lane_launch(L, st, src, max_pts, tgt, prm, profile);
    }
    if (!any) break;
  }
  if (two) {
    launches += second->launches; hot_ms += second->hot_ms; hot_launches += second->hot_launches; point_evals += second->point_evals;
  }
  for (size_t i = 0; i < n_scans; i++) {
    for (int q = 0; q < 16; q++) T[i * 16 + q] = double(st[i].final_T[q]);  // NdtRegister.cpp:28
    if (converged) converged[i] = st[i].converged ? 1 : 0;
    if (iters) iters[i] = st[i].nr_iterations;
    const size_t ns = offs[i + 1] - offs[i];
    if (trans_prob) trans_prob[i] = st[i].score / double(ns);
    total_evals += st[i].n_evals;
    total_hess += st[i].n_hess;
    total_pairs += st[i].n_pairs;
  }
  return 0;
}

}  // namespace pcr
