// Exact k-nearest-neighbour search (k <= 32) for density-skewed clouds (raw 64/128-beam scans: thousands of points per
// 0.5 m cell next to the sensor, a handful 100 m away): a nested multi-resolution grid. Points are sorted by a 63-bit
// Morton code of their finest-level cell (1/32 m), so that a cell of ANY of the 9 levels (1/32 m ... 8 m, factor 2) is one
// contiguous run of the sorted array; per level an open-addressing hash maps cell code -> [start, end).
// A query picks the finest level whose own cell holds enough points, searches the 27 cells of that level nearest-first
// (one warp per query, candidates read coalesced, cells pruned against the current k-th distance) and accepts the result
// when the k-th distance is provably inside the visited cube; otherwise it climbs one level. At the top level the cube
// grows ring by ring. Metric and ordering: FLANN L2_Simple<float> (float diff, float square, float accumulate x->y->z),
// ties by (d2, original index) — what replaces pcl::search::KdTree in FastGICP::calculate_covariances
// (fast_gicp_impl.hpp:241-298) and pcl::Registration::getFitnessScore.
#pragma once
#include "common.cuh"
#include "voxel.cuh"

namespace pcr {

constexpr int kKnnLevels = 9;
constexpr int kKnnMaxK = 32;

struct MortonGrid {
  float mn[3] = {0, 0, 0};
  float h0 = 1.0f / 32.0f, inv_h0 = 32.0f;
  int dim0[3] = {1, 1, 1};   // cells per axis at level 0
  float smax[3] = {0, 0, 0}; // largest scaled coordinate still inside the last cell
  double slack0 = 1e-3;      // float rounding of the cell assignment, in level-0 cells
  uint32_t cap = 0;          // slots per level (power of two)
  size_t n = 0;
  bool built = false;
  DevBuf<float4> pts;                // Morton-sorted, w = original index bits
  DevBuf<uint4> tables;              // kKnnLevels x cap slots {key_lo, key_hi, start, end}; key = cell code + 1, 0 = empty
  DevBuf<unsigned long long> c0, c1; // codes (unsorted / sorted)
  DevBuf<uint32_t> v0, v1;
  DevBuf<unsigned char> tmp;
};

struct MortonView {
  const float4* pts;
  const uint4* tables;
  float mn[3];
  float inv_h0, h0;
  int dim0[3];
  float smax[3];
  float slack0;
  uint32_t cap_mask, cap;
  int n;
};

inline MortonView view_of(const MortonGrid& g) {
  MortonView v;
  v.pts = g.pts.p; v.tables = g.tables.p;
  for (int a = 0; a < 3; a++) { v.mn[a] = g.mn[a]; v.dim0[a] = g.dim0[a]; v.smax[a] = g.smax[a]; }
  v.inv_h0 = g.inv_h0; v.h0 = g.h0; v.slack0 = float(g.slack0);
  v.cap = g.cap; v.cap_mask = g.cap - 1; v.n = int(g.n);
  return v;
}

// Returns a PCR error code (0 ok, -5 extent too large for 21 bits per axis).
int build_morton_grid(const float4* pts, size_t n, MortonGrid& grid, BBoxWork& bw, cudaStream_t s);

#ifdef __CUDACC__
__host__ __device__ __forceinline__ unsigned long long morton_spread(unsigned v) {  // 21 bits -> every third bit
  unsigned long long x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__host__ __device__ __forceinline__ unsigned long long morton3(unsigned x, unsigned y, unsigned z) {
  return morton_spread(x) | (morton_spread(y) << 1) | (morton_spread(z) << 2);
}
__device__ __forceinline__ uint32_t knn_hash(unsigned long long key) {
  key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ull; key ^= key >> 33;
  return uint32_t(key);
}
// level-0 cell coordinate of a float coordinate: the ONE quantisation used by the build and by every query
__device__ __forceinline__ float knn_scaled(float x, float mn, float inv_h0) { return __fmul_rn(__fsub_rn(x, mn), inv_h0); }

// [start, end) of cell (cx, cy, cz) of `level`; empty -> (0, 0)
__device__ __forceinline__ int2 knn_cell(const MortonView& g, int level, int cx, int cy, int cz) {
  const unsigned long long key = morton3(unsigned(cx), unsigned(cy), unsigned(cz)) + 1ull;
  const uint4* tab = g.tables + size_t(level) * g.cap;
  uint32_t h = knn_hash(key) & g.cap_mask;
  for (;;) {
    const uint4 sl = __ldg(tab + h);
    const unsigned long long k = (unsigned long long)sl.x | ((unsigned long long)sl.y << 32);
    if (k == key) return make_int2(int(sl.z), int(sl.w));
    if (k == 0ull) return make_int2(0, 0);
    h = (h + 1) & g.cap_mask;
  }
}

__device__ __forceinline__ float dist2_f32(float qx, float qy, float qz, const float4& m) {
  const float dx = __fsub_rn(qx, m.x), dy = __fsub_rn(qy, m.y), dz = __fsub_rn(qz, m.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

struct WarpKnn {
  float bd;   // this lane's entry of the result (lane < k), +inf when empty. UNSORTED while a search runs: knn_sort_result
  int bi;
  float td;   // current k-th best = the largest kept entry (threshold), +inf while fewer than k are kept; warp-uniform
  int ti;
  int tl;     // lane that holds the threshold entry
  int cnt;    // entries found so far (<= k), warp-uniform
};

constexpr unsigned kFull = 0xffffffffu;

// largest kept entry by (d2, index): two hardware redux.sync max + one ballot (d2 >= 0: float bits order like unsigned)
__device__ __forceinline__ void knn_refresh_threshold(WarpKnn& st, int k, int lane) {
  const unsigned kh = lane < k ? __float_as_uint(st.bd) : 0u;
  const unsigned mh = __reduce_max_sync(kFull, kh);
  const bool e1 = lane < k && kh == mh;
  const unsigned ml = __reduce_max_sync(kFull, e1 ? unsigned(st.bi) : 0u);
  st.tl = __ffs(__ballot_sync(kFull, e1 && unsigned(st.bi) == ml)) - 1;
  st.td = __uint_as_float(mh);
  st.ti = int(ml);
}

// ascending (d2, index) order over the lanes (empty lanes = +inf sort to the end): warp bitonic network, 15 shuffle steps
__device__ __forceinline__ void knn_bitonic_sort(float& d, int& id, int lane) {
#pragma unroll
  for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      const float od = __shfl_xor_sync(kFull, d, jj);
      const int oi = __shfl_xor_sync(kFull, id, jj);
      const bool keep_min = ((lane & jj) == 0) == ((lane & kk) == 0);
      const bool other_less = od < d || (od == d && oi < id);
      const bool other_more = od > d || (od == d && oi > id);
      if (keep_min ? other_less : other_more) { d = od; id = oi; }
    }
  }
}
__device__ __forceinline__ void knn_sort_result(WarpKnn& st, int lane) { knn_bitonic_sort(st.bd, st.bi, lane); }

// scan the contiguous run [lo, hi) of the sorted array: candidates are read coalesced (one float4 per lane). The kept set
// lives UNSORTED in the lanes (lane < k): a candidate that beats the threshold replaces the largest kept entry and the new
// largest is found with redux.sync — ~15 warp instructions per accepted candidate instead of a ballot / shuffle-up sorted
// insertion (~28). The (d2, index) order is total, so the kept SET is the same whatever the arrival order.
// `seen` (level >= 0): the 27 cells around (cx, cy, cz) of that level were examined already by the finer attempt of the same
// search; their points are skipped (they are kept, or were rejected against a bound that has only shrunk since).
struct KnnSeen { int level, cx, cy, cz; };

__device__ __forceinline__ void knn_scan_run(WarpKnn& st, const MortonView& g, int lo, int hi, float qx, float qy, float qz,
                                             int k, int lane, const KnnSeen& seen) {
  const float4* __restrict__ pts = g.pts;
  for (int base = lo; base < hi; base += 32) {
    const int j = base + lane;
    float d2 = 0.f;
    int idx = 0;
    bool pass = false;
    bool valid = j < hi;
    if (valid) {
      const float4 m = __ldg(pts + j);
      d2 = dist2_f32(qx, qy, qz, m);
      idx = __float_as_int(m.w);
      if (seen.level >= 0) {  // the candidate's cell at the finer level, by the build's own quantisation
        const int ax = min(max(int(floorf(knn_scaled(m.x, g.mn[0], g.inv_h0))), 0), g.dim0[0] - 1) >> seen.level;
        const int ay = min(max(int(floorf(knn_scaled(m.y, g.mn[1], g.inv_h0))), 0), g.dim0[1] - 1) >> seen.level;
        const int az = min(max(int(floorf(knn_scaled(m.z, g.mn[2], g.inv_h0))), 0), g.dim0[2] - 1) >> seen.level;
        if (abs(ax - seen.cx) <= 1 && abs(ay - seen.cy) <= 1 && abs(az - seen.cz) <= 1) valid = false;
      }
      pass = valid && (d2 < st.td || (d2 == st.td && idx < st.ti));
    }
    if (st.cnt == 0 && seen.level < 0 && hi - base <= k) {
      // first chunk of a search and it fits: the lanes simply keep their own candidates (the kept set is unsorted anyway)
      const int valid = hi - base;
      st.cnt = valid;
      st.bd = lane < valid ? d2 : INFINITY;
      st.bi = lane < valid ? idx : 0x7fffffff;
      if (valid == k) knn_refresh_threshold(st, k, lane);
      continue;
    }
    if (st.cnt == 0 && seen.level < 0) {
      // first chunk of a search, more candidates than slots: instead of up to 32 serial insertions sort the chunk with a
      // warp bitonic network on (d2, index) and adopt its k smallest
      float d = j < hi ? d2 : INFINITY;
      int id = j < hi ? idx : 0x7fffffff;
      knn_bitonic_sort(d, id, lane);
      const int valid = min(hi - base, 32);
      st.cnt = min(k, valid);
      st.bd = lane < st.cnt ? d : INFINITY;
      st.bi = lane < st.cnt ? id : 0x7fffffff;
      st.td = __shfl_sync(kFull, st.bd, k - 1);   // +inf while fewer than k are kept
      st.ti = __shfl_sync(kFull, st.bi, k - 1);
      st.tl = k - 1;
      continue;
    }
    unsigned mask = __ballot_sync(kFull, pass);
    if (mask == 0u) continue;
    if (st.cnt < k) {
      // empty slots left: lane cnt + r adopts the r-th passing candidate directly
      const int m = min(k - st.cnt, __popc(mask));
      const int want = lane - st.cnt;
      const bool take = want >= 0 && want < m;
      const int srcl = take ? int(__fns(mask, 0, want + 1)) : 0;
      const float sd = __shfl_sync(kFull, d2, srcl);
      const int si = __shfl_sync(kFull, idx, srcl);
      if (take) { st.bd = sd; st.bi = si; }
      st.cnt += m;
      const int rank = __popc(mask & ((1u << lane) - 1u));
      mask = __ballot_sync(kFull, pass && rank >= m);  // the candidates that did not find an empty slot
      if (st.cnt == k) knn_refresh_threshold(st, k, lane);
    }
    while (mask) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1;
      const float cd = __shfl_sync(kFull, d2, s);
      const int ci = __shfl_sync(kFull, idx, s);
      if (!(cd < st.td || (cd == st.td && ci < st.ti))) continue;  // the threshold moved since the ballot
      if (lane == st.tl) { st.bd = cd; st.bi = ci; }
      knn_refresh_threshold(st, k, lane);
    }
  }
}

__device__ __forceinline__ void knn_reset(WarpKnn& st) {
  st.bd = INFINITY; st.bi = 0x7fffffff; st.td = INFINITY; st.ti = 0x7fffffff; st.tl = 0; st.cnt = 0;
}

// the 27 cells of a level, nearest first: centre, 6 faces, 12 edges, 8 corners; packed (dx+1) | (dy+1)<<2 | (dz+1)<<4
static __constant__ unsigned char c_knn_nb[27] = {
    21, 22, 20, 25, 17, 37, 5,                              // centre (1,1,1); faces
    26, 24, 18, 16, 38, 36, 6, 4, 41, 33, 9, 1,             // edges
    42, 40, 34, 32, 10, 8, 2, 0};                           // corners

// Warp-cooperative exact k-NN. `min_pop`: own-cell population that selects the starting level.
// On return lanes [0, k) hold the cnt neighbours UNSORTED (empty lanes +inf); knn_sort_result orders them by (d2, idx).
__device__ __forceinline__ WarpKnn knn_warp_morton(const MortonView& g, float qx, float qy, float qz, int k, int min_pop, int lane) {
  WarpKnn st;
  knn_reset(st);
  KnnSeen seen{-1, 0, 0, 0};
  const KnnSeen none{-1, 0, 0, 0};
  // Level-0 scaled coordinates of the query. A query outside the cloud's bounding box is searched from its projection q'
  // onto the box: for every point p of the (convex) box |p - q|^2 >= |p - q'|^2 + |q' - q|^2, so all "everything outside
  // the visited cube is farther than ..." bounds are taken around q' and get extra2 = |q' - q|^2 added. Distances
  // themselves are always measured to the real query.
  const float q[3] = {qx, qy, qz};
  float s0[3];
  int i0[3];
  float extra2 = 0.f;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float sr = knn_scaled(q[a], g.mn[a], g.inv_h0);
    s0[a] = fminf(fmaxf(sr, 0.f), g.smax[a]);
    const float off = fminf(fabsf(sr - s0[a]), 1.0e9f) * g.h0;
    extra2 += off * off;
    i0[a] = int(floorf(s0[a]));
  }
  extra2 *= 0.9999f;
  const bool inside0 = true;
  // ---- starting level: finest level whose own cell holds at least min_pop points (lane l probes level l)
  int level = kKnnLevels - 1;
  {
    int pop = 0;
    if (lane < kKnnLevels && inside0) {
      const int2 r = knn_cell(g, lane, i0[0] >> lane, i0[1] >> lane, i0[2] >> lane);
      pop = r.y - r.x;
    }
    const unsigned ok = __ballot_sync(kFull, pop >= min_pop) & ((1u << kKnnLevels) - 1u);
    if (ok) level = __ffs(ok) - 1;
  }
  for (;;) {
    const float scale = 1.0f / float(1 << level);
    const float h = g.h0 * float(1 << level);
    const float slack = g.slack0 * scale + 1e-6f;
    // cell and in-cell position at this level (floor division also for negative level-0 indices)
    int c[3];
    float fr[3];
    float margin = 1.f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      c[a] = i0[a] >> level;
      fr[a] = (s0[a] - float(c[a]) * float(1 << level)) * scale;
      margin = fminf(margin, fminf(fr[a], 1.f - fr[a]));
    }
    // ---- 27 cells of this level: lane j probes the j-th nearest, all probes in flight together
    int lo = 0, hi = 0;
    float cmin2 = 0.f;
    if (lane < 27) {
      const unsigned pk = c_knn_nb[lane];
      const int dx = int(pk & 3u) - 1, dy = int((pk >> 2) & 3u) - 1, dz = int((pk >> 4) & 3u) - 1;
      const int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
      if (x >= 0 && y >= 0 && z >= 0 && x <= (g.dim0[0] >> level) && y <= (g.dim0[1] >> level) && z <= (g.dim0[2] >> level)) {
        const int2 r = knn_cell(g, level, x, y, z);
        lo = r.x; hi = r.y;
      }
      const float ax = fmaxf((dx == 0 ? 0.f : (dx > 0 ? 1.f - fr[0] : fr[0])) - slack, 0.f);
      const float ay = fmaxf((dy == 0 ? 0.f : (dy > 0 ? 1.f - fr[1] : fr[1])) - slack, 0.f);
      const float az = fmaxf((dz == 0 ? 0.f : (dz > 0 ? 1.f - fr[2] : fr[2])) - slack, 0.f);
      cmin2 = (ax * ax + ay * ay + az * az) * h * h * 0.99999f + extra2;  // every point of the cell is at least this far
    }
    unsigned cells = __ballot_sync(kFull, hi > lo);
    while (cells) {
      const int sl = __ffs(cells) - 1;
      cells &= cells - 1;
      const int rlo = __shfl_sync(kFull, lo, sl), rhi = __shfl_sync(kFull, hi, sl);
      const float cm = __shfl_sync(kFull, cmin2, sl);
      if (st.cnt == k && cm > st.td) continue;  // cannot hold anything better than the current k-th
      knn_scan_run(st, g, rlo, rhi, qx, qy, qz, k, lane, seen);
    }
    // every point outside the 27 cells is farther than reach
    const float reach = (1.f + margin - slack) * h;
    if (st.cnt == k && reach > 0.f && st.td < reach * reach * 0.99999f + extra2) return st;
    if (level < kKnnLevels - 1) {
      // climb one level (the coarser cube contains the finer one) and KEEP what was found: the points of the cube examined so
      // far are skipped at the next level, everything else competes against the bound already reached
      seen.level = level; seen.cx = c[0]; seen.cy = c[1]; seen.cz = c[2];
      level++;
      continue;
    }
    // ---- top level: grow the cube ring by ring (Chebyshev shells) until the k-th distance is inside it or the grid is exhausted
    const int dimx = (g.dim0[0] >> level) + 1, dimy = (g.dim0[1] >> level) + 1, dimz = (g.dim0[2] >> level) + 1;
    int rmax = 0;
    rmax = max(rmax, max(c[0], dimx - 1 - c[0]));
    rmax = max(rmax, max(c[1], dimy - 1 - c[1]));
    rmax = max(rmax, max(c[2], dimz - 1 - c[2]));
    rmax = min(rmax, 1 << 20);
    for (int r = 2; r <= rmax; r++) {
      const int side = 2 * r + 1;
      const int slab = side * side, per = 4 * side - 4;
      const int nshell = 2 * slab + (side - 2) * per;  // = side^3 - (side-2)^3
      for (int base = 0; base < nshell; base += 32) {
        const int e = base + lane;
        int l2 = 0, h2 = 0;
        if (e < nshell) {
          int dx, dy, dz;
          if (e < 2 * slab) {  // bottom / top z-slabs
            const int rem = e % slab;
            dz = (e / slab) ? r : -r;
            dy = rem / side - r;
            dx = rem % side - r;
          } else {  // perimeter of the middle layers
            const int e2 = e - 2 * slab, p = e2 % per;
            dz = -r + 1 + e2 / per;
            if (p < side) { dy = -r; dx = p - r; }
            else if (p < 2 * side) { dy = r; dx = p - side - r; }
            else { const int qq = p - 2 * side; dy = -r + 1 + (qq >> 1); dx = (qq & 1) ? r : -r; }
          }
          const int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
          if (x >= 0 && y >= 0 && z >= 0 && x < dimx && y < dimy && z < dimz) {
            const int2 rr = knn_cell(g, level, x, y, z);
            l2 = rr.x; h2 = rr.y;
          }
        }
        unsigned cl = __ballot_sync(kFull, h2 > l2);
        while (cl) {
          const int sl = __ffs(cl) - 1;
          cl &= cl - 1;
          knn_scan_run(st, g, __shfl_sync(kFull, l2, sl), __shfl_sync(kFull, h2, sl), qx, qy, qz, k, lane, none);
        }
      }
      if (st.cnt == k) {
        const float rch = (float(r) + margin - slack) * h;
        if (rch > 0.f && st.td < rch * rch * 0.99999f + extra2) break;
      }
    }
    return st;
  }
}
#endif  // __CUDACC__

}  // namespace pcr
