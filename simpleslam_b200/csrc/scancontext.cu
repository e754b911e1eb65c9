// ScanContext descriptor and descriptor distance (scancontext.cuh). One block per cloud / per descriptor pair.
#include "scancontext.cuh"
#include <cfloat>

namespace pcr {

// order-preserving int encoding of a float (atomicMax on heights)
__device__ __forceinline__ int sc_enc(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float sc_dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) sc_make_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ offs, float lidar_height,
                                                      double* __restrict__ desc, double* __restrict__ ring_key, double* __restrict__ sector_key) {
  __shared__ int bins[kScSize];
  __shared__ double vals[kScSize];
  const int cloud = blockIdx.x;
  const int no_point = sc_enc(-1000.0f);  // NO_POINT (:159)
  for (int b = threadIdx.x; b < kScSize; b += blockDim.x) bins[b] = no_point;
  __syncthreads();
  const uint32_t begin = offs[cloud], end = offs[cloud + 1];
  for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const float4 p = __ldg(pts + i);
    const float z = __fadd_rn(p.z, lidar_height);                                           // :169
    const float range = __fsqrt_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)));   // :172
    // xy2theta (:28-33): float res = atan2f(y, x) + M_PI (double sum, rounded to float); clamp; rad2deg in double -> float.
    // atan2 is evaluated in double and rounded: the correctly rounded float arctangent.
    float res = float(double(float(atan2(double(p.y), double(p.x)))) + 3.14159265358979323846);
    res = fmaxf(0.0f, fminf(float(2.0 * 3.14159265358979323846), res));
    const float angle = float(double(res) * 180.0 / 3.14159265358979323846);
    if (range > kScMaxRadius) continue;                                                     // :176-177
    const int ring = max(min(kScRings, int(ceilf(__fmul_rn(__fdiv_rn(range, kScMaxRadius), float(kScRings))))), 1);
    const int sector = max(min(kScSectors, int(ceil((double(angle) / 360.0) * double(kScSectors)))), 1);
    atomicMax(&bins[(ring - 1) * kScSectors + (sector - 1)], sc_enc(z));                    // :183-184 maximum z of the bin
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kScSize; b += blockDim.x) {
    const double v = bins[b] == no_point ? 0.0 : double(sc_dec(bins[b]));                   // :188-191
    vals[b] = v;
    desc[size_t(cloud) * kScSize + b] = v;
  }
  __syncthreads();
  // ring key = row means (:200-212), sector key = column means (:215-230); sums of float-valued doubles are exact here
  if (threadIdx.x < kScRings) {
    double sum = 0.0;
    for (int c = 0; c < kScSectors; c++) sum += vals[threadIdx.x * kScSectors + c];
    ring_key[size_t(cloud) * kScRings + threadIdx.x] = sum / double(kScSectors);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kScSectors) {
    const int c = threadIdx.x - 64;
    double sum = 0.0;
    for (int r = 0; r < kScRings; r++) sum += vals[r * kScSectors + c];
    sector_key[size_t(cloud) * kScSectors + c] = sum / double(kScRings);
  }
}

void scancontext_make(const float4* pts, const uint32_t* d_offs, int n_clouds, float lidar_height, double* desc, double* ring_key,
                      double* sector_key, cudaStream_t s) {
  if (n_clouds <= 0) return;
  sc_make_kernel<<<n_clouds, 256, 0, s>>>(pts, d_offs, lidar_height, desc, ring_key, sector_key);
}

// one block of 64 threads per pair: thread c owns column c
__global__ void __launch_bounds__(64) sc_distance_kernel(const double* __restrict__ desc, const int2* __restrict__ pairs, int search_radius,
                                                         int key_align, double* __restrict__ dist, int32_t* __restrict__ shift_out) {
  __shared__ double a[kScSize], b[kScSize];
  __shared__ double ka[kScSectors], kb[kScSectors], na[kScSectors], nb[kScSectors];
  __shared__ double diffs[kScSectors], sims[kScSectors];
  __shared__ int s_align;
  const int2 pr = pairs[blockIdx.x];
  const int t = threadIdx.x;
  for (int k = t; k < kScSize; k += 64) { a[k] = desc[size_t(pr.x) * kScSize + k]; b[k] = desc[size_t(pr.y) * kScSize + k]; }
  __syncthreads();
  if (t < kScSectors) {  // sector keys (column means) and column norms
    double sa = 0, sb = 0, qa = 0, qb = 0;
    for (int r = 0; r < kScRings; r++) {
      const double va = a[r * kScSectors + t], vb = b[r * kScSectors + t];
      sa += va; sb += vb; qa += va * va; qb += vb * vb;
    }
    ka[t] = sa / double(kScRings); kb[t] = sb / double(kScRings);
    na[t] = sqrt(qa); nb[t] = sqrt(qb);
  }
  __syncthreads();
  // fastAlignUsingVkey (:89-108): shift of key 2 minimising |k1 - shift(k2)|, first minimum wins
  if (t < kScSectors) {
    double q = 0;
    for (int c = 0; c < kScSectors; c++) {
      const double d = ka[c] - kb[(c - t + kScSectors) % kScSectors];  // circshift: column c of the shifted key is column c - shift
      q = __dadd_rn(q, __dmul_rn(d, d));  // unfused, sequential: near-ties between shifts must resolve like the CPU restatement
    }
    diffs[t] = sqrt(q);
  }
  __syncthreads();
  if (t == 0) {
    // Reference quirk: the sector key is stored as an Eigen::VectorXd and converted to a 60 x 1 MatrixXd at the call site
    // (ScanContext.cpp:122-124), so fastAlignUsingVkey's loop over `_vkey1.cols()` (:93) runs for shift 0 only and the
    // alignment is always 0. key_align = 0 reproduces that; key_align = 1 searches all 60 shifts as the IROS'18 code does.
    int arg = 0;
    double best = DBL_MAX;
    const int nshift = key_align ? kScSectors : 1;
    for (int sft = 0; sft < nshift; sft++)
      if (diffs[sft] < best) { best = diffs[sft]; arg = sft; }
    s_align = arg;
  }
  __syncthreads();
  // candidate shifts: align, align +- 1 .. +- radius (mod 60), visited in ascending order (:127-134); columnwise cosine distance
  const int align = s_align;
  int cand[2 * 30 + 1];
  int nc = 0;
  for (int sft = 0; sft < kScSectors; sft++) {  // ascending order with duplicates kept exactly as std::sort of the list would
    int mult = 0;
    if (sft == align) mult++;
    for (int ii = 1; ii <= search_radius; ii++) {
      if (sft == (align + ii + kScSectors) % kScSectors) mult++;
      if (sft == (align - ii + kScSectors) % kScSectors) mult++;
    }
    for (int m = 0; m < mult && nc < 61; m++) cand[nc++] = sft;
  }
  double best = DBL_MAX;
  int best_shift = 0;
  for (int k = 0; k < nc; k++) {
    const int sft = cand[k];
    if (t < kScSectors) {
      const int cb = (t - sft + kScSectors) % kScSectors;
      double sim = 2.0;  // marker: column not counted
      if (!(na[t] == 0.0 || nb[cb] == 0.0)) {
        double dot = 0;
        for (int r = 0; r < kScRings; r++) dot += a[r * kScSectors + t] * b[r * kScSectors + cb];
        sim = dot / (na[t] * nb[cb]);
      }
      sims[t] = sim;
    }
    __syncthreads();
    double sum = 0;
    int eff = 0;
    for (int c = 0; c < kScSectors; c++)  // every thread repeats the sequential sum of the reference (:66-79): warp-uniform result
      if (sims[c] != 2.0) { sum = sum + sims[c]; eff++; }
    const double d = 1.0 - sum / double(eff);  // eff == 0 -> NaN, never smaller than the running minimum, like the reference
    if (d < best) { best = d; best_shift = sft; }
    __syncthreads();
  }
  if (t == 0) { dist[blockIdx.x] = best; shift_out[blockIdx.x] = best_shift; }
}

void scancontext_distance(const double* desc, const int2* pairs, int n_pairs, int search_radius, int key_align, double* dist, int32_t* shift,
                          cudaStream_t s) {
  if (n_pairs <= 0) return;
  sc_distance_kernel<<<n_pairs, 64, 0, s>>>(desc, pairs, search_radius, key_align, dist, shift);
}

}  // namespace pcr
