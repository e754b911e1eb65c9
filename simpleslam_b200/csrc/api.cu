// C ABI of the PCR CUDA library (include/pcr_cuda.h). No exceptions cross this boundary: every entry point maps
// failures to an error code + pcr_last_error() text (SURVEY.md §8b "Errors").
#include "../../include/pcr_cuda.h"
#include "common.cuh"
#include "hostpack.hpp"
#include "voxel.cuh"
#include "loam.cuh"
#include "ndt.cuh"
#include "vgicp.cuh"
#include "scancontext.cuh"
#include "host_math.hpp"
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <map>
#include <memory>
#include <string>
#include <condition_variable>
#include <thread>
#include <vector>

using namespace pcr;

struct pcr_ctx {
  pcr_params prm{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  bool profiling = false;
  pcr_stats stats{};
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;

  // staging
  DevBuf<unsigned char> raw_src, raw_dst;
  DevBuf<float4> src, dst, ds_in, nf_scratch;
  DevBuf<unsigned char> ds_out;
  PinBuf<unsigned char> pin;
  KeySort ks, ks_ds;
  BBoxWork bw;
  // downsample introspection
  GridSpec ds_grid{};
  size_t ds_n = 0, ds_m = 0;
  bool ds_overflow = false;

  // target state
  bool has_target = false;
  size_t n_target = 0;
  CellGrid loam_grid;
  NdtTarget ndt;
  VgicpTarget vg;

  // drivers
  LoamDriver loam;
  NdtDriver ndtd;
  VgicpDriver vgd;

  // submap assembly: device copies of immutable keyframe clouds, keyed by the caller's keyframe id, LRU within a byte budget
  struct CachedCloud { DevBuf<float4> pts; size_t n = 0; uint64_t last_use = 0; };
  std::map<int64_t, std::unique_ptr<CachedCloud>> kf_cache;
  size_t kf_budget = size_t(1) << 30;
  uint64_t kf_clock = 0;
  DevBuf<float4> sub_uncached;  // clouds of a build without ids
  DevBuf<float4> sub_concat;
  DevBuf<unsigned char> sub_meta;
  PinBuf<unsigned char> sub_meta_h;

  // pcr_batch_align of a large host batch: uploads of later chunks overlap the registration of earlier ones
  cudaStream_t copy_stream = nullptr;
  DevBuf<unsigned char> raw_chunk[2];
  std::vector<cudaEvent_t> ev_up;

  // uploads from pageable host memory: `cores` host threads pack into pinned staging, chunk by chunk (hostpack.hpp)
  std::unique_ptr<HostPacker> packer;
  PinBuf<float> pack_stage[2];
  cudaEvent_t ev_pack[2] = {nullptr, nullptr};
  size_t host_packed_bytes = 0;  // diagnostics: bytes that took this path in the last call

  KnnProfile knn_prof;  // VGICP k-NN kernel timing while profiling is on (target build + source covariances)

  // VGICP: last registration (for getFitnessScore)
  size_t last_ns = 0, last_off = 0;
  double last_T[16];
  bool has_last = false;
};

static thread_local std::string g_create_error;

// process-wide logger (pcr_set_logger); default: warnings and errors to stderr
static std::mutex g_log_mu;
static pcr_log_fn g_log_cb = nullptr;
static void* g_log_user = nullptr;
static void pcr_log(int level, const std::string& msg) {
  pcr_log_fn cb;
  void* user;
  {
    std::lock_guard<std::mutex> lk(g_log_mu);
    cb = g_log_cb; user = g_log_user;
  }
  if (cb) cb(level, msg.c_str(), user);
  else if (level >= 2) std::fprintf(stderr, "[pcr %s] %s\n", level >= 3 ? "error" : "warning", msg.c_str());
}
extern "C" void pcr_set_logger(pcr_log_fn cb, void* user) {
  std::lock_guard<std::mutex> lk(g_log_mu);
  g_log_cb = cb;
  g_log_user = user;
}

static LoamParams loam_params(const pcr_params& p) {
  LoamParams lp;
  lp.max_knn_d2 = double(p.loam_max_knn_d2);
  lp.plane_thresh = double(p.loam_plane_thresh);
  lp.point_thresh = double(p.loam_point_thresh);
  lp.pos_conv = double(p.loam_pos_converge);
  lp.rot_conv = double(p.loam_rot_converge);
  lp.max_iters = p.loam_max_iters;
  return lp;
}

// the caller's current device is restored when the call returns (a host with its own CUDA work on another GPU)
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    PCR_CUDA_CHECK(cudaSetDevice(dev));
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};

#define PCR_API_BEGIN(c)                                   \
  if (!(c)) return PCR_ERR_INVALID;                        \
  try {                                                    \
    DeviceScope pcr_device_scope((c)->device);
#define PCR_API_END(c)                                     \
  }                                                        \
  catch (const CudaError& e) {                             \
    (c)->err = e.what();                                   \
    pcr_log(3, (c)->err);                                  \
    cudaGetLastError();                                    \
    return PCR_ERR_CUDA;                                   \
  }                                                        \
  catch (const std::exception& e) {                        \
    (c)->err = e.what();                                   \
    pcr_log(3, (c)->err);                                  \
    return PCR_ERR_INVALID;                                \
  }

static int fail(pcr_ctx* c, int code, const char* msg) {
  c->err = msg;
  pcr_log(3, c->err);
  return code;
}

extern "C" void pcr_default_params(int32_t method, pcr_params* p) {
  memset(p, 0, sizeof(*p));
  p->method = method;
  p->device = 0;
  p->cores = 4;  // config/params.json:5
  p->loam_max_iters = 8;
  p->loam_max_knn_d2 = 1.0f;
  p->loam_plane_thresh = 0.2f;
  p->loam_point_thresh = 0.1f;
  p->loam_pos_converge = 5e-3f;
  p->loam_rot_converge = 5e-3f;
  p->ndt_resolution = 1.0f;
  p->ndt_search = PCR_NDT_DIRECT7;
  p->ndt_max_iters = 35;
  p->ndt_step_size = 0.1;
  p->ndt_outlier_ratio = 0.55;
  p->ndt_trans_eps = 0.1;
  p->ndt_min_points = 6;
  p->ndt_eig_mult = 0.01;
  p->vgicp_resolution = 1.0;
  p->vgicp_k = 20;
  p->vgicp_max_iters = 64;
  p->vgicp_optimizer = PCR_LSQ_LM;
  p->vgicp_lm_max_iters = 10;
  p->vgicp_rot_eps = 2e-3;
  p->vgicp_trans_eps = 5e-4;
  p->vgicp_lm_init_lambda = 1e-9;
}

extern "C" int pcr_create(const pcr_params* p, pcr_ctx** out) {
  if (!p || !out) return PCR_ERR_INVALID;
  *out = nullptr;
  if (p->method < PCR_LOAM || p->method > PCR_VGICP) { g_create_error = "unknown method (expected loam|ndt|vgicp)"; return PCR_ERR_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    pcr_log(3, g_create_error);
    return PCR_ERR_NO_DEVICE;
  }
  if (p->device < 0 || p->device >= ndev) { g_create_error = "device ordinal out of range"; return PCR_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess || prop.major < 10) {
    g_create_error = "device is not sm_100-class (B200); kernels are built for sm_100a only";
    return PCR_ERR_NO_DEVICE;
  }
  pcr_ctx* c = new pcr_ctx();
  c->prm = *p;
  c->device = p->device;
  try {
    DeviceScope scope(c->device);
    PCR_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PCR_CUDA_CHECK(cudaEventCreate(&c->ev_a));
    PCR_CUDA_CHECK(cudaEventCreate(&c->ev_b));
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    delete c;
    return PCR_ERR_CUDA;
  }
  *out = c;
  return PCR_OK;
}

extern "C" void pcr_destroy(pcr_ctx* c) {
  if (!c) return;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  cudaSetDevice(c->device);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  for (cudaEvent_t e : c->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_pack) if (e) cudaEventDestroy(e);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  delete c;
  if (prev >= 0) cudaSetDevice(prev);
}

extern "C" const char* pcr_last_error(const pcr_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

extern "C" int pcr_vgicp_init_for_lc(pcr_ctx* c) {
  if (!c) return PCR_ERR_INVALID;
  // VgicpRegister.cpp:21-28. setMaxCorrespondenceDistance(150), EuclideanFitnessEpsilon and RANSAC are no-ops for the
  // voxel-correspondence LSQ path.
  c->prm.vgicp_max_iters = 100;
  c->prm.vgicp_trans_eps = 1e-6;
  return PCR_OK;
}

extern "C" int pcr_set_profiling(pcr_ctx* c, int enable) {
  if (!c) return PCR_ERR_INVALID;
  c->profiling = enable != 0;
  c->vg.prof = c->profiling ? &c->knn_prof : nullptr;
  c->vgd.prof = c->vg.prof;
  c->knn_prof.reset();
  return PCR_OK;
}

extern "C" int pcr_get_stats(const pcr_ctx* c, pcr_stats* s) {
  if (!c || !s) return PCR_ERR_INVALID;
  *s = c->stats;
  return PCR_OK;
}

// ---- uploads -------------------------------------------------------------------------------------------------------
// 0 = ordinary pageable host memory, 1 = pinned / registered host memory, 2 = anything else (device, managed)
static int host_memory_kind(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (a.type == cudaMemoryTypeUnregistered) return 0;
  return a.type == cudaMemoryTypeHost ? 1 : 2;
}

constexpr size_t kPackChunkPts = size_t(1) << 20;  // 16 MB of float4 per staging buffer
constexpr int kPackPinnedCores = 8;                // host threads from which packing beats the plain DMA of pinned 32-byte records
constexpr size_t kPackPinnedBytes = size_t(256) << 20;  // a 67 MB map gains nothing (1.58 vs 1.50 ms): only batches that stream for tens of ms

// Pageable sources are always packed by the host threads (the driver's own staging copy is single-threaded). Pinned
// sources cross PCIe at ~50 GB/s as they are; packing them first halves the bytes but needs >= 8 threads to keep up
// (measured on the 1024-scan job, 4.26 GB: plain DMA 85 ms, packed with 4 / 8 / 12 threads 115 / 82 / 76 ms).
static bool use_host_pack(const pcr_ctx* c, const void* host, size_t bytes) {
  const char* e = std::getenv("PCR_HOST_PACK");  // read on every call: a test / tuning knob
  const int force = e ? (std::atoi(e) != 0 ? 1 : 0) : -1;
  static const size_t min_bytes = [] { const char* m = std::getenv("PCR_HOST_PACK_MIN_KB"); return size_t(m ? std::max(1, std::atoi(m)) : 1024) << 10; }();
  if (force >= 0) return force == 1;
  if (c->prm.cores <= 0 || bytes < min_bytes) return false;
  const int kind = host_memory_kind(host);
  return kind == 0 || (kind == 1 && c->prm.cores >= kPackPinnedCores && bytes >= kPackPinnedBytes);
}

// host AoS records -> float4 records at `out` (device), ordered on stream s. Pageable sources of some size are packed by the
// context's host threads into pinned staging and cross PCIe as 16-byte records while the next chunk is packed; everything
// else is copied as it is and packed by pack_kernel. PCR_HOST_PACK=0 / 1 forces the choice (tests, A/B).
// pack_mode: -1 = decide here from the size and kind of the source, 0 / 1 = the caller decided for a whole batch of chunks.
static void upload_host_cloud(pcr_ctx* c, const void* host, size_t n, size_t stride, DevBuf<unsigned char>& raw, float4* out, cudaStream_t s,
                              int pack_mode = -1) {
  if (n == 0) return;
  if (!(pack_mode < 0 ? use_host_pack(c, host, n * stride) : pack_mode != 0)) {
    raw.ensure(n * stride);
    PCR_CUDA_CHECK(cudaMemcpyAsync(raw.p, host, n * stride, cudaMemcpyHostToDevice, s));
    pack_points(raw.p, n, stride, out, s);
    return;
  }
  if (!c->packer) c->packer.reset(new HostPacker(c->prm.cores));
  const size_t chunk = std::min(n, kPackChunkPts);
  for (int b = 0; b < 2; b++) {
    if (!c->ev_pack[b]) PCR_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_pack[b], cudaEventDisableTiming));
    if (c->pack_stage[b].cap < chunk * 4) {
      PCR_CUDA_CHECK(cudaEventSynchronize(c->ev_pack[b]));  // a copy of an earlier call may still read the buffer
      c->pack_stage[b].ensure(kPackChunkPts * 4);
    }
  }
  const unsigned char* bytes = static_cast<const unsigned char*>(host);
  size_t k = 0;
  for (size_t p0 = 0; p0 < n; p0 += kPackChunkPts, k++) {
    const size_t cnt = std::min(kPackChunkPts, n - p0);
    const int b = int(k & 1);
    PCR_CUDA_CHECK(cudaEventSynchronize(c->ev_pack[b]));  // the copy that last read this staging buffer has completed
    c->packer->pack(bytes + p0 * stride, cnt, stride, c->pack_stage[b].p);
    PCR_CUDA_CHECK(cudaMemcpyAsync(out + p0, c->pack_stage[b].p, cnt * sizeof(float4), cudaMemcpyHostToDevice, s));
    PCR_CUDA_CHECK(cudaEventRecord(c->ev_pack[b], s));
  }
  c->host_packed_bytes += n * stride;
}

static const float4* upload_points(pcr_ctx* c, const void* host, size_t n, size_t stride, DevBuf<unsigned char>& raw, DevBuf<float4>& out) {
  out.ensure(n + 1);
  if (n == 0) return out.p;
  upload_host_cloud(c, host, n, stride, raw, out.p, c->stream);
  return out.p;
}
static const float4* adopt_points(pcr_ctx* c, const void* dev, size_t n, size_t stride, DevBuf<float4>& out) {
  out.ensure(n + 1);
  if (n) pack_points(dev, n, stride, out.p, c->stream);
  return out.p;
}

static int build_target_once(pcr_ctx* c, const float4* pts, size_t n) {
  switch (c->prm.method) {
    case PCR_LOAM: return loam_build_target(pts, n, double(c->prm.loam_max_knn_d2), c->loam_grid, c->ks, c->bw, c->stream);
    case PCR_NDT: return ndt_build_target(pts, n, c->prm, c->ndt, c->ks, c->bw, c->stream);
    case PCR_VGICP: return vgicp_build_target(pts, n, c->prm, c->vg, c->ks, c->bw, c->stream);
  }
  return PCR_ERR_INVALID;
}

// pts: a context-owned packed copy of the cloud (c->dst). Records with a NaN / Inf coordinate are dropped (order kept) the
// way the reference never lets them reach a register (LidarDataProxy.cpp:47) and PCL's voxel grids skip them.
static int build_target(pcr_ctx* c, const float4* pts, size_t n) {
  c->has_target = false;
  c->n_target = n;
  c->has_last = false;
  int rc = build_target_once(c, pts, n);
  if (rc == kRetryNonFinite) {
    n = drop_nonfinite(const_cast<float4*>(pts), n, c->nf_scratch, c->ks.tmp, c->ks.d_count, c->ks.h_count, c->stream);
    c->n_target = n;
    rc = build_target_once(c, pts, n);
  }
  if (rc == PCR_ERR_GRID_TOO_LARGE) return fail(c, rc, "target bounding box needs a cell table larger than the dense-table budget");
  if (rc) return fail(c, rc, "target build failed");
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PCR_CUDA_CHECK(cudaGetLastError());
  c->has_target = true;
  return PCR_OK;
}

extern "C" int pcr_set_target(pcr_ctx* c, const void* pts, size_t n, size_t stride) {
  PCR_API_BEGIN(c)
  if ((n && !pts) || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad target cloud");
  const float4* d = upload_points(c, pts, n, stride, c->raw_dst, c->dst);
  return build_target(c, d, n);
  PCR_API_END(c)
}

extern "C" int pcr_set_target_device(pcr_ctx* c, const void* dev_pts, size_t n, size_t stride) {
  PCR_API_BEGIN(c)
  if ((n && !dev_pts) || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad target cloud");
  const float4* d = adopt_points(c, dev_pts, n, stride, c->dst);
  return build_target(c, d, n);
  PCR_API_END(c)
}

// ---- align ---------------------------------------------------------------------------------------------------------
static int align_packed(pcr_ctx* c, const float4* src, const size_t* offs, size_t n_scans, double* T, int32_t* converged,
                        const NdtArrivals* arrivals = nullptr) {
  if (!c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no target set");
  pcr_stats& st = c->stats;
  memset(&st, 0, sizeof(st));
  st.n_source = int64_t(offs[n_scans] - offs[0]);
  st.n_target = int64_t(c->n_target);
  PCR_CUDA_CHECK(cudaEventRecord(c->ev_a, c->stream));
  int rc = 0;
  std::vector<int32_t> conv(n_scans, 0), iters(n_scans, 0);
  switch (c->prm.method) {
    case PCR_LOAM: {
      std::vector<int64_t> nl(n_scans, 0);
      rc = c->loam.align(src, offs, n_scans, c->loam_grid, loam_params(c->prm), T, conv.data(), iters.data(), nl.data(), c->profiling,
                         c->stream);
      st.iterations = iters[0];
      st.evaluations = 0;
      for (size_t i = 0; i < n_scans; i++) st.evaluations += iters[i];
      st.n_residuals = nl[0];
      st.kernel_launches = c->loam.launches + 1;  // + pack kernel
      st.n_pairs = c->loam.cand_total;
      st.n_point_evals = c->loam.pt_evals;
      st.n_index_reads = c->loam.rows_total;
      st.ms_hot_kernel = c->loam.hot_ms;
      st.hot_kernel_launches = c->loam.hot_launches;
      break;
    }
    case PCR_NDT: {
      std::vector<double> tp(n_scans, 0.0);
      rc = c->ndtd.align(src, offs, n_scans, c->ndt, c->prm, T, conv.data(), iters.data(), tp.data(), c->profiling, c->stream, arrivals);
      st.iterations = iters[0];
      st.evaluations = c->ndtd.total_evals;
      st.hessian_evals = c->ndtd.total_hess;
      st.score = tp[0];
      st.kernel_launches = c->ndtd.launches + 1;
      st.n_pairs = c->ndtd.total_pairs;
      st.n_point_evals = c->ndtd.point_evals;
      st.n_index_reads = c->ndtd.point_evals * (c->prm.ndt_search == PCR_NDT_DIRECT1 ? 1 : c->prm.ndt_search == PCR_NDT_DIRECT7 ? 7 : c->prm.ndt_search == PCR_NDT_DIRECT26 ? 26 : 27);
      st.ms_hot_kernel = c->ndtd.hot_ms;
      st.hot_kernel_launches = c->ndtd.hot_launches;
      break;
    }
    case PCR_VGICP: {
      c->vgd.launches = 0;
      float hot = 0.f;
      int hotl = 0, evals = 0;
      long long corr = 0;
      size_t last_off = 0, last_cnt = 0;
      for (size_t i = 0; i < n_scans && rc == 0; i++) {  // independent scans, processed one after the other
        size_t ns = offs[i + 1] - offs[i];
        const float4* sp = src + (offs[i] - offs[0]);
        rc = c->vgd.compute_source_covs(sp, ns, c->prm.vgicp_k, c->ks, c->bw, c->stream);
        if (rc == kRetryNonFinite) {  // NaN / Inf records of this scan are dropped inside its own segment of the staging copy
          ns = drop_nonfinite(const_cast<float4*>(sp), ns, c->nf_scratch, c->ks.tmp, c->ks.d_count, c->ks.h_count, c->stream);
          rc = c->vgd.compute_source_covs(sp, ns, c->prm.vgicp_k, c->ks, c->bw, c->stream);
        }
        if (rc) break;
        last_off = offs[i] - offs[0];
        last_cnt = ns;
        rc = c->vgd.align(sp, ns, c->vg, c->prm, T + i * 16, &conv[i], &iters[i], c->profiling, c->stream);
        hot += c->vgd.hot_ms;
        hotl += c->vgd.hot_launches;
        evals += c->vgd.n_linearize + c->vgd.n_error;
        corr += c->vgd.total_corr;
      }
      st.iterations = iters[0];
      st.evaluations = evals;
      st.n_residuals = c->vgd.last_corr;
      st.score = c->vgd.last_cost;
      st.kernel_launches = c->vgd.launches + 1;
      st.n_pairs = corr;
      st.n_point_evals = int64_t(evals) * int64_t(offs[1] - offs[0]);
      st.n_index_reads = st.n_point_evals;
      st.ms_hot_kernel = hot;
      st.hot_kernel_launches = hotl;
      if (c->profiling) {  // k-NN kernel time since the last align: the target build of this registration + the source scans
        PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->knn_prof.collect();
        st.ms_aux_kernel = c->knn_prof.ms;
        st.aux_kernel_launches = c->knn_prof.launches;
        st.n_aux_items = c->knn_prof.queries;
        c->knn_prof.reset();
      }
      // remember the last scan for getFitnessScore (pcl keeps input_ + final_transformation_)
      c->has_last = false;
      if (rc == 0) {
        c->last_ns = last_cnt;
        c->last_off = last_off;
        memcpy(c->last_T, T + (n_scans - 1) * 16, sizeof(double) * 16);
        c->has_last = true;
      }
      break;
    }
  }
  if (rc == PCR_ERR_UNSUPPORTED) return fail(c, rc, "unsupported option");
  if (rc) return fail(c, rc, "align failed");
  PCR_CUDA_CHECK(cudaEventRecord(c->ev_b, c->stream));
  PCR_CUDA_CHECK(cudaEventSynchronize(c->ev_b));
  PCR_CUDA_CHECK(cudaEventElapsedTime(&st.ms_total, c->ev_a, c->ev_b));
  st.converged = conv[0];
  if (converged)
    for (size_t i = 0; i < n_scans; i++) converged[i] = conv[i];
  return PCR_OK;
}

extern "C" int pcr_align(pcr_ctx* c, const void* src, size_t n, size_t stride, double T[16], int32_t* converged) {
  PCR_API_BEGIN(c)
  if ((n && !src) || !T || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad source cloud");
  const float4* d = upload_points(c, src, n, stride, c->raw_src, c->src);
  size_t offs[2] = {0, n};
  return align_packed(c, d, offs, 1, T, converged);
  PCR_API_END(c)
}

extern "C" int pcr_align_device(pcr_ctx* c, const void* dev_src, size_t n, size_t stride, double T[16], int32_t* converged) {
  PCR_API_BEGIN(c)
  if ((n && !dev_src) || !T || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad source cloud");
  const float4* d = adopt_points(c, dev_src, n, stride, c->src);
  size_t offs[2] = {0, n};
  return align_packed(c, d, offs, 1, T, converged);
  PCR_API_END(c)
}

extern "C" int pcr_scan2map(pcr_ctx* c, const void* src, size_t ns, size_t sstride, const void* dst, size_t nm, size_t dstride,
                            double T[16], int32_t* converged) {
  int rc = pcr_set_target(c, dst, nm, dstride);
  if (rc) return rc;
  return pcr_align(c, src, ns, sstride, T, converged);
}

// Large host batches (loc.cpp mode: hundreds of scans per call) are cut into chunks of whole scans: while chunk k registers
// on the context's stream, chunks k+1 and k+2 cross PCIe on a copy stream (two staging buffers) and are packed into their
// slice of the resident float4 array. Results and statistics are those of one call.
static void merge_stats(pcr_stats& acc, const pcr_stats& s, bool first) {
  if (first) { acc = s; return; }
  acc.evaluations += s.evaluations; acc.hessian_evals += s.hessian_evals;
  acc.n_source += s.n_source;
  acc.kernel_launches += s.kernel_launches; acc.n_pairs += s.n_pairs; acc.n_point_evals += s.n_point_evals; acc.n_index_reads += s.n_index_reads;
  acc.ms_total += s.ms_total; acc.ms_hot_kernel += s.ms_hot_kernel; acc.hot_kernel_launches += s.hot_kernel_launches;
  acc.ms_aux_kernel += s.ms_aux_kernel; acc.aux_kernel_launches += s.aux_kernel_launches; acc.n_aux_items += s.n_aux_items;
}

extern "C" int pcr_batch_align(pcr_ctx* c, const void* src, const size_t* offsets, size_t n_scans, size_t stride, double* T,
                               int32_t* converged) {
  PCR_API_BEGIN(c)
  if (!offsets || !T || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad batch");
  if (n_scans == 0) return PCR_OK;
  const size_t n = offsets[n_scans] - offsets[0];
  const unsigned char* base = static_cast<const unsigned char*>(src) + offsets[0] * stride;
  const char* chunk_env = std::getenv("PCR_BATCH_CHUNK_MB");  // tuning / test knob
  const size_t chunk_bytes = size_t(chunk_env ? std::max(1, std::atoi(chunk_env)) : 256) << 20;
  if (n_scans < 8 || n * stride < 2 * chunk_bytes || !c->has_target) {
    const float4* d = upload_points(c, base, n, stride, c->raw_src, c->src);
    return align_packed(c, d, offsets, n_scans, T, converged);
  }
  // ---- chunks of whole scans, about chunk_bytes each
  std::vector<size_t> cut(1, 0);
  for (size_t i = 1; i <= n_scans; i++)
    if (i == n_scans || (offsets[i] - offsets[cut.back()]) * stride >= chunk_bytes) cut.push_back(i);
  const size_t nc = cut.size() - 1;
  if (!c->copy_stream) PCR_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  while (c->ev_up.size() < nc) {
    cudaEvent_t e;
    PCR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_up.push_back(e);
  }
  c->src.ensure(n + 1);
  const bool pack = use_host_pack(c, base, n * stride);  // one decision for the whole batch: every chunk takes the same path
  if (!pack) {  // raw copies of two chunks in flight; sized here, not on the uploader thread
    size_t max_bytes = 0;
    for (size_t k = 0; k < nc; k++) max_bytes = std::max(max_bytes, (offsets[cut[k + 1]] - offsets[cut[k]]) * stride);
    c->raw_chunk[0].ensure(max_bytes);
    c->raw_chunk[1].ensure(max_bytes);
  }
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));  // nothing of an earlier call still reads c->src
  auto upload = [&](size_t k) {
    const size_t p0 = offsets[cut[k]] - offsets[0], pn = offsets[cut[k + 1]] - offsets[cut[k]];
    if (pn)  // the copy stream is in order: staging buffer k & 1 is free again afterwards
      upload_host_cloud(c, base + p0 * stride, pn, stride, c->raw_chunk[k & 1], c->src.p + p0, c->copy_stream, pack ? 1 : 0);
    PCR_CUDA_CHECK(cudaEventRecord(c->ev_up[k], c->copy_stream));
  };
  // the uploads run on their own host thread: a copy from pageable memory blocks its caller while the driver stages it, and
  // that must not stall the thread that drives the registrations
  std::mutex mu;
  std::condition_variable cv;
  size_t uploaded = 0;
  std::string up_err;
  std::thread uploader([&]() {
    try {
      PCR_CUDA_CHECK(cudaSetDevice(c->device));
      for (size_t k = 0; k < nc; k++) {
        upload(k);
        { std::lock_guard<std::mutex> lk(mu); uploaded = k + 1; }
        cv.notify_all();
      }
    } catch (const std::exception& e) {
      { std::lock_guard<std::mutex> lk(mu); up_err = e.what(); uploaded = nc; }
      cv.notify_all();
    }
  });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{uploader};  // also on the exception paths
  auto wait_enqueued = [&](size_t k) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return uploaded > k; });
    if (!up_err.empty()) throw CudaError(up_err);
  };
  pcr_stats total{};
  int rc = PCR_OK;
  if (c->prm.method == PCR_NDT && n_scans <= NdtDriver::max_streaming_scans) {
    // NDT: ONE batch; the scans of a chunk join the running evaluation rounds as soon as their points are on the device
    NdtArrivals arr;
    arr.n_groups = nc;
    arr.first = cut.data();
    arr.ready = [&](size_t k) {
      { std::lock_guard<std::mutex> lk(mu); if (uploaded <= k) return false; if (!up_err.empty()) throw CudaError(up_err); }
      const cudaError_t q = cudaEventQuery(c->ev_up[k]);
      if (q != cudaSuccess && q != cudaErrorNotReady) PCR_CUDA_CHECK(q);
      return q == cudaSuccess;
    };
    arr.wait = wait_enqueued;   // the registration stream itself then waits for the event
    arr.event = [&](size_t k) { return c->ev_up[k]; };
    rc = align_packed(c, c->src.p, offsets, n_scans, T, converged, &arr);
    if (rc == PCR_OK) total = c->stats;
  } else {
    for (size_t k = 0; k < nc && rc == PCR_OK; k++) {
      wait_enqueued(k);
      const size_t p0 = offsets[cut[k]] - offsets[0];
      PCR_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_up[k], 0));
      rc = align_packed(c, c->src.p + p0, offsets + cut[k], cut[k + 1] - cut[k], T + 16 * cut[k], converged ? converged + cut[k] : nullptr);
      if (rc == PCR_OK) {
        merge_stats(total, c->stats, k == 0);
        if (c->prm.method == PCR_VGICP) c->last_off += p0;  // getFitnessScore looks the last scan up in the whole staging array
      }
    }
  }
  uploader.join();
  if (rc != PCR_OK) { cudaStreamSynchronize(c->copy_stream); cudaGetLastError(); pcr_log(3, c->err); return rc; }
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
  if (rc == PCR_OK) c->stats = total;
  return rc;
  PCR_API_END(c)
}

extern "C" int pcr_batch_align_device(pcr_ctx* c, const void* dev_src, const size_t* offsets, size_t n_scans, size_t stride, double* T,
                                      int32_t* converged) {
  PCR_API_BEGIN(c)
  if (!offsets || !T || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad batch");
  if (n_scans == 0) return PCR_OK;
  const size_t n = offsets[n_scans] - offsets[0];
  const unsigned char* base = static_cast<const unsigned char*>(dev_src) + offsets[0] * stride;
  const float4* d = adopt_points(c, base, n, stride, c->src);
  return align_packed(c, d, offsets, n_scans, T, converged);
  PCR_API_END(c)
}

extern "C" int pcr_fitness(pcr_ctx* c, double* score) {
  PCR_API_BEGIN(c)
  if (!score) return PCR_ERR_INVALID;
  *score = 0.0;  // PointCloudRegister::getFitnessScore() base implementation returns 0 (LOAM / NDT)
  if (c->prm.method != PCR_VGICP) return PCR_OK;
  // pcl::Registration::getFitnessScore answers DBL_MAX when it has nothing to measure: a failed call must never look like a
  // perfect match to the loop-closure gate (`fitness < 0.3`, LoopClosureManager.cpp:98)
  *score = DBL_MAX;
  if (!c->has_target || !c->has_last) return fail(c, PCR_ERR_NO_TARGET, "getFitnessScore before scan2Map");
  // the last source scan is still resident in the staging copy c->src
  const float4* sp = c->src.p + c->last_off;
  return c->vgd.fitness(sp, c->last_ns, c->vg, c->last_T, DBL_MAX, score, c->stream);
  PCR_API_END(c)
}

// ---- voxel downsample ----------------------------------------------------------------------------------------------
static int downsample_packed(pcr_ctx* c, const float4* pts, size_t n, float leaf, void* dev_out32, size_t cap, size_t* m) {
  c->ds_n = n;
  c->ds_m = 0;
  c->ds_overflow = false;
  if (n == 0) { *m = 0; return PCR_OK; }
  float mn[3], mx[3];
  bbox_blocking(pts, n, mn, mx, c->bw, c->stream);
  if (c->bw.n_nonfinite) {  // pcl::VoxelGrid skips non-finite points of a non-dense cloud (voxel_grid.hpp: `if (!isFinite(point)) continue`)
    n = drop_nonfinite(const_cast<float4*>(pts), n, c->nf_scratch, c->ks_ds.tmp, c->ks_ds.d_count, c->ks_ds.h_count, c->stream);
    c->ds_n = n;
    if (n == 0) { *m = 0; return PCR_OK; }
    bbox_blocking(pts, n, mn, mx, c->bw, c->stream);
  }
  if (!make_grid_spec(mn, mx, leaf, c->ds_grid)) {
    // PCL: "Leaf size is too small for the input dataset" -> the input is returned unchanged
    c->ds_overflow = true;
    if (cap < n) return fail(c, PCR_ERR_INVALID, "output capacity too small");
    write_xyzi32(pts, n, dev_out32, c->stream);
    *m = n;
    c->ds_m = n;
    return PCR_OK;
  }
  c->ks_ds.sort(pts, n, c->ds_grid, c->stream);
  c->ks_ds.segment(c->stream);
  if (c->ks_ds.nseg > cap) return fail(c, PCR_ERR_INVALID, "output capacity too small");
  voxel_centroids(pts, c->ks_ds, dev_out32, c->stream);
  *m = c->ks_ds.nseg;
  c->ds_m = *m;
  return PCR_OK;
}

extern "C" int pcr_voxel_downsample(pcr_ctx* c, const void* pts, size_t n, size_t stride, float leaf, void* out, size_t cap, size_t* m) {
  PCR_API_BEGIN(c)
  if ((n && !pts) || !m || !(leaf > 0.f) || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad arguments");
  const float4* d = upload_points(c, pts, n, stride, c->raw_src, c->ds_in);
  c->ds_out.ensure(std::max<size_t>(n, 1) * 32);
  int rc = downsample_packed(c, d, n, leaf, c->ds_out.p, std::min(cap, n), m);
  if (rc) return rc;
  if (*m) PCR_CUDA_CHECK(cudaMemcpyAsync(out, c->ds_out.p, *m * 32, cudaMemcpyDeviceToHost, c->stream));
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PCR_CUDA_CHECK(cudaGetLastError());
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_voxel_downsample_device(pcr_ctx* c, const void* dev_pts, size_t n, size_t stride, float leaf, void* dev_out, size_t cap,
                                           size_t* m) {
  PCR_API_BEGIN(c)
  if ((n && !dev_pts) || !m || !(leaf > 0.f) || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad arguments");
  const float4* d = adopt_points(c, dev_pts, n, stride, c->ds_in);
  int rc = downsample_packed(c, d, n, leaf, dev_out, cap, m);
  if (rc) return rc;
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PCR_CUDA_CHECK(cudaGetLastError());
  return PCR_OK;
  PCR_API_END(c)
}

// One frame of LidarOdometry::generateOdom (LidarOdometry.cpp:170-184): mVoxelGrid.filter(scan) -> mPcr->scan2Map(ds, submap,
// pose) against the resident target. The downsampled scan stays on the device (the two-call form downloads it and uploads it
// again); its records go through the same 32-byte layout and pack step, so the registration sees the same bits.
extern "C" int pcr_downsample_align(pcr_ctx* c, const void* scan, size_t n, size_t stride, float leaf, double T[16], int32_t* converged,
                                    size_t* m_out) {
  PCR_API_BEGIN(c)
  if ((n && !scan) || !T || !(leaf > 0.f) || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad arguments");
  if (!c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no target set");
  const float4* d = upload_points(c, scan, n, stride, c->raw_src, c->ds_in);
  c->ds_out.ensure(std::max<size_t>(n, 1) * 32);
  size_t m = 0;
  int rc = downsample_packed(c, d, n, leaf, c->ds_out.p, n, &m);
  if (rc) return rc;
  if (m_out) *m_out = m;
  const float4* ds = adopt_points(c, c->ds_out.p, m, 32, c->src);
  size_t offs[2] = {0, m};
  return align_packed(c, ds, offs, 1, T, converged);
  PCR_API_END(c)
}

// ---- submap assembly (MapManager::updateMap / loopFindNearKeyframes) ------------------------------------------------
struct SubmapPart {
  const float4* src;
  unsigned long long offset;  // first output record of this cloud
  unsigned long long n;
  float T[12];                // rows of the float 3x4 [R | t]
};

__global__ void __launch_bounds__(256) submap_transform_kernel(const SubmapPart* __restrict__ parts, int n_parts, size_t total,
                                                               float4* __restrict__ out) {
  extern __shared__ unsigned long long s_off[];  // n_parts + 1 offsets
  for (int k = threadIdx.x; k <= n_parts; k += blockDim.x) s_off[k] = k < n_parts ? parts[k].offset : (unsigned long long)total;
  __syncthreads();
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = n_parts - 1;  // last part whose offset <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  const SubmapPart& pt = parts[lo];
  const float4 p = __ldg(pt.src + (i - pt.offset));
  // pcp::transformPointCloud: pto = tr * pfrom with tr = pose.cast<float>()  ->  ((r0 x + r1 y) + r2 z) + t, intensity kept
  float4 o;
  o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pt.T[0], p.x), __fmul_rn(pt.T[1], p.y)), __fmul_rn(pt.T[2], p.z)), pt.T[3]);
  o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pt.T[4], p.x), __fmul_rn(pt.T[5], p.y)), __fmul_rn(pt.T[6], p.z)), pt.T[7]);
  o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pt.T[8], p.x), __fmul_rn(pt.T[9], p.y)), __fmul_rn(pt.T[10], p.z)), pt.T[11]);
  o.w = p.w;
  out[i] = o;
}

extern "C" int pcr_submap_cache_clear(pcr_ctx* c) {
  if (!c) return PCR_ERR_INVALID;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->kf_cache.clear();
  return PCR_OK;
}

extern "C" int pcr_submap_cache_budget(pcr_ctx* c, size_t bytes) {
  if (!c) return PCR_ERR_INVALID;
  c->kf_budget = bytes;
  return PCR_OK;
}

extern "C" int pcr_submap_cache_info(const pcr_ctx* c, size_t* bytes, size_t* entries) {
  if (!c) return PCR_ERR_INVALID;
  size_t b = 0;
  for (const auto& kv : c->kf_cache) b += kv.second->pts.bytes_;
  if (bytes) *bytes = b;
  if (entries) *entries = c->kf_cache.size();
  return PCR_OK;
}

extern "C" int pcr_submap_build(pcr_ctx* c, const void* const* clouds, const size_t* counts, const int64_t* ids, size_t n_clouds, size_t stride,
                                const double* poses, float leaf, void* out, size_t cap, size_t* m) {
  PCR_API_BEGIN(c)
  if (!m || !(leaf > 0.f) || stride < 12 || stride % 4 || (n_clouds && (!clouds || !counts || !poses))) return fail(c, PCR_ERR_INVALID, "bad arguments");
  *m = 0;
  static const bool trace = std::getenv("PCR_TRACE") != nullptr;  // stage timings on stderr (tuning aid)
  auto now = [&]() { if (trace) cudaStreamSynchronize(c->stream); return std::chrono::steady_clock::now(); };
  auto t_start = now();
  SubmapPart* hp = reinterpret_cast<SubmapPart*>(c->sub_meta_h.ensure((n_clouds + 1) * sizeof(SubmapPart)));
  size_t total = 0, np = 0, uncached_total = 0;
  if (!ids) {
    for (size_t k = 0; k < n_clouds; k++) uncached_total += counts[k];
    c->sub_uncached.ensure(uncached_total + 1);
  }
  const uint64_t stamp = ++c->kf_clock;
  size_t unc_off = 0;
  for (size_t k = 0; k < n_clouds; k++) {
    if (counts[k] == 0) continue;
    if (!clouds[k]) return fail(c, PCR_ERR_INVALID, "null keyframe cloud");
    const float4* dev_pts = nullptr;
    if (ids) {
      auto it = c->kf_cache.find(ids[k]);
      if (it == c->kf_cache.end() || it->second->n != counts[k]) {
        std::unique_ptr<pcr_ctx::CachedCloud> cc(new pcr_ctx::CachedCloud());
        cc->n = counts[k];
        upload_points(c, clouds[k], counts[k], stride, c->raw_src, cc->pts);
        // raw_src is reused by the next upload: the pack kernel of this one is already queued on the same stream
        if (it != c->kf_cache.end()) { cudaStreamSynchronize(c->stream); c->kf_cache.erase(it); }
        it = c->kf_cache.emplace(ids[k], std::move(cc)).first;
      }
      it->second->last_use = stamp;
      dev_pts = it->second->pts.p;
    } else {
      c->raw_src.ensure(counts[k] * stride);
      PCR_CUDA_CHECK(cudaMemcpyAsync(c->raw_src.p, clouds[k], counts[k] * stride, cudaMemcpyHostToDevice, c->stream));
      pack_points(c->raw_src.p, counts[k], stride, c->sub_uncached.p + unc_off, c->stream);
      dev_pts = c->sub_uncached.p + unc_off;
      unc_off += counts[k];
    }
    SubmapPart& p = hp[np++];
    p.src = dev_pts;
    p.offset = total;
    p.n = counts[k];
    const double* T = poses + k * 16;
    for (int r = 0; r < 3; r++) {
      p.T[r * 4 + 0] = float(T[r]); p.T[r * 4 + 1] = float(T[4 + r]); p.T[r * 4 + 2] = float(T[8 + r]); p.T[r * 4 + 3] = float(T[12 + r]);
    }
    total += counts[k];
  }
  if (ids) {  // bound the cache: least-recently-used entries that are not part of this submap go first
    size_t bytes = 0;
    for (const auto& kv : c->kf_cache) bytes += kv.second->pts.bytes_;
    bool synced = false;
    while (bytes > c->kf_budget) {
      auto victim = c->kf_cache.end();
      for (auto it = c->kf_cache.begin(); it != c->kf_cache.end(); ++it)
        if (it->second->last_use != stamp && (victim == c->kf_cache.end() || it->second->last_use < victim->second->last_use)) victim = it;
      if (victim == c->kf_cache.end()) break;  // everything left belongs to the current submap
      if (!synced) { PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream)); synced = true; }  // an older build may still read it
      bytes -= victim->second->pts.bytes_;
      c->kf_cache.erase(victim);
    }
  }
  c->has_target = false;
  c->has_last = false;
  c->n_target = 0;
  if (total == 0) {  // empty submap: registered as an empty target
    c->dst.ensure(1);
    return build_target(c, c->dst.p, 0);
  }
  auto t_up = now();
  c->sub_meta.ensure(np * sizeof(SubmapPart));
  PCR_CUDA_CHECK(cudaMemcpyAsync(c->sub_meta.p, hp, np * sizeof(SubmapPart), cudaMemcpyHostToDevice, c->stream));
  c->sub_concat.ensure(total);
  if ((np + 1) * sizeof(unsigned long long) > 40 * 1024) return fail(c, PCR_ERR_INVALID, "too many keyframe clouds in one submap (limit 5000)");
  submap_transform_kernel<<<unsigned((total + 255) / 256), 256, (np + 1) * sizeof(unsigned long long), c->stream>>>(
      reinterpret_cast<const SubmapPart*>(c->sub_meta.p), int(np), total, c->sub_concat.p);
  PCR_CUDA_CHECK(cudaGetLastError());
  auto t_tr = now();
  c->ds_out.ensure(total * 32);
  size_t mm = 0;
  int rc = downsample_packed(c, c->sub_concat.p, total, leaf, c->ds_out.p, total, &mm);
  if (rc) return rc;
  if (out) {
    if (mm > cap) return fail(c, PCR_ERR_INVALID, "output capacity too small");
    PCR_CUDA_CHECK(cudaMemcpyAsync(out, c->ds_out.p, mm * 32, cudaMemcpyDeviceToHost, c->stream));
  }
  *m = mm;
  auto t_ds = now();
  const float4* d = adopt_points(c, c->ds_out.p, mm, 32, c->dst);
  rc = build_target(c, d, mm);
  if (trace) {
    auto t_end = now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::fprintf(stderr, "[pcr trace] submap_build: %zu clouds, %zu pts -> %zu | upload %.3f transform %.3f downsample %.3f index %.3f ms (ncell %lld, ring %d)\n",
                 np, total, mm, ms(t_start, t_up), ms(t_up, t_tr), ms(t_tr, t_ds), ms(t_ds, t_end), c->loam_grid.g.ncell, c->loam_grid.max_ring);
  }
  return rc;
  PCR_API_END(c)
}

__global__ void seg_info_kernel(const uint32_t* keys, const uint32_t* seg_start, size_t nseg, int32_t* okeys, int32_t* ocounts) {
  size_t v = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (v >= nseg) return;
  okeys[v] = int32_t(keys[seg_start[v]]);
  ocounts[v] = int32_t(seg_start[v + 1] - seg_start[v]);
}

extern "C" int pcr_debug_voxel(pcr_ctx* c, int32_t* keys, int32_t* out_keys, int32_t* out_counts, int32_t grid[9]) {
  PCR_API_BEGIN(c)
  if (grid)
    for (int a = 0; a < 3; a++) { grid[a] = c->ds_grid.min_b[a]; grid[3 + a] = c->ds_grid.div_b[a]; grid[6 + a] = c->ds_grid.mul[a]; }
  if (c->ds_n == 0) return PCR_OK;
  if (c->ds_overflow) {
    if (keys) for (size_t i = 0; i < c->ds_n; i++) keys[i] = -1;
    return PCR_OK;
  }
  if (keys) PCR_CUDA_CHECK(cudaMemcpy(keys, c->ks_ds.keys_unsorted, c->ds_n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if ((out_keys || out_counts) && c->ds_m) {
    DevBuf<int32_t> tmp;
    tmp.ensure(c->ds_m * 2);
    seg_info_kernel<<<unsigned((c->ds_m + 127) / 128), 128, 0, c->stream>>>(c->ks_ds.keys, c->ks_ds.seg_start.p, c->ds_m, tmp.p, tmp.p + c->ds_m);
    PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (out_keys) PCR_CUDA_CHECK(cudaMemcpy(out_keys, tmp.p, c->ds_m * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (out_counts) PCR_CUDA_CHECK(cudaMemcpy(out_counts, tmp.p + c->ds_m, c->ds_m * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  return PCR_OK;
  PCR_API_END(c)
}

// ---- target blob (multi-GPU broadcast) ----------------------------------------------------------------------------
namespace {
struct BlobHeader {
  uint64_t magic;
  int32_t method;
  int32_t pad;
  uint64_t n_target;
  uint64_t sizes[8];  // byte sizes of the sections that follow (each 256-byte aligned)
  GridSpec g;
  double d[8];
  int32_t i[16];
  float f[16];
  uint64_t u[4];
  uint64_t params_hash;  // of the build parameters the index depends on (an index built with other parameters is refused)
  uint64_t reserved;
};
constexpr uint64_t kMagic = 0x33505242323030ull;  // "PCRB200" v3

// FNV-1a over the parameters a built index depends on
uint64_t params_hash(const pcr_params& p) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* v, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(v);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  };
  mix(&p.method, sizeof(p.method));
  switch (p.method) {
    case PCR_LOAM: mix(&p.loam_max_knn_d2, sizeof(p.loam_max_knn_d2)); break;
    case PCR_NDT:
      mix(&p.ndt_resolution, sizeof(p.ndt_resolution)); mix(&p.ndt_min_points, sizeof(p.ndt_min_points));
      mix(&p.ndt_eig_mult, sizeof(p.ndt_eig_mult)); mix(&p.ndt_outlier_ratio, sizeof(p.ndt_outlier_ratio));
      break;
    case PCR_VGICP: mix(&p.vgicp_resolution, sizeof(p.vgicp_resolution)); mix(&p.vgicp_k, sizeof(p.vgicp_k)); break;
  }
  return h;
}

// a blob is only trusted after every section has been checked against the blob length and against the grid / counts
// the header itself states (a truncated, corrupt or older-layout file must not drive copies or allocations)
const char* validate_blob(const BlobHeader& h, size_t bytes) {
  size_t off = (sizeof(BlobHeader) + 255) & ~size_t(255);
  for (int k = 0; k < 8; k++) {
    if (h.sizes[k] > bytes) return "section larger than the blob";
    const size_t a = (size_t(h.sizes[k]) + 255) & ~size_t(255);
    if (off + a > bytes || off + a < off) return "sections run past the end of the blob";
    off += a;
  }
  auto grid_ok = [](const GridSpec& g) {
    if (g.ncell <= 0 || g.ncell > (1ll << 29)) return false;
    for (int a = 0; a < 3; a++)
      if (g.div_b[a] <= 0 || g.max_b[a] - g.min_b[a] + 1 != g.div_b[a] || !(g.leaf[a] > 0.f)) return false;
    if ((long long)g.div_b[0] * g.div_b[1] * g.div_b[2] != g.ncell) return false;
    return g.mul[0] == 1 && g.mul[1] == g.div_b[0] && (long long)g.mul[2] == (long long)g.div_b[0] * g.div_b[1];
  };
  if (h.method == PCR_LOAM) {
    if (h.i[0] == 0) return (h.sizes[0] || h.sizes[1]) ? "unbuilt LOAM index with sections" : nullptr;
    if (!grid_ok(h.g)) return "LOAM grid inconsistent";
    if (h.sizes[0] != h.n_target * sizeof(float4) || h.sizes[0] > (size_t(1) << 36)) return "LOAM point section does not match n_target";
    if (h.sizes[1] != (size_t(h.g.ncell) + 1) * sizeof(int32_t)) return "LOAM start table does not match the grid";
    if (h.i[1] != 1 && h.i[1] != 2) return "LOAM ring count out of range";
  } else if (h.method == PCR_NDT) {
    if (h.i[0] != 0 || h.i[1] == 0) return nullptr;  // overflow / no leaves: nothing follows
    if (h.i[1] < 0 || !grid_ok(h.g)) return "NDT grid inconsistent";
    const size_t L = size_t(h.i[1]);
    const size_t want[8] = {L * sizeof(NdtLeafRec), size_t(h.g.ncell) * sizeof(int32_t), L * 24, L * 72, L * 72, L * 4, L * 4, L * sizeof(float4)};
    for (int k = 0; k < 8; k++)
      if (h.sizes[k] != want[k]) return "NDT section size does not match the leaf count / grid";
  } else if (h.method == PCR_VGICP) {
    if (h.sizes[0] == 0) return nullptr;
    const size_t n = size_t(h.n_target), cap = size_t(h.u[0]), nvox = size_t(h.u[3]);
    if (h.u[1] != h.n_target || cap < 1024 || (cap & (cap - 1)) || cap > (size_t(1) << 33)) return "VGICP hash capacity inconsistent";
    if (h.sizes[0] != n * sizeof(float4) || h.sizes[1] != size_t(kKnnLevels) * cap * sizeof(uint4) || h.sizes[2] != n * 6 * sizeof(double))
      return "VGICP kNN sections do not match n_target";
    if (nvox) {
      if (h.u[2] == 0 || h.u[2] > (1ull << 29)) return "VGICP voxel grid too large";
      long long nc = 1;
      for (int a = 0; a < 3; a++) { if (h.i[4 + a] <= 0) return "VGICP voxel grid inconsistent"; nc *= h.i[4 + a]; }
      if ((unsigned long long)nc != h.u[2]) return "VGICP voxel grid inconsistent";
      if (h.sizes[3] != nvox * sizeof(VoxelRec) || h.sizes[4] != nvox * sizeof(int32_t) || h.sizes[5] != size_t(h.u[2]) * sizeof(int32_t))
        return "VGICP voxel sections do not match the voxel count / grid";
    }
  } else {
    return "unknown method";
  }
  return nullptr;
}
size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct Section { const void* src; void* dst_holder; size_t bytes; };

// describes the device arrays making up the built target of ctx c
int collect_sections(pcr_ctx* c, BlobHeader& h, const void* ptrs[8]) {
  memset(&h, 0, sizeof(h));
  h.magic = kMagic;
  h.method = c->prm.method;
  h.n_target = c->n_target;
  h.params_hash = params_hash(c->prm);
  for (int k = 0; k < 8; k++) ptrs[k] = nullptr;
  switch (c->prm.method) {
    case PCR_LOAM:
      h.g = c->loam_grid.g;
      h.i[0] = c->loam_grid.built ? 1 : 0;
      h.i[1] = c->loam_grid.max_ring;
      if (c->loam_grid.built) {
        ptrs[0] = c->loam_grid.pts.p; h.sizes[0] = c->loam_grid.n * sizeof(float4);
        ptrs[1] = c->loam_grid.start.p; h.sizes[1] = (size_t(c->loam_grid.g.ncell) + 1) * sizeof(int32_t);
      }
      break;
    case PCR_NDT:
      h.g = c->ndt.g;
      h.i[0] = c->ndt.overflow ? 1 : 0;
      h.i[1] = int32_t(c->ndt.nleaves);
      h.d[0] = c->ndt.d1; h.d[1] = c->ndt.d2; h.d[2] = c->ndt.d3; h.d[3] = c->ndt.resolution;
      if (!c->ndt.overflow && c->ndt.nleaves) {
        const size_t L = c->ndt.nleaves;
        ptrs[0] = c->ndt.recs.p; h.sizes[0] = L * sizeof(NdtLeafRec);
        ptrs[1] = c->ndt.table.p; h.sizes[1] = size_t(c->ndt.g.ncell) * sizeof(int32_t);
        ptrs[2] = c->ndt.mean.p; h.sizes[2] = L * 3 * sizeof(double);
        ptrs[3] = c->ndt.icov.p; h.sizes[3] = L * 9 * sizeof(double);
        ptrs[4] = c->ndt.cov.p; h.sizes[4] = L * 9 * sizeof(double);
        ptrs[5] = c->ndt.keys.p; h.sizes[5] = L * sizeof(int32_t);
        ptrs[6] = c->ndt.npts.p; h.sizes[6] = L * sizeof(int32_t);
        ptrs[7] = c->ndt.centroids.p; h.sizes[7] = L * sizeof(float4);
      }
      break;
    case PCR_VGICP: {
      const VgicpTarget& t = c->vg;
      const MortonGrid& mg = t.grid;
      h.i[0] = t.built ? 1 : 0;
      for (int a = 0; a < 3; a++) { h.i[1 + a] = t.cmin[a]; h.i[4 + a] = t.cdim[a]; h.i[7 + a] = mg.dim0[a]; h.f[a] = mg.mn[a]; h.f[3 + a] = mg.smax[a]; }
      h.i[10] = mg.built ? 1 : 0;
      h.f[6] = mg.h0; h.f[7] = mg.inv_h0;
      h.d[0] = t.resolution; h.d[1] = mg.slack0;
      h.u[0] = mg.cap; h.u[1] = mg.n; h.u[2] = uint64_t(t.ncell); h.u[3] = t.nvox;
      if (mg.built && t.n) {
        ptrs[0] = mg.pts.p; h.sizes[0] = mg.n * sizeof(float4);
        ptrs[1] = mg.tables.p; h.sizes[1] = size_t(kKnnLevels) * mg.cap * sizeof(uint4);
        ptrs[2] = t.covs.p; h.sizes[2] = t.n * 6 * sizeof(double);
        if (t.nvox) {
          ptrs[3] = t.vox.p; h.sizes[3] = t.nvox * sizeof(VoxelRec);
          ptrs[4] = t.vox_key.p; h.sizes[4] = t.nvox * sizeof(int32_t);
          ptrs[5] = t.table.p; h.sizes[5] = size_t(t.ncell) * sizeof(int32_t);
        }
      }
      break;
    }
    default:
      return PCR_ERR_UNSUPPORTED;
  }
  return PCR_OK;
}
}  // namespace

extern "C" int pcr_target_blob_size(pcr_ctx* c, size_t* bytes) {
  PCR_API_BEGIN(c)
  if (!bytes) return PCR_ERR_INVALID;
  if (!c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no target set");
  BlobHeader h;
  const void* ptrs[8];
  int rc = collect_sections(c, h, ptrs);
  if (rc) return fail(c, rc, "target export: unknown method");
  size_t total = align256(sizeof(BlobHeader));
  for (int k = 0; k < 8; k++) total += align256(h.sizes[k]);
  *bytes = total;
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_target_export(pcr_ctx* c, void* dev_blob, size_t cap) {
  PCR_API_BEGIN(c)
  if (!c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no target set");
  BlobHeader h;
  const void* ptrs[8];
  int rc = collect_sections(c, h, ptrs);
  if (rc) return fail(c, rc, "target export: unknown method");
  size_t off = align256(sizeof(BlobHeader));
  unsigned char* base = static_cast<unsigned char*>(dev_blob);
  for (int k = 0; k < 8; k++) {
    if (off + align256(h.sizes[k]) > cap) return fail(c, PCR_ERR_INVALID, "blob capacity too small");
    if (h.sizes[k]) PCR_CUDA_CHECK(cudaMemcpyAsync(base + off, ptrs[k], h.sizes[k], cudaMemcpyDeviceToDevice, c->stream));
    off += align256(h.sizes[k]);
  }
  PCR_CUDA_CHECK(cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_target_import(pcr_ctx* c, const void* dev_blob, size_t bytes) {
  PCR_API_BEGIN(c)
  if (!dev_blob || bytes < sizeof(BlobHeader)) return fail(c, PCR_ERR_INVALID, "bad blob");
  BlobHeader h;
  PCR_CUDA_CHECK(cudaMemcpy(&h, dev_blob, sizeof(h), cudaMemcpyDeviceToHost));
  if (h.magic != kMagic) return fail(c, PCR_ERR_INVALID, "not a target index blob of this library version");
  if (h.method != c->prm.method) return fail(c, PCR_ERR_INVALID, "blob does not match this context's method");
  if (h.params_hash != params_hash(c->prm)) return fail(c, PCR_ERR_INVALID, "index was built with other parameters than this context's");
  if (const char* why = validate_blob(h, bytes)) return fail(c, PCR_ERR_INVALID, why);
  const unsigned char* base = static_cast<const unsigned char*>(dev_blob);
  size_t off = align256(sizeof(BlobHeader));
  auto take = [&](void* dst, int k) {
    if (h.sizes[k]) PCR_CUDA_CHECK(cudaMemcpyAsync(dst, base + off, h.sizes[k], cudaMemcpyDeviceToDevice, c->stream));
    off += align256(h.sizes[k]);
  };
  c->has_target = false;
  c->n_target = h.n_target;
  if (h.method == PCR_LOAM) {
    c->loam_grid.g = h.g;
    c->loam_grid.n = h.sizes[0] / sizeof(float4);
    c->loam_grid.built = h.i[0] != 0;
    c->loam_grid.max_ring = h.i[1];
    if (c->loam_grid.built) {
      c->loam_grid.pts.ensure(c->loam_grid.n);
      c->loam_grid.start.ensure(size_t(h.g.ncell) + 1);
      c->loam_grid.has_start = true;
      take(c->loam_grid.pts.p, 0);
      take(c->loam_grid.start.p, 1);
    }
  } else if (h.method == PCR_NDT) {
    NdtTarget& t = c->ndt;
    t.g = h.g;
    t.overflow = h.i[0] != 0;
    t.nleaves = size_t(h.i[1]);
    t.d1 = h.d[0]; t.d2 = h.d[1]; t.d3 = h.d[2]; t.resolution = float(h.d[3]);
    if (!t.overflow && t.nleaves) {
      const size_t L = t.nleaves;
      t.recs.ensure(L); t.table.ensure(size_t(h.g.ncell)); t.mean.ensure(L * 3); t.icov.ensure(L * 9); t.cov.ensure(L * 9);
      t.keys.ensure(L); t.npts.ensure(L); t.centroids.ensure(L);
      take(t.recs.p, 0); take(t.table.p, 1); take(t.mean.p, 2); take(t.icov.p, 3); take(t.cov.p, 4); take(t.keys.p, 5); take(t.npts.p, 6);
      take(t.centroids.p, 7);
    }
    t.built = true;
  } else if (h.method == PCR_VGICP) {
    VgicpTarget& t = c->vg;
    MortonGrid& mg = t.grid;
    t.built = false;
    t.n = size_t(h.n_target);
    t.resolution = h.d[0];
    mg.slack0 = h.d[1];
    for (int a = 0; a < 3; a++) { t.cmin[a] = h.i[1 + a]; t.cdim[a] = h.i[4 + a]; mg.dim0[a] = h.i[7 + a]; mg.mn[a] = h.f[a]; mg.smax[a] = h.f[3 + a]; }
    mg.h0 = h.f[6]; mg.inv_h0 = h.f[7];
    mg.cap = uint32_t(h.u[0]); mg.n = size_t(h.u[1]); t.ncell = (long long)h.u[2]; t.nvox = size_t(h.u[3]);
    mg.built = h.i[10] != 0;
    if (h.sizes[0]) {
      mg.pts.ensure(mg.n); mg.tables.ensure(size_t(kKnnLevels) * mg.cap); t.covs.ensure(t.n * 6);
      take(mg.pts.p, 0); take(mg.tables.p, 1); take(t.covs.p, 2);
      if (t.nvox) {
        t.vox.ensure(t.nvox); t.vox_key.ensure(t.nvox); t.table.ensure(size_t(t.ncell));
        take(t.vox.p, 3); take(t.vox_key.p, 4); take(t.table.p, 5);
      }
    }
    t.built = h.i[0] != 0;
    c->has_last = false;
  } else {
    return fail(c, PCR_ERR_UNSUPPORTED, "unknown method in the target blob");
  }
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->has_target = true;
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_host_register(const void* p, size_t bytes) {
  if (!p || bytes == 0) return PCR_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return PCR_OK; }
  if (e != cudaSuccess) {
    cudaGetLastError();
    pcr_log(3, std::string("cudaHostRegister failed: ") + cudaGetErrorString(e));
    return PCR_ERR_CUDA;
  }
  return PCR_OK;
}

extern "C" int pcr_host_unregister(const void* p) {
  if (!p) return PCR_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(const_cast<void*>(p));
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorHostMemoryNotRegistered ? PCR_OK : PCR_ERR_CUDA; }
  return PCR_OK;
}

extern "C" int pcr_trim_device_cache(size_t* freed_bytes) {
  const size_t before = DevPool::cached_bytes();
  cudaDeviceSynchronize();
  DevPool::trim();
  if (freed_bytes) *freed_bytes = before - DevPool::cached_bytes();  // buffers of a slab with a sibling in use stay parked
  return PCR_OK;
}

// ---- several GPUs in one process (loc.cpp mode) --------------------------------------------------------------------
struct pcr_multi {
  std::vector<pcr_ctx*> ctx;
  std::vector<int> dev;
  std::string err;
  size_t blob_bytes = 0;
  double copy_ms = 0.0;
};

extern "C" int pcr_multi_create(const pcr_params* p, const int32_t* devices, size_t n_devices, pcr_multi** out) {
  if (!p || !devices || !out || n_devices == 0) return PCR_ERR_INVALID;
  *out = nullptr;
  pcr_multi* m = new pcr_multi();
  for (size_t k = 0; k < n_devices; k++) {
    pcr_params q = *p;
    q.device = devices[k];
    pcr_ctx* c = nullptr;
    const int rc = pcr_create(&q, &c);
    if (rc) {
      for (pcr_ctx* x : m->ctx) pcr_destroy(x);
      delete m;
      return rc;  // pcr_last_error(NULL) has the text
    }
    m->ctx.push_back(c);
    m->dev.push_back(devices[k]);
  }
  *out = m;
  return PCR_OK;
}

extern "C" void pcr_multi_destroy(pcr_multi* m) {
  if (!m) return;
  for (pcr_ctx* c : m->ctx) pcr_destroy(c);
  delete m;
}

extern "C" const char* pcr_multi_last_error(const pcr_multi* m) { return m ? m->err.c_str() : ""; }

extern "C" int pcr_multi_get_broadcast(const pcr_multi* m, size_t* blob_bytes, double* copy_ms) {
  if (!m) return PCR_ERR_INVALID;
  if (blob_bytes) *blob_bytes = m->blob_bytes;
  if (copy_ms) *copy_ms = m->copy_ms;
  return PCR_OK;
}

extern "C" int pcr_multi_set_target(pcr_multi* m, const void* pts, size_t n, size_t stride) {
  if (!m) return PCR_ERR_INVALID;
  pcr_ctx* c0 = m->ctx[0];
  int rc = pcr_set_target(c0, pts, n, stride);
  if (rc) { m->err = pcr_last_error(c0); return rc; }
  m->blob_bytes = 0;
  m->copy_ms = 0.0;
  if (m->ctx.size() == 1) return PCR_OK;
  size_t bytes = 0;
  rc = pcr_target_blob_size(c0, &bytes);
  if (rc) { m->err = pcr_last_error(c0); return rc; }
  void* blob0 = nullptr;
  std::vector<void*> blobs(m->ctx.size(), nullptr);
  auto cleanup = [&]() {
    for (size_t k = 0; k < blobs.size(); k++)
      if (blobs[k]) { cudaSetDevice(m->dev[k]); cudaFree(blobs[k]); }
  };
  try {
    PCR_CUDA_CHECK(cudaSetDevice(m->dev[0]));
    PCR_CUDA_CHECK(cudaMalloc(&blob0, bytes));
    blobs[0] = blob0;
    rc = pcr_target_export(c0, blob0, bytes);
    if (rc) { m->err = pcr_last_error(c0); cleanup(); return rc; }
    // all copies are queued before the first one is waited for: NVSwitch gives every peer the full link rate
    for (size_t k = 1; k < m->ctx.size(); k++) {
      PCR_CUDA_CHECK(cudaSetDevice(m->dev[k]));
      PCR_CUDA_CHECK(cudaMalloc(&blobs[k], bytes));
      if (m->dev[k] != m->dev[0]) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->dev[k], m->dev[0]);
        if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[0], 0); if (e != cudaSuccess) cudaGetLastError(); }  // already enabled is fine
      }
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t k = 1; k < m->ctx.size(); k++) {
      PCR_CUDA_CHECK(cudaSetDevice(m->dev[k]));
      PCR_CUDA_CHECK(cudaMemcpyPeerAsync(blobs[k], m->dev[k], blob0, m->dev[0], bytes, m->ctx[k]->stream));
    }
    for (size_t k = 1; k < m->ctx.size(); k++) {
      PCR_CUDA_CHECK(cudaSetDevice(m->dev[k]));
      PCR_CUDA_CHECK(cudaStreamSynchronize(m->ctx[k]->stream));
    }
    m->copy_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    m->blob_bytes = bytes;
    for (size_t k = 1; k < m->ctx.size(); k++) {
      rc = pcr_target_import(m->ctx[k], blobs[k], bytes);
      if (rc) { m->err = pcr_last_error(m->ctx[k]); cleanup(); return rc; }
    }
  } catch (const std::exception& e) {
    m->err = e.what();
    cudaGetLastError();
    cleanup();
    return PCR_ERR_CUDA;
  }
  cleanup();
  return PCR_OK;
}

extern "C" int pcr_multi_batch_align(pcr_multi* m, const void* src, const size_t* offsets, size_t n_scans, size_t stride, double* T,
                                     int32_t* converged) {
  if (!m || !offsets || !T) return PCR_ERR_INVALID;
  const size_t nd = m->ctx.size();
  std::vector<int> rcs(nd, PCR_OK);
  std::vector<std::thread> th;
  auto shard = [&](size_t k) {  // contiguous blocks, the first n % nd shards one scan longer (multigpu.shard)
    const size_t base = n_scans / nd, rem = n_scans % nd;
    const size_t lo = k * base + std::min(k, rem);
    return std::make_pair(lo, lo + base + (k < rem ? 1 : 0));
  };
  for (size_t k = 0; k < nd; k++) {
    th.emplace_back([&, k]() {
      const auto r = shard(k);
      if (r.second == r.first) return;
      rcs[k] = pcr_batch_align(m->ctx[k], src, offsets + r.first, r.second - r.first, stride, T + 16 * r.first, converged ? converged + r.first : nullptr);
    });
  }
  for (auto& t : th) t.join();
  for (size_t k = 0; k < nd; k++)
    if (rcs[k]) { m->err = pcr_last_error(m->ctx[k]); return rcs[k]; }
  return PCR_OK;
}

// ---- on-disk index cache + PCD reader -----------------------------------------------------------
namespace {
struct FileHeader {
  char magic[8];       // "PCRIDX02"
  uint64_t blob_bytes;
};
}  // namespace

extern "C" int pcr_target_save(pcr_ctx* c, const char* path) {
  PCR_API_BEGIN(c)
  if (!path) return fail(c, PCR_ERR_INVALID, "null path");
  size_t bytes = 0;
  int rc = pcr_target_blob_size(c, &bytes);
  if (rc) return rc;
  DevBuf<unsigned char> blob;
  blob.ensure(bytes);
  rc = pcr_target_export(c, blob.p, bytes);
  if (rc) return rc;
  std::vector<unsigned char> host(bytes);
  PCR_CUDA_CHECK(cudaMemcpy(host.data(), blob.p, bytes, cudaMemcpyDeviceToHost));
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(c, PCR_ERR_INVALID, "cannot open index file for writing");
  FileHeader h;
  std::memcpy(h.magic, "PCRIDX02", 8);
  h.blob_bytes = bytes;
  const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(host.data(), 1, bytes, f) == bytes;
  std::fclose(f);
  if (!ok) return fail(c, PCR_ERR_INVALID, "short write of the index file");
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_target_load(pcr_ctx* c, const char* path) {
  PCR_API_BEGIN(c)
  if (!path) return fail(c, PCR_ERR_INVALID, "null path");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(c, PCR_ERR_INVALID, "cannot open index file");
  FileHeader h;
  if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "PCRIDX02", 8) != 0) { std::fclose(f); return fail(c, PCR_ERR_INVALID, "not an index file"); }
  std::fseek(f, 0, SEEK_END);
  const long fsize = std::ftell(f);
  std::fseek(f, long(sizeof(h)), SEEK_SET);
  if (fsize < 0 || h.blob_bytes < sizeof(BlobHeader) || h.blob_bytes > size_t(fsize) - sizeof(h)) {
    std::fclose(f);
    return fail(c, PCR_ERR_INVALID, "truncated index file");
  }
  unsigned char* host = c->pin.ensure(h.blob_bytes);
  const bool ok = std::fread(host, 1, h.blob_bytes, f) == h.blob_bytes;
  std::fclose(f);
  if (!ok) return fail(c, PCR_ERR_INVALID, "truncated index file");
  DevBuf<unsigned char> blob;
  blob.ensure(h.blob_bytes);
  PCR_CUDA_CHECK(cudaMemcpyAsync(blob.p, host, h.blob_bytes, cudaMemcpyHostToDevice, c->stream));
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return pcr_target_import(c, blob.p, h.blob_bytes);
  PCR_API_END(c)
}

extern "C" int pcr_read_pcd(const char* path, void* out, size_t cap, size_t* n) {
  if (!path || !n) return PCR_ERR_INVALID;
  *n = 0;
  std::ifstream in(path, std::ios::binary);
  if (!in) return PCR_ERR_INVALID;
  std::vector<std::string> fields, types;
  std::vector<int> sizes, counts;
  size_t points = 0;
  std::string data_kind, line;
  while (std::getline(in, line)) {  // header: "KEY v1 v2 ..." lines up to DATA
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key, tok;
    ls >> key;
    if (key == "FIELDS") { while (ls >> tok) fields.push_back(tok); }
    else if (key == "SIZE") { while (ls >> tok) sizes.push_back(std::atoi(tok.c_str())); }
    else if (key == "TYPE") { while (ls >> tok) types.push_back(tok); }
    else if (key == "COUNT") { while (ls >> tok) counts.push_back(std::atoi(tok.c_str())); }
    else if (key == "POINTS") { ls >> points; }
    else if (key == "DATA") { ls >> data_kind; break; }
  }
  const size_t nf = fields.size();
  if (!nf || sizes.size() != nf || types.size() != nf || (data_kind != "ascii" && data_kind != "binary")) return PCR_ERR_UNSUPPORTED;
  if (counts.empty()) counts.assign(nf, 1);
  int off[4] = {-1, -1, -1, -1}, col[4] = {-1, -1, -1, -1};  // byte offset / ascii column of x, y, z, intensity
  int rec = 0, cols = 0;
  for (size_t k = 0; k < nf; k++) {
    const char* want[4] = {"x", "y", "z", "intensity"};
    for (int w = 0; w < 4; w++)
      if (fields[k] == want[w]) {
        if (sizes[k] != 4 || types[k] != "F" || counts[k] != 1) return PCR_ERR_UNSUPPORTED;
        off[w] = rec; col[w] = cols;
      }
    rec += sizes[k] * counts[k];
    cols += counts[k];
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) return PCR_ERR_UNSUPPORTED;
  *n = points;
  if (!out) return PCR_OK;
  float* o = static_cast<float*>(out);
  const size_t take = std::min(points, cap);
  if (data_kind == "binary") {
    std::vector<unsigned char> buf(static_cast<size_t>(rec) * 4096, 0);
    for (size_t base = 0; base < take; base += 4096) {
      const size_t m = std::min<size_t>(4096, take - base);
      in.read(reinterpret_cast<char*>(buf.data()), std::streamsize(m * size_t(rec)));
      if (size_t(in.gcount()) != m * size_t(rec)) return PCR_ERR_INVALID;
      for (size_t i = 0; i < m; i++) {
        const unsigned char* r = buf.data() + i * size_t(rec);
        float* q = o + (base + i) * 8;
        std::memcpy(q + 0, r + off[0], 4); std::memcpy(q + 1, r + off[1], 4); std::memcpy(q + 2, r + off[2], 4);
        q[3] = 1.0f;
        q[4] = 0.0f;
        if (off[3] >= 0) std::memcpy(q + 4, r + off[3], 4);
        q[5] = q[6] = q[7] = 0.0f;
      }
    }
  } else {
    std::vector<double> v(static_cast<size_t>(cols), 0.0);
    for (size_t i = 0; i < take; i++) {
      for (int k = 0; k < cols; k++)
        if (!(in >> v[size_t(k)])) return PCR_ERR_INVALID;
      float* q = o + i * 8;
      q[0] = float(v[size_t(col[0])]); q[1] = float(v[size_t(col[1])]); q[2] = float(v[size_t(col[2])]);
      q[3] = 1.0f;
      q[4] = col[3] >= 0 ? float(v[size_t(col[3])]) : 0.0f;
      q[5] = q[6] = q[7] = 0.0f;
    }
  }
  return PCR_OK;
}

extern "C" int pcr_static_map_load(pcr_ctx* c, const char* pcd_path, float leaf, size_t* m) {
  PCR_API_BEGIN(c)
  if (!pcd_path || !m || !(leaf > 0.f)) return fail(c, PCR_ERR_INVALID, "bad arguments");
  size_t n = 0;
  int rc = pcr_read_pcd(pcd_path, nullptr, 0, &n);
  if (rc) return fail(c, rc, "can't load globalmap (unreadable or unsupported PCD)");  // MapManager.cpp:68-73
  unsigned char* host = c->pin.ensure(std::max<size_t>(n, 1) * 32);
  rc = pcr_read_pcd(pcd_path, host, n, &n);
  if (rc) return fail(c, rc, "can't load globalmap (truncated PCD)");
  const float4* d = upload_points(c, host, n, 32, c->raw_src, c->ds_in);
  c->ds_out.ensure(std::max<size_t>(n, 1) * 32);
  rc = downsample_packed(c, d, n, leaf, c->ds_out.p, n, m);   // pcp::voxelDownSample(mSubmap, mGridSize), MapManager.cpp:77
  if (rc) return rc;
  const float4* t = adopt_points(c, c->ds_out.p, *m, 32, c->dst);
  return build_target(c, t, *m);
  PCR_API_END(c)
}

// ---- ScanContext ---------------------------------------------------------------------------------
extern "C" int pcr_scancontext_make(pcr_ctx* c, const void* pts, const size_t* offsets, size_t n_clouds, size_t stride, float lidar_height,
                                    double* desc, double* ring_key, double* sector_key) {
  PCR_API_BEGIN(c)
  if (!offsets || stride < 12 || stride % 4) return fail(c, PCR_ERR_INVALID, "bad arguments");
  if (n_clouds == 0) return PCR_OK;
  const size_t n = offsets[n_clouds] - offsets[0];
  if (n && !pts) return fail(c, PCR_ERR_INVALID, "null cloud");
  const unsigned char* base = static_cast<const unsigned char*>(pts) + offsets[0] * stride;
  const float4* d = upload_points(c, base, n, stride, c->raw_src, c->ds_in);
  DevBuf<uint32_t> offs;
  offs.ensure(n_clouds + 1);
  std::vector<uint32_t> ho(n_clouds + 1);
  for (size_t k = 0; k <= n_clouds; k++) ho[k] = uint32_t(offsets[k] - offsets[0]);
  PCR_CUDA_CHECK(cudaMemcpyAsync(offs.p, ho.data(), (n_clouds + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  DevBuf<double> out;
  const size_t per = size_t(kScSize) + kScRings + kScSectors;
  out.ensure(n_clouds * per);
  double* dd = out.p;
  double* dr = dd + n_clouds * kScSize;
  double* dsk = dr + n_clouds * kScRings;
  scancontext_make(d, offs.p, int(n_clouds), lidar_height, dd, dr, dsk, c->stream);
  if (desc) PCR_CUDA_CHECK(cudaMemcpyAsync(desc, dd, n_clouds * kScSize * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (ring_key) PCR_CUDA_CHECK(cudaMemcpyAsync(ring_key, dr, n_clouds * kScRings * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (sector_key) PCR_CUDA_CHECK(cudaMemcpyAsync(sector_key, dsk, n_clouds * kScSectors * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PCR_CUDA_CHECK(cudaGetLastError());
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_scancontext_distance(pcr_ctx* c, const double* descs, size_t n_desc, const int32_t* pairs, size_t n_pairs, float search_ratio,
                                        int32_t sector_key_align, double* dist, int32_t* shift) {
  PCR_API_BEGIN(c)
  if (n_pairs == 0) return PCR_OK;
  if (!descs || !pairs || !dist || !shift || n_desc == 0) return fail(c, PCR_ERR_INVALID, "bad arguments");
  for (size_t k = 0; k < 2 * n_pairs; k++)
    if (pairs[k] < 0 || size_t(pairs[k]) >= n_desc) return fail(c, PCR_ERR_INVALID, "pair index out of range");
  const int radius = int(std::lround(0.5 * double(search_ratio) * double(kScSectors)));  // SEARCH_RADIUS (:127)
  if (radius < 0 || radius > 30) return fail(c, PCR_ERR_INVALID, "search ratio out of range");
  DevBuf<double> dd, ddist;
  DevBuf<int2> dp;
  DevBuf<int32_t> dshift;
  dd.ensure(n_desc * kScSize); dp.ensure(n_pairs); ddist.ensure(n_pairs); dshift.ensure(n_pairs);
  PCR_CUDA_CHECK(cudaMemcpyAsync(dd.p, descs, n_desc * kScSize * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  PCR_CUDA_CHECK(cudaMemcpyAsync(dp.p, pairs, n_pairs * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
  scancontext_distance(dd.p, dp.p, int(n_pairs), radius, sector_key_align ? 1 : 0, ddist.p, dshift.p, c->stream);
  PCR_CUDA_CHECK(cudaMemcpyAsync(dist, ddist.p, n_pairs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PCR_CUDA_CHECK(cudaMemcpyAsync(shift, dshift.p, n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PCR_CUDA_CHECK(cudaGetLastError());
  return PCR_OK;
  PCR_API_END(c)
}

// ---- parity / introspection ---------------------------------------------------------------------------------------
extern "C" int pcr_loam_linearize(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double T[16], int32_t* knn_idx,
                                  int32_t* status, double JtJ[36], double JtE[6], int64_t* n_acc) {
  PCR_API_BEGIN(c)
  if (c->prm.method != PCR_LOAM) return fail(c, PCR_ERR_INVALID, "not a LOAM context");
  if (!c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no target set");
  const float4* d = upload_points(c, src, ns, stride, c->raw_src, c->src);
  return c->loam.linearize(d, ns, c->loam_grid, loam_params(c->prm), T, knn_idx, status, JtJ, JtE, n_acc, c->stream);
  PCR_API_END(c)
}

extern "C" int pcr_loam_last_shape(pcr_ctx* c, int32_t shape[3]) {
  if (!c || !shape) return PCR_ERR_INVALID;
  shape[0] = c->loam.last_lpq; shape[1] = c->loam.last_tile; shape[2] = c->loam.last_split ? 1 : 0;
  return PCR_OK;
}

extern "C" int pcr_loam_get_logs(pcr_ctx* c, pcr_loam_iter_log* logs, int32_t cap, int32_t* n) {
  if (!c || !n) return PCR_ERR_INVALID;
  int cnt = std::min(c->loam.last_log_count, cap);
  if (logs && c->loam.h_logs.p)
    for (int i = 0; i < cnt; i++) logs[i] = c->loam.h_logs.p[i];
  *n = cnt;
  return PCR_OK;
}

extern "C" int pcr_ndt_num_leaves(pcr_ctx* c, size_t* n, int32_t grid[9]) {
  if (!c || !n) return PCR_ERR_INVALID;
  if (c->prm.method != PCR_NDT || !c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no NDT target");
  *n = c->ndt.overflow ? 0 : c->ndt.nleaves;
  if (grid)
    for (int a = 0; a < 3; a++) { grid[a] = c->ndt.g.min_b[a]; grid[3 + a] = c->ndt.g.max_b[a]; grid[6 + a] = c->ndt.g.div_b[a]; }
  return PCR_OK;
}

extern "C" int pcr_ndt_get_leaves(pcr_ctx* c, int32_t* keys, int32_t* npts, double* mean, double* cov, double* icov) {
  PCR_API_BEGIN(c)
  if (c->prm.method != PCR_NDT || !c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no NDT target");
  const size_t L = c->ndt.overflow ? 0 : c->ndt.nleaves;
  if (!L) return PCR_OK;
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (keys) PCR_CUDA_CHECK(cudaMemcpy(keys, c->ndt.keys.p, L * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (npts) PCR_CUDA_CHECK(cudaMemcpy(npts, c->ndt.npts.p, L * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (mean) PCR_CUDA_CHECK(cudaMemcpy(mean, c->ndt.mean.p, L * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  if (cov) PCR_CUDA_CHECK(cudaMemcpy(cov, c->ndt.cov.p, L * 9 * sizeof(double), cudaMemcpyDeviceToHost));
  if (icov) PCR_CUDA_CHECK(cudaMemcpy(icov, c->ndt.icov.p, L * 9 * sizeof(double), cudaMemcpyDeviceToHost));
  return PCR_OK;
  PCR_API_END(c)
}

static int ndt_eval_one(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double p[6], const float* Tf, int kind, int hess,
                        NdtEvalResult& out) {
  if (c->prm.method != PCR_NDT || !c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no NDT target");
  const float4* d = upload_points(c, src, ns, stride, c->raw_src, c->src);
  NdtEvalParams ep;
  memset(&ep, 0, sizeof(ep));
  float Tm[16];
  if (Tf) memcpy(Tm, Tf, sizeof(Tm)); else hm::ndt_pose_matrix_f32(p, Tm);
  memcpy(ep.Tf, Tm, sizeof(Tm));
  hm::ndt_angle_tables(p, ep.j_ang, ep.h_ang, ep.j_ang_d, ep.h_ang_d);
  ep.compute_hessian = hess; ep.kind = kind; ep.scan = 0; ep.pad = 0;
  c->ndtd.evaluate_one(d, ns, c->ndt, c->prm.ndt_search, ep, out, c->stream);
  return PCR_OK;
}

extern "C" int pcr_ndt_derivatives(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double p[6], const float* Tf,
                                   int32_t compute_hessian, double* score, double g[6], double H[36]) {
  PCR_API_BEGIN(c)
  NdtEvalResult r;
  int rc = ndt_eval_one(c, src, ns, stride, p, Tf, 0, compute_hessian ? 1 : 0, r);
  if (rc) return rc;
  if (score) *score = r.v[0];
  if (g) for (int i = 0; i < 6; i++) g[i] = r.v[1 + i];
  if (H) {
    int k = 7;
    for (int a = 0; a < 6; a++)
      for (int b = a; b < 6; b++) { H[a * 6 + b] = compute_hessian ? r.v[k] : 0.0; H[b * 6 + a] = H[a * 6 + b]; k++; }
  }
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_ndt_hessian(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double p[6], double H[36]) {
  PCR_API_BEGIN(c)
  NdtEvalResult r;
  int rc = ndt_eval_one(c, src, ns, stride, p, nullptr, 1, 1, r);
  if (rc) return rc;
  int k = 7;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) { H[a * 6 + b] = r.v[k]; H[b * 6 + a] = r.v[k]; k++; }
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_gicp_covariances(pcr_ctx* c, const void* pts, size_t n, size_t stride, int32_t k, double* covs, int32_t* knn_idx) {
  PCR_API_BEGIN(c)
  if (k < 1 || k > 32) return fail(c, PCR_ERR_INVALID, "k must be in [1, 32]");
  if (n == 0) return PCR_OK;
  const float4* d = upload_points(c, pts, n, stride, c->raw_src, c->src);
  c->has_last = false;
  MortonGrid& grid = c->vgd.src_grid;
  int rc = build_morton_grid(d, n, grid, c->bw, c->stream);
  if (rc == kRetryNonFinite) return fail(c, PCR_ERR_INVALID, "cloud holds NaN / Inf records (per-point outputs would not line up): strip them first");
  if (rc) return fail(c, rc, "grid too large");
  c->vgd.src_covs.ensure(n * 6);
  int32_t* dk = c->vgd.knn_dbg.ensure(n * size_t(k));
  gicp_covariances(d, n, grid, k, c->vgd.src_covs.p, dk, c->stream, nullptr, knn_idx != nullptr);
  PCR_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PCR_CUDA_CHECK(cudaGetLastError());
  if (covs) {
    std::vector<double> h(n * 6);
    PCR_CUDA_CHECK(cudaMemcpy(h.data(), c->vgd.src_covs.p, n * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) {
      const double* s = &h[i * 6];
      double* o = covs + i * 9;
      o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[1]; o[4] = s[3]; o[5] = s[4]; o[6] = s[2]; o[7] = s[4]; o[8] = s[5];
    }
  }
  if (knn_idx) PCR_CUDA_CHECK(cudaMemcpy(knn_idx, dk, n * size_t(k) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_vgicp_num_voxels(pcr_ctx* c, size_t* n) {
  if (!c || !n) return PCR_ERR_INVALID;
  if (c->prm.method != PCR_VGICP || !c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no VGICP target");
  *n = c->vg.nvox;
  return PCR_OK;
}

extern "C" int pcr_vgicp_get_voxels(pcr_ctx* c, int32_t* coords, int32_t* npts, double* mean, double* cov) {
  PCR_API_BEGIN(c)
  if (c->prm.method != PCR_VGICP || !c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no VGICP target");
  const size_t V = c->vg.nvox;
  if (!V) return PCR_OK;
  std::vector<VoxelRec> h(V);
  std::vector<int32_t> keys(V);
  PCR_CUDA_CHECK(cudaMemcpy(h.data(), c->vg.vox.p, V * sizeof(VoxelRec), cudaMemcpyDeviceToHost));
  PCR_CUDA_CHECK(cudaMemcpy(keys.data(), c->vg.vox_key.p, V * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (size_t v = 0; v < V; v++) {
    if (coords) {
      long long k = keys[v];
      coords[v * 3] = int32_t(k % c->vg.cdim[0]) + c->vg.cmin[0];
      coords[v * 3 + 1] = int32_t((k / c->vg.cdim[0]) % c->vg.cdim[1]) + c->vg.cmin[1];
      coords[v * 3 + 2] = int32_t(k / ((long long)c->vg.cdim[0] * c->vg.cdim[1])) + c->vg.cmin[2];
    }
    if (npts) npts[v] = int32_t(h[v].n + 0.5);
    if (mean) for (int a = 0; a < 3; a++) mean[v * 3 + a] = h[v].mean[a];
    if (cov) {
      const double* s = h[v].cov;
      double* o = cov + v * 9;
      o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[1]; o[4] = s[3]; o[5] = s[4]; o[6] = s[2]; o[7] = s[4]; o[8] = s[5];
    }
  }
  return PCR_OK;
  PCR_API_END(c)
}

extern "C" int pcr_vgicp_evaluate(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double T0[16], const double Ti[16],
                                  double* cost, double H[36], double b[6], int64_t* n_corr) {
  PCR_API_BEGIN(c)
  if (c->prm.method != PCR_VGICP || !c->has_target) return fail(c, PCR_ERR_NO_TARGET, "no VGICP target");
  const float4* d = upload_points(c, src, ns, stride, c->raw_src, c->src);
  c->has_last = false;  // c->src no longer holds the last aligned scan (getFitnessScore needs a new scan2Map)
  int rc = c->vgd.compute_source_covs(d, ns, c->prm.vgicp_k, c->ks, c->bw, c->stream);
  if (rc == kRetryNonFinite) return fail(c, PCR_ERR_INVALID, "source cloud holds NaN / Inf records: strip them first");
  if (rc) return fail(c, rc, "source covariance build failed");
  uint32_t* ho = c->vgd.h_offsets.ensure(2);
  ho[0] = 0; ho[1] = uint32_t(ns);
  c->vgd.offsets.ensure(2);
  PCR_CUDA_CHECK(cudaMemcpyAsync(c->vgd.offsets.p, ho, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  VgicpEvalParams* ep = c->vgd.h_params.ensure(1);
  memcpy(ep->T0, T0, sizeof(double) * 16);
  memcpy(ep->Ti, Ti, sizeof(double) * 16);
  ep->want_hb = (H || b) ? 1 : 0;
  ep->scan = 0;
  ep->pad[0] = ep->pad[1] = 0;
  c->vgd.evaluate(d, c->vgd.src_covs.p, c->vgd.offsets.p, ns, c->vg, 1, false, c->stream);
  const VgicpEvalResult& r = c->vgd.h_results.p[0];
  if (cost) *cost = r.v[0];
  if (H) {
    int k = 1;
    for (int a = 0; a < 6; a++)
      for (int q = a; q < 6; q++) { H[a * 6 + q] = r.v[k]; H[q * 6 + a] = r.v[k]; k++; }
  }
  if (b) for (int a = 0; a < 6; a++) b[a] = r.v[22 + a];
  if (n_corr) *n_corr = int64_t(r.v[28] + 0.5);
  return PCR_OK;
  PCR_API_END(c)
}
