/* pcr_cuda.h — C ABI of the B200-native PCR (PointCloudRegister) hot path.
 *
 * This is the drop-in boundary for SimpleSLAM's registration path: everything the reference's
 *   PCR::PointCloudRegister::scan2Map / getFitnessScore   (PCR/include/PCR/PointCloudRegister.hpp:34-35)
 * and its three config-selected back ends
 *   PCR::LoamRegister   (PCR/src/LoamRegister.cpp:99-223)
 *   PCR::NdtRegister    (PCR/src/NdtRegister.cpp:21-31      -> pclomp::NormalDistributionsTransform)
 *   PCR::VgicpRegister  (PCR/src/VgicpRegister.cpp:21-45    -> fast_gicp::FastVGICP)
 * plus the frontend's voxel downsample (common/pcp/pcp.hpp:15-28, frontend/src/LidarOdometry.cpp:170-171)
 * need from a GPU. Plain pointers and sizes only; no C++/torch/PCL/Eigen types cross this boundary.
 * The header-only C++ adaptor simpleslam_b200/cpp/PCR/ (one .hpp per register) re-creates the reference's class interface on top.
 *
 * Conventions
 *  - Clouds: array of records, `stride` bytes apart, float x,y,z at byte offset 0 and (if stride >= 20)
 *    float intensity at byte offset 16 — i.e. pcl::PointXYZI (32 B) or a bare float4/float3+pad (16 B).
 *  - Poses: double[16], column-major 4x4 (exactly Eigen::Isometry3d::matrix().data()), T_map<-scan,
 *    in = initial guess, out = refined pose.
 *  - Every function returns PCR_OK (0) or a negative error code; pcr_last_error() gives the text.
 *    There is NO CPU fallback: without an sm_100 device pcr_create fails with PCR_ERR_NO_DEVICE.
 *  - A context is single-threaded (one CUDA stream + device buffers per instance), like a reference
 *    register instance (SURVEY.md §8b "Threading"). Different contexts may be used from different threads.
 */
#ifndef PCR_CUDA_H
#define PCR_CUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PCR_OK 0
#define PCR_ERR_INVALID -1
#define PCR_ERR_NO_DEVICE -2
#define PCR_ERR_CUDA -3
#define PCR_ERR_NO_TARGET -4
#define PCR_ERR_GRID_TOO_LARGE -5
#define PCR_ERR_UNSUPPORTED -6

enum { PCR_LOAM = 0, PCR_NDT = 1, PCR_VGICP = 2 };                 /* cfg["frontend"]["pcr"] = loam|ndt|vgicp */
enum { PCR_NDT_KDTREE = 0, PCR_NDT_DIRECT26 = 1, PCR_NDT_DIRECT7 = 2, PCR_NDT_DIRECT1 = 3 }; /* pclomp::NeighborSearchMethod */
enum { PCR_LSQ_LM = 0, PCR_LSQ_GN = 1 };                           /* fast_gicp::LSQ_OPTIMIZER_TYPE */

typedef struct pcr_ctx pcr_ctx;

/* Tunables. pcr_default_params fills the reference's shipped values. */
typedef struct pcr_params {
  int32_t method;            /* PCR_LOAM | PCR_NDT | PCR_VGICP */
  int32_t device;            /* CUDA device ordinal */
  int32_t cores;             /* the reference's `cores` (PointCloudRegister.hpp:30-31): host threads that pack clouds uploaded from
                                pageable memory into pinned staging (from 8 threads up also clouds in pinned memory: half the PCIe bytes);
                                0 = plain cudaMemcpy. Never used for arithmetic. */
  /* LOAM (PCR/include/PCR/LoamRegister.hpp:30-40) */
  int32_t loam_max_iters;    /* 8 */
  float loam_max_knn_d2;     /* 1.0  (compared with a SQUARED distance, LoamRegister.cpp:59) */
  float loam_plane_thresh;   /* 0.2 */
  float loam_point_thresh;   /* 0.1 */
  float loam_pos_converge;   /* 5e-3 */
  float loam_rot_converge;   /* 5e-3 */
  /* NDT (NdtRegister.hpp:11, NdtRegister.cpp:12-18, ndt_omp_impl.hpp:49-51,71-74) */
  float ndt_resolution;      /* 1.0 */
  int32_t ndt_search;        /* PCR_NDT_DIRECT7 */
  int32_t ndt_max_iters;     /* 35 */
  double ndt_step_size;      /* 0.1 */
  double ndt_outlier_ratio;  /* 0.55 */
  double ndt_trans_eps;      /* 0.1 */
  int32_t ndt_min_points;    /* 6    (voxel_grid_covariance_omp.h:210) */
  double ndt_eig_mult;       /* 0.01 (voxel_grid_covariance_omp.h:211) */
  /* VGICP (VgicpRegister.cpp:13,21-28, fast_gicp_impl.hpp:16-20, lsq_registration_impl.hpp:11-18) */
  double vgicp_resolution;   /* 1.0 */
  int32_t vgicp_k;           /* 20 */
  int32_t vgicp_max_iters;   /* 64 (initForLC: 100) */
  int32_t vgicp_optimizer;   /* PCR_LSQ_LM */
  int32_t vgicp_lm_max_iters;/* 10 */
  double vgicp_rot_eps;      /* 2e-3 */
  double vgicp_trans_eps;    /* 5e-4 (initForLC: 1e-6) */
  double vgicp_lm_init_lambda; /* 1e-9 */
} pcr_params;

typedef struct pcr_stats {
  int32_t iterations;        /* LOAM: linearisations run; NDT: outer iterations; VGICP: LM/GN outer steps */
  int32_t evaluations;       /* kernel evaluations of the cost (LOAM: = iterations; NDT: computeDerivatives calls; VGICP: linearize + compute_error) */
  int32_t hessian_evals;     /* NDT computeHessian calls */
  int32_t converged;
  int64_t n_source;          /* points registered */
  int64_t n_target;
  int64_t n_residuals;       /* LOAM: accepted residuals of the last linearisation; VGICP: correspondences */
  int64_t kernel_launches;   /* hand-written kernels launched by the last align/scan2map call */
  int64_t n_pairs;           /* LOAM: map points examined by the pruned neighbour search (sum over iterations and scans); NDT: (point, leaf)
                                pairs evaluated (sum over evaluations); VGICP: correspondences (sum over evaluations) */
  int64_t n_point_evals;     /* source points pushed through the hot kernel, summed over its launches (roofline numerator) */
  int64_t n_index_reads;     /* spatial-index entries read by the hot kernel: LOAM x-row lookups (2 x 4 B each), NDT voxel-table
                                lookups (4 B each), VGICP voxel-table lookups (4 B each); summed over launches */
  double score;              /* NDT trans_probability; VGICP last cost */
  float ms_total;            /* device time of the last align call (CUDA events on the context's stream) */
  float ms_hot_kernel;       /* summed device time of the dominant correspondence/accumulation kernel */
  int32_t hot_kernel_launches;
  int32_t pad;
  /* second kernel worth a line of its own — VGICP: gicp_knn_kernel (exact k-NN behind the source / target covariances, the
   * largest share of a scan-to-scan registration); 0 for LOAM / NDT */
  float ms_aux_kernel;
  int32_t aux_kernel_launches;
  int64_t n_aux_items;        /* VGICP: k-NN queries answered */
} pcr_stats;

void pcr_default_params(int32_t method, pcr_params* p);
/* replaces: `new PCR::LoamRegister / NdtRegister / VgicpRegister` (frontend/src/LidarOdometry.cpp:44-53) */
int pcr_create(const pcr_params* p, pcr_ctx** out);
void pcr_destroy(pcr_ctx* c);
const char* pcr_last_error(const pcr_ctx* c); /* c may be NULL: error of the last failed pcr_create on this thread */
/* replaces: VgicpRegister::initForLC (PCR/src/VgicpRegister.cpp:21-28) */
int pcr_vgicp_init_for_lc(pcr_ctx* c);
/* 1 = record CUDA-event time of the hot kernel into pcr_stats (used by bench.py's roofline leg) */
int pcr_set_profiling(pcr_ctx* c, int enable);

/* Target (map) registration: upload + spatial index build. Replaces the per-call index build inside scan2Map:
 * nanoflann buildIndex (LoamRegister.cpp:110), VoxelGridCovariance::filter (ndt_omp.h:122-127,276-283),
 * FastVGICP::setInputTarget + calculate_covariances + create_voxelmap (fast_vgicp_impl.hpp:56-70,120-123). */
int pcr_set_target(pcr_ctx* c, const void* pts, size_t n, size_t stride);
/* same, `pts` is a DEVICE pointer on the context's device */
int pcr_set_target_device(pcr_ctx* c, const void* dev_pts, size_t n, size_t stride);

/* Align a scan to the current target. Replaces the body of scan2Map after the index build. */
int pcr_align(pcr_ctx* c, const void* src, size_t n, size_t stride, double T[16], int32_t* converged);
int pcr_align_device(pcr_ctx* c, const void* dev_src, size_t n, size_t stride, double T[16], int32_t* converged);
/* Exact reference semantics of PointCloudRegister::scan2Map(src, dst, res): (re)build the target index
 * from `dst`, then align. HOST buffers; all host<->device copies happen inside the call. */
int pcr_scan2map(pcr_ctx* c, const void* src, size_t ns, size_t sstride, const void* dst, size_t nm, size_t dstride,
                 double T[16], int32_t* converged);
/* Batched independent registrations against the current target (loc.cpp localisation pattern):
 * scans are concatenated, scan i = records [offsets[i], offsets[i+1]). T: 16*n_scans doubles in/out. */
int pcr_batch_align(pcr_ctx* c, const void* src, const size_t* offsets, size_t n_scans, size_t stride, double* T,
                    int32_t* converged);
int pcr_batch_align_device(pcr_ctx* c, const void* dev_src, const size_t* offsets, size_t n_scans, size_t stride, double* T,
                           int32_t* converged);
/* replaces: PointCloudRegister::getFitnessScore (VgicpRegister.cpp:42-45 -> pcl::Registration::getFitnessScore);
 * LOAM / NDT return 0 like the base class (PointCloudRegister.hpp:34). */
int pcr_fitness(pcr_ctx* c, double* score);
int pcr_get_stats(const pcr_ctx* c, pcr_stats* s);

/* Voxel-grid downsample. Replaces pcp::voxelDownSample -> pcl::VoxelGrid<PointXYZI>::filter
 * (common/pcp/pcp.hpp:15-28). `out` receives up to `cap` 32-byte PointXYZI records (x y z 1 | intensity 0 0 0),
 * voxels in ascending key order; *m = number of voxels. Works with any method's context. */
int pcr_voxel_downsample(pcr_ctx* c, const void* pts, size_t n, size_t stride, float leaf, void* out, size_t cap, size_t* m);
/* device in / device out variant (out: 32-byte records) */
int pcr_voxel_downsample_device(pcr_ctx* c, const void* dev_pts, size_t n, size_t stride, float leaf, void* dev_out, size_t cap,
                                size_t* m);

/* One frame of frontend::LidarOdometry::generateOdom (frontend/src/LidarOdometry.cpp:170-184): mVoxelGrid.filter(scan) followed by
 * mPcr->scan2Map(downsampled, submap, pose) against the resident target (pcr_set_target / pcr_submap_build). Same result as
 * pcr_voxel_downsample + pcr_align, without moving the downsampled scan to the host and back. m (nullable): points left. */
int pcr_downsample_align(pcr_ctx* c, const void* scan, size_t n, size_t stride, float leaf, double T[16], int32_t* converged, size_t* m);


/* Submap assembly on the device. Replaces the body of MapManager::updateMap (frontend/src/MapManager.cpp:177-192) and of
 * LoopClosureManager::loopFindNearKeyframes (backend/src/LoopClosureManager.cpp:40-60): every keyframe cloud i is
 * transformed by poses[i] (cast to float, pcp::transformPointCloud, common/pcp/pcp.hpp:38-62), the clouds are
 * concatenated in the given order and voxel-downsampled at `leaf` (pcp::voxelDownSample). The result becomes the
 * context's current target (index built as by pcr_set_target) and stays on the device; `out` (nullable, HOST) receives
 * up to `cap` 32-byte PointXYZI records, *m their number. poses: 16 doubles per cloud, column-major.
 * Keyframe clouds are immutable in the reference (KeyFrame::pc is a const shared_ptr), so a keyframe should cross PCIe once:
 * `ids` (nullable) names every cloud with a caller-chosen keyframe id (the keyframe's index in the reference's keyframe
 * deque); device copies are cached BY ID (never by host address) and re-uploaded when the count under an id changes.
 * The cache is bounded: entries not part of the submap being built are evicted least-recently-used first once the
 * cache exceeds its byte budget (pcr_submap_cache_budget, default 1 GiB). ids == NULL: nothing is cached. */
int pcr_submap_build(pcr_ctx* c, const void* const* clouds, const size_t* counts, const int64_t* ids, size_t n_clouds, size_t stride,
                     const double* poses, float leaf, void* out, size_t cap, size_t* m);
int pcr_submap_cache_budget(pcr_ctx* c, size_t bytes);
/* bytes / entries currently cached (either pointer may be NULL) */
int pcr_submap_cache_info(const pcr_ctx* c, size_t* bytes, size_t* entries);
int pcr_submap_cache_clear(pcr_ctx* c);

/* Multi-GPU: serialise the built target index into one contiguous DEVICE blob so that it can be broadcast with NCCL
 * (torch.distributed) and imported on the other ranks without rebuilding. */
int pcr_target_blob_size(pcr_ctx* c, size_t* bytes);
int pcr_target_export(pcr_ctx* c, void* dev_blob, size_t cap);
int pcr_target_import(pcr_ctx* c, const void* dev_blob, size_t bytes);

/* Logging (the reference logs through its spdlog singleton, common/utils/Logger.hpp:16-76: `lg->error(...)` at LoamRegister.cpp:174,
 * `lg->warn(...)` at LidarOdometry.cpp:199). Every error text this library produces (what pcr_last_error returns) and its warnings
 * are also handed to the process-wide callback: level 0 = debug, 1 = info, 2 = warning, 3 = error. Default (cb == NULL): errors and
 * warnings go to stderr. The callback may be invoked from any thread that uses a context. */
typedef void (*pcr_log_fn)(int32_t level, const char* message, void* user);
void pcr_set_logger(pcr_log_fn cb, void* user);

/* Page-lock a host buffer the caller keeps handing to the library (cudaHostRegister, portable): clouds in registered
 * memory cross PCIe by plain DMA at full rate instead of being staged. Worth it for memory that is uploaded many times
 * or is large (the static map of loc.cpp, a batch of scans); registering costs about a millisecond per 10 MB. The
 * adaptor's target-cache mode can do it for `dst` (CudaRegister::enableTargetCache(true, true)). */
int pcr_host_register(const void* p, size_t bytes);
int pcr_host_unregister(const void* p);

/* Device allocations released by contexts are parked in a process-wide, mutex-protected cache (cudaMalloc / cudaFree cost
 * milliseconds inside host-driven loops). A long-running caller can hand the parked buffers back to the driver at a quiet
 * moment: *freed_bytes (nullable) = bytes returned. Buffers owned by live contexts are not touched. */
int pcr_trim_device_cache(size_t* freed_bytes);

/* Several GPUs in ONE process (test/loc.cpp with a static map, SURVEY.md 8e / 8b "pcr_batch_align(ctxs/devices...)"): one context per
 * entry of devices[] (an ordinal may repeat). pcr_multi_set_target builds the index on devices[0], serialises it
 * (pcr_target_export) and copies the blob to every other device over NVLink (cudaMemcpyPeerAsync), where it is imported
 * without a rebuild. pcr_multi_batch_align shards the scans in contiguous blocks (scan i of n -> device i * n_devices / n),
 * runs the shards concurrently (one host thread per device, no collective while a registration runs) and returns all
 * poses / flags in the caller's arrays. Same argument meaning as pcr_set_target / pcr_batch_align. */
typedef struct pcr_multi pcr_multi;
int pcr_multi_create(const pcr_params* p, const int32_t* devices, size_t n_devices, pcr_multi** out);
void pcr_multi_destroy(pcr_multi* m);
const char* pcr_multi_last_error(const pcr_multi* m);
int pcr_multi_set_target(pcr_multi* m, const void* pts, size_t n, size_t stride);
int pcr_multi_batch_align(pcr_multi* m, const void* src, const size_t* offsets, size_t n_scans, size_t stride, double* T, int32_t* converged);
/* bytes of the index blob the last pcr_multi_set_target copied to each peer, and the milliseconds the copies took */
int pcr_multi_get_broadcast(const pcr_multi* m, size_t* blob_bytes, double* copy_ms);

/* On-disk index cache: the built target index (the same blob pcr_target_export produces) written to /
 * read from a file, so that the localisation mode (test/loc.cpp -> MapManager(pcd_file), frontend/src/MapManager.cpp:52-84)
 * does not re-downsample and re-index its static map at every start. */
int pcr_target_save(pcr_ctx* c, const char* path);
int pcr_target_load(pcr_ctx* c, const char* path);
/* Minimal PCD reader (pcp::loadPCDFile -> pcl::io::loadPCDFile<PointXYZI>, common/pcp/pcp.hpp:71-75): DATA ascii | binary,
 * float32 fields x y z and optionally intensity (other fields skipped). Writes up to `cap` 32-byte PointXYZI records
 * (x y z 1 | intensity 0 0 0) to the HOST buffer `out` (NULL: only count); *n = points in the file. Needs no GPU. */
int pcr_read_pcd(const char* path, void* out, size_t cap, size_t* n);
/* MapManager(pcd_file) in one call: read the PCD, voxel-downsample at `leaf` (pcp::voxelDownSample, MapManager.cpp:77) and
 * register the result as the static target. *m = points of the downsampled map. */
int pcr_static_map_load(pcr_ctx* c, const char* pcd_path, float leaf, size_t* m);

/* ScanContext place-recognition descriptor (backend/src/ScanContext.cpp). `pts`: n_clouds clouds
 * concatenated, cloud i = records [offsets[i], offsets[i+1]) (HOST). Per cloud: desc 1200 doubles (20 rings x 60 sectors,
 * row-major, maximum z + lidar_height per bin, radius <= 80 m; makeScanContext :152-196), ring_key 20 doubles (row means,
 * :198-212), sector_key 60 doubles (column means, :215-230). Output pointers are HOST and may be NULL. */
int pcr_scancontext_make(pcr_ctx* c, const void* pts, const size_t* offsets, size_t n_clouds, size_t stride, float lidar_height,
                         double* desc, double* ring_key, double* sector_key);
/* distanceBtnScanContext (:116-150) for a batch of descriptor pairs: descs = n_desc x 1200 doubles (HOST), pairs = 2 x
 * n_pairs indices (i0 j0 i1 j1 ...); search_ratio as in the config (0.1). dist[p] = minimum columnwise cosine distance,
 * shift[p] = the column shift of descriptor j that attains it. sector_key_align = 0 reproduces the reference, whose
 * fastAlignUsingVkey only ever evaluates shift 0 (its sector key is a 60 x 1 matrix and the loop runs over cols(), :93,122);
 * 1 = align over all 60 shifts of the sector key as the IROS'18 implementation intends. */
int pcr_scancontext_distance(pcr_ctx* c, const double* descs, size_t n_desc, const int32_t* pairs, size_t n_pairs, float search_ratio,
                             int32_t sector_key_align, double* dist, int32_t* shift);

/* ------------------------------------------------------------------------------------------------------------
 * Parity / introspection entry points (used by tests/ to compare every intermediate with the oracle).
 * All output pointers are HOST pointers and may be NULL.
 * ---------------------------------------------------------------------------------------------------------- */
/* voxel keys of the last pcr_voxel_downsample call: keys[n] per input point, out_keys[m], out_counts[m],
 * grid[9] = min_b[3], div_b[3], mul[3] */
int pcr_debug_voxel(pcr_ctx* c, int32_t* keys, int32_t* out_keys, int32_t* out_counts, int32_t grid[9]);

typedef struct pcr_loam_iter_log {
  double T_before[16];
  double JtJ[36]; /* row-major */
  double JtE[6];
  double x[6];
  int64_t n;
  int32_t converged;
  int32_t pad;
} pcr_loam_iter_log;
/* one LOAM linearisation at pose T (no update): per-point knn_idx[ns*5] (original target indices, -1 = gate
 * failed before 5 were found), status[ns] (0 gate, 1 plane invalid, 2 weight, 3 accepted), JtJ[36], JtE[6], n */
int pcr_loam_linearize(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double T[16], int32_t* knn_idx,
                       int32_t* status, double JtJ[36], double JtE[6], int64_t* n_acc);
/* kernel shape the last pcr_align / pcr_batch_align / pcr_loam_linearize of a LOAM context ran with: shape[0] = lanes per query
 * (1, 2, 4, 8), shape[1] = queries per warp pass, shape[2] = 1 when the iteration ran as the search kernel + the fit kernel
 * (large batches), 0 for the fused kernel. Lets the parity tests assert WHICH kernel variant they compared with the oracle. */
int pcr_loam_last_shape(pcr_ctx* c, int32_t shape[3]);
/* logs of the last pcr_align (scan 0 of a batch): returns count in *n */
int pcr_loam_get_logs(pcr_ctx* c, pcr_loam_iter_log* logs, int32_t cap, int32_t* n);

/* NDT leaves, ascending key: keys[L], npts[L] (-1 rejected), mean[3L], cov[9L], icov[9L]; grid[9] = min_b, max_b, div_b */
int pcr_ndt_num_leaves(pcr_ctx* c, size_t* n, int32_t grid[9]);
int pcr_ndt_get_leaves(pcr_ctx* c, int32_t* keys, int32_t* npts, double* mean, double* cov, double* icov);
/* computeDerivatives at transform vector p (xyz + euler xyz); Tf (column-major float[16]) NULL = built from p */
int pcr_ndt_derivatives(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double p[6], const float* Tf,
                        int32_t compute_hessian, double* score, double g[6], double H[36]);
int pcr_ndt_hessian(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double p[6], double H[36]);

/* GICP covariances of a cloud (V1): covs[n*9] row-major, knn_idx[n*k] */
int pcr_gicp_covariances(pcr_ctx* c, const void* pts, size_t n, size_t stride, int32_t k, double* covs, int32_t* knn_idx);
/* VGICP voxel map of the current target sorted by (z,y,x): coords[3V], npts[V], mean[3V], cov[9V] */
int pcr_vgicp_num_voxels(pcr_ctx* c, size_t* n);
int pcr_vgicp_get_voxels(pcr_ctx* c, int32_t* coords, int32_t* npts, double* mean, double* cov);
/* linearize at T0 evaluated at Ti (Ti == T0 for FastVGICP::linearize; Ti != T0 for compute_error). H, b may be NULL. */
int pcr_vgicp_evaluate(pcr_ctx* c, const void* src, size_t ns, size_t stride, const double T0[16], const double Ti[16],
                       double* cost, double H[36], double b[6], int64_t* n_corr);

#ifdef __cplusplus
}
#endif
#endif /* PCR_CUDA_H */
