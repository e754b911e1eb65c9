/* ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the SimpleSLAM PCR hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product path (simpleslam_b200/, include/pcr_cuda.h) never links it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors / fixtures for this path and its
 * registration libraries need PCL + Eigen + FLANN (absent here), so they cannot be compiled in this
 * container. The one piece that does compile — the vendored nanoflann kd-tree —
 * is built into oracle/_ref and used to cross-check the kNN restatement (tests/test_oracle_knn.py).
 *
 * Conventions: clouds are float32 arrays with a stride given in floats (8 for pcl::PointXYZI,
 * xyz at 0..2, intensity at 4); poses are column-major double[16] (Eigen::Isometry3d::matrix()).
 */
#ifndef PCR_ORACLE_H
#define PCR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- A1: pcl::VoxelGrid semantics (pcp.hpp:15-28,191-210) ---------------- */
/* keys_out[n] (nullable): voxel key per input point. out_pts: capacity n*8 floats (PointXYZI
 * records, pad=1). out_keys[m], out_counts[m] nullable. Returns 0, or 1 if the grid overflowed
 * int32 (PCL then returns the input unchanged: out = in, m = n). */
int orc_voxel_downsample(const float* pts, size_t n, size_t stride_f, float leaf,
                         int32_t* keys_out, float* out_pts, int32_t* out_keys, int32_t* out_counts,
                         size_t* m, int32_t grid_out[9] /* min_b[3], div_b[3], mul[3]; nullable */);

/* ---- A3/A4: exact kNN, (d2, idx) tie-break. metric_float=0: double metric (nanoflann
 * L2_Simple<double>), queries are doubles; metric_float=1: FLANN float metric, queries are read as
 * doubles and rounded to float first. brute=1 forces the O(n*m) scan. idx_out [nq*k] (-1 padded),
 * d2_out [nq*k] double. */
int orc_knn(const float* map, size_t nm, size_t stride_f, const double* queries, size_t nq, int k,
            int metric_float, int brute, float cell, int threads, int64_t* idx_out, double* d2_out);

/* C-bar of SURVEY.md 8(d): mean population of the 27 cells (cell = gate radius) around a query */
double orc_neighbourhood27(const float* map, size_t nm, size_t stride_f, const double* queries, size_t nq, float cell, int threads);

/* ---- LOAM (PCR/src/LoamRegister.cpp:99-223) ----------------------------------------------------- */
typedef struct orc_loam_iter_log {
  double T_before[16]; /* pose the iteration linearised at */
  double JtJ[36];      /* row-major 6x6 */
  double JtE[6];
  double x[6];         /* LDLT solution (translation first, rotation last) */
  int64_t n;           /* number of accepted residuals */
  int32_t converged;   /* this iteration hit the convergence test (update not applied) */
  int32_t pad;
} orc_loam_iter_log;

/* One linearisation at pose T. Per-point outputs (all nullable): knn_idx[ns*5] (-1 when fewer than 5
 * found), knn_d2[ns*5], status[ns] (0 = kNN gate failed, 1 = plane invalid, 2 = weight <= 0.1,
 * 3 = accepted), resid[ns] (s*dist), J[ns*6]. Accumulates JtJ/JtE/n in ascending point order. */
int orc_loam_linearize(const float* src, size_t ns, size_t sstride, const float* dst, size_t nm,
                       size_t dstride, const double T[16], int threads, int64_t* knn_idx,
                       double* knn_d2, int32_t* status, double* resid, double* J, double JtJ[36],
                       double JtE[6], int64_t* n_acc);

/* Full scan2Map. logs: capacity max_iters entries. Returns 0. */
int orc_loam_align(const float* src, size_t ns, size_t sstride, const float* dst, size_t nm,
                   size_t dstride, double T[16], int threads, int max_iters,
                   orc_loam_iter_log* logs, int32_t* n_iters, int32_t* converged);

/* manifolds::exp (common/geometry/manifolds.hpp:33-60) and trans::T2SE3 (trans.hpp:54-65). */
void orc_se3_exp(const double x[6], double T[16]);
void orc_t2se3(double T[16]);

/* ---- NDT (third_parties/pclomp ndt_omp_impl.hpp, voxel_grid_covariance_omp_impl.hpp) ------------ */
typedef struct orc_ndt orc_ndt;
orc_ndt* orc_ndt_create(const float* dst, size_t nm, size_t dstride, float resolution);
void orc_ndt_destroy(orc_ndt*);
/* grid_out: min_b[3], max_b[3], div_b[3]. */
size_t orc_ndt_num_leaves(const orc_ndt*, int32_t grid_out[9]);
/* per leaf, ascending key: key, nr_points (-1 = rejected), mean[3], cov[9], icov[9] (row-major) */
void orc_ndt_get_leaves(const orc_ndt*, int32_t* keys, int32_t* npts, double* mean, double* cov, double* icov);
/* computeDerivatives at transform vector p[6] (xyz + euler xyz); the cloud is transformed from src by
 * the float matrix built from p exactly as computeStepLengthMT does (:827-833). search: 0 KDTREE,
 * 1 DIRECT26, 2 DIRECT7, 3 DIRECT1. H row-major 6x6. nb_count[ns] nullable: neighbours per point. */
double orc_ndt_derivatives(const orc_ndt*, const float* src, size_t ns, size_t sstride,
                           const double p[6], int search, int compute_hessian, int threads,
                           double g[6], double H[36], int32_t* nb_count);
/* same but with an explicit float 4x4 (column-major) for the cloud transform (first call path :100) */
double orc_ndt_derivatives_T(const orc_ndt*, const float* src, size_t ns, size_t sstride,
                             const float Tf[16], const double p[6], int search, int compute_hessian,
                             int threads, double g[6], double H[36], int32_t* nb_count);
/* computeHessian (double path, :541-645) */
void orc_ndt_hessian(const orc_ndt*, const float* src, size_t ns, size_t sstride, const double p[6],
                     int search, double H[36]);
typedef struct orc_ndt_result {
  double T[16];            /* final pose (float matrix widened) */
  double trans_probability;
  int32_t converged;
  int32_t nr_iterations;
  int32_t n_derivative_evals;
  int32_t n_hessian_evals;
  double p_final[6];
} orc_ndt_result;
/* NdtRegister::scan2Map minus the target build (pass a built orc_ndt). T in = guess. */
int orc_ndt_align(const orc_ndt*, const float* src, size_t ns, size_t sstride, const double T[16],
                  int search, int threads, int max_iterations, double trans_eps, double step_size,
                  orc_ndt_result* out);
void orc_euler_xyz_f32(const float R[9] /* row-major */, float out[3]);

/* ---- VGICP (fast_gicp_impl.hpp, fast_vgicp_impl.hpp, fast_vgicp_voxel.hpp, lsq_registration) ----- */
/* V1: covs_out [n*9] row-major 3x3 (upper-left block of the reference's 4x4). knn_idx_out nullable [n*k]. */
int orc_gicp_covariances(const float* pts, size_t n, size_t stride_f, int k, int threads,
                         double* covs_out, int64_t* knn_idx_out);
typedef struct orc_vgicp orc_vgicp;
/* builds target covariances (unless given, nullable) + voxel map */
orc_vgicp* orc_vgicp_create(const float* dst, size_t nm, size_t dstride, double resolution, int k,
                            int threads, const double* target_covs /* nullable [nm*9] */);
void orc_vgicp_destroy(orc_vgicp*);
size_t orc_vgicp_num_voxels(const orc_vgicp*);
/* sorted by (z,y,x) coordinate: coords[3*v], npts[v], mean[3*v], cov[9*v] */
void orc_vgicp_get_voxels(const orc_vgicp*, int32_t* coords, int32_t* npts, double* mean, double* cov);
/* linearize at T (correspondences + mahalanobis from T); returns cost. H row-major 6x6 (rot first). */
double orc_vgicp_linearize(const orc_vgicp*, const float* src, size_t ns, size_t sstride,
                           const double* src_covs, const double T[16], int threads, double H[36],
                           double b[6], int64_t* n_corr);
/* compute_error at Ti using correspondences/mahalanobis of T0 */
double orc_vgicp_error(const orc_vgicp*, const float* src, size_t ns, size_t sstride,
                       const double* src_covs, const double T0[16], const double Ti[16], int threads);
typedef struct orc_vgicp_result {
  double T[16];
  int32_t converged;
  int32_t nr_iterations;
  int32_t n_linearize;
  int32_t n_error_evals;
} orc_vgicp_result;
/* optimizer: 0 LM (reference default), 1 GN. */
int orc_vgicp_align(const orc_vgicp*, const float* src, size_t ns, size_t sstride,
                    const double* src_covs /* nullable: computed */, const double T[16], int threads,
                    int optimizer, int max_iterations, double rot_eps, double trans_eps,
                    orc_vgicp_result* out);
/* V6: pcl::Registration::getFitnessScore(max_range) with a float transform */
double orc_fitness(const float* src, size_t ns, size_t sstride, const float* dst, size_t nm,
                   size_t dstride, const double T[16], double max_range, int threads);

#ifdef __cplusplus
}
#endif
#endif
