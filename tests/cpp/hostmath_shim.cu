// Host-side shim over the PRODUCT's small dense linear algebra (simpleslam_b200/csrc/dev_linalg.cuh is __host__ __device__,
// host_math.hpp is plain C++), so that tests/test_product_linalg.py can exercise it on the CPU — no GPU needed — against
// numpy / scipy and against the oracle's independent restatement.
#include "../../simpleslam_b200/csrc/dev_linalg.cuh"
#include "../../simpleslam_b200/csrc/host_math.hpp"
#include "../../simpleslam_b200/csrc/ndt_logic.cuh"
#include "../../simpleslam_b200/csrc/vgicp_logic.cuh"
#include "../../simpleslam_b200/csrc/hostpack.hpp"
#include <cstring>

extern "C" {
void shim_cpqr5x3(const double* A, const double* b, double* x) {
  double a[5][3], bb[5], xx[3];
  for (int r = 0; r < 5; r++) { bb[r] = b[r]; for (int c = 0; c < 3; c++) a[r][c] = A[r * 3 + c]; }
  pcr::cpqr5x3_solve(a, bb, xx);
  for (int c = 0; c < 3; c++) x[c] = xx[c];
}
void shim_ldlt6(const double* A, const double* b, double* x) { pcr::ldlt6_solve(A, b, x); }
void shim_se3_exp(const double* k, double* E) { pcr::se3_exp(k, E); }
void shim_t2se3(double* T) { pcr::t2se3(T); }
void shim_mat4_mul(const double* A, const double* B, double* C) { pcr::mat4_mul(A, B, C); }
void shim_eig_sym3(const double* A, double* w, double* V) {
  double a[3][3], ww[3], vv[3][3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = A[r * 3 + c];
  pcr::eig_sym3(a, ww, vv);
  for (int r = 0; r < 3; r++) { w[r] = ww[r]; for (int c = 0; c < 3; c++) V[r * 3 + c] = vv[r][c]; }
}
void shim_inv3(const double* A, double* O) {
  double a[3][3], o[3][3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = A[r * 3 + c];
  pcr::inv3(a, o);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) O[r * 3 + c] = o[r][c];
}
void shim_svd6(const double* A, const double* b, double* x) { pcr::hm::svd6_solve(A, b, x); }
void shim_ndt_pose(const double* p, float* M) { pcr::hm::ndt_pose_matrix_f32(p, M); }
void shim_euler(const float* R, float* e) { pcr::hm::euler_xyz_f32(R, e); }
void shim_so3_exp(const double* om, double* R) { pcr::hm::so3_exp_matrix(om, R); }
void shim_angle_tables(const double* p, float* jf, float* hf, double* jd, double* hd) {
  float j[8][3], h[15][3];
  double jdd[8][3], hdd[15][3];
  pcr::hm::ndt_angle_tables(p, j, h, jdd, hdd);
  std::memcpy(jf, j, sizeof(j)); std::memcpy(hf, h, sizeof(h)); std::memcpy(jd, jdd, sizeof(jdd)); std::memcpy(hd, hdd, sizeof(hdd));
}
void shim_solve6_newton(const double* A, const double* b, double* x) { pcr::hm::solve6_newton(A, b, x); }
// NDT Newton / More-Thuente state machine (ndt_logic.cuh), the code the evaluation kernels' tails run on the device
size_t shim_ndt_state_size() { return sizeof(pcr::NdtScanState); }
void shim_ndt_start(void* st, const double* Tguess) {
  pcr::ndt_logic::start(*static_cast<pcr::NdtScanState*>(st), Tguess, 0);
  pcr::ndt_logic::fill_request(*static_cast<pcr::NdtScanState*>(st));
}
void shim_ndt_on_result(void* st, const double* v29, double step_size, double trans_eps, int max_iters) {
  pcr::NdtCfg cfg{step_size, trans_eps, max_iters, 0};
  pcr::ndt_logic::on_result(*static_cast<pcr::NdtScanState*>(st), v29, cfg);
  pcr::ndt_logic::fill_request(*static_cast<pcr::NdtScanState*>(st));
}
// pending request: pend (0 none, 1 float derivatives, 2 double Hessian), compute_hessian, p[6] of the evaluation, Tf[16]
void shim_ndt_pending(const void* stv, int* pend, int* hess, double* p, float* Tf) {
  const pcr::NdtScanState& st = *static_cast<const pcr::NdtScanState*>(stv);
  *pend = st.pend;
  *hess = st.next.compute_hessian;
  for (int i = 0; i < 6; i++) p[i] = st.eval_p[i];
  for (int i = 0; i < 16; i++) Tf[i] = st.next.Tf[i];
}
void shim_ndt_result(const void* stv, float* final_T, int* converged, int* nr_iterations, int* n_evals, int* n_hess, double* score) {
  const pcr::NdtScanState& st = *static_cast<const pcr::NdtScanState*>(stv);
  for (int i = 0; i < 16; i++) final_T[i] = st.final_T[i];
  *converged = st.converged; *nr_iterations = st.nr_iterations; *n_evals = st.n_evals; *n_hess = st.n_hess; *score = st.score;
}
// fast_gicp LM / GN state machine (vgicp_logic.cuh), the code vgicp_eval_kernel's tail runs on the device
size_t shim_vgicp_state_size() { return sizeof(pcr::VgicpState); }
static pcr::VgicpCfg shim_vg_cfg(int optimizer, int max_iters, int lm_max_iters, double rot_eps, double trans_eps, double lm_init_lambda) {
  pcr::VgicpCfg c{};
  c.optimizer = optimizer; c.max_iters = max_iters; c.lm_max_iters = lm_max_iters;
  c.rot_eps = rot_eps; c.trans_eps = trans_eps; c.lm_init_lambda = lm_init_lambda;
  return c;
}
void shim_vgicp_start(void* st, const double* Tguess, int optimizer, int max_iters, int lm_max_iters, double rot_eps, double trans_eps, double lm_init_lambda) {
  pcr::vgicp_logic::start(*static_cast<pcr::VgicpState*>(st), Tguess, 0, shim_vg_cfg(optimizer, max_iters, lm_max_iters, rot_eps, trans_eps, lm_init_lambda));
}
void shim_vgicp_on_result(void* st, const double* v29, int optimizer, int max_iters, int lm_max_iters, double rot_eps, double trans_eps, double lm_init_lambda) {
  pcr::vgicp_logic::on_result(*static_cast<pcr::VgicpState*>(st), v29, shim_vg_cfg(optimizer, max_iters, lm_max_iters, rot_eps, trans_eps, lm_init_lambda));
}
void shim_vgicp_pending(const void* stv, int* pend, int* want_hb, double* T0, double* Ti) {
  const pcr::VgicpState& st = *static_cast<const pcr::VgicpState*>(stv);
  *pend = st.pend; *want_hb = st.next.want_hb;
  for (int i = 0; i < 16; i++) { T0[i] = st.next.T0[i]; Ti[i] = st.next.Ti[i]; }
}
void shim_vgicp_result(const void* stv, double* T, int* converged, int* nr_iterations, int* n_linearize, int* n_error) {
  const pcr::VgicpState& st = *static_cast<const pcr::VgicpState*>(stv);
  for (int i = 0; i < 16; i++) T[i] = st.x0[i];
  *converged = st.converged; *nr_iterations = st.nr_iterations; *n_linearize = st.n_linearize; *n_error = st.n_error;
}
// host-side record packing of pageable uploads (hostpack.hpp): `repeat` calls on one pool
void shim_host_pack(const unsigned char* src, size_t n, size_t stride, int threads, int repeat, float* out) {
  pcr::HostPacker pk(threads);
  for (int r = 0; r < repeat; r++) pk.pack(src, n, stride, out);
}
}
