// fast_gicp LsqRegistration control flow — computeTransformation / step_lm / step_gn / is_converged
// (third_parties/pclomp/src/lsq_registration_impl.hpp:53-172) — as a per-scan state machine that advances by ONE evaluation
// result at a time (linearize at x0, or compute_error of a Levenberg-Marquardt trial). Host + device: it runs in the tail
// of vgicp_eval_kernel (one thread of the last block), so a registration needs no host round trip between evaluations;
// compiled for the host it is driven by the CPU oracle's linearize / compute_error in tests/test_vgicp_logic.py.
#pragma once
#include "dev_linalg.cuh"
#include "host_math.hpp"
#include <stdint.h>

namespace pcr {

struct VgicpEvalParams {
  double T0[16];  // linearisation point (correspondences, Mahalanobis matrices)
  double Ti[16];  // where the error is evaluated (Ti == T0 for linearize)
  int want_hb;
  int scan;
  int pad[2];
};

struct VgicpCfg {
  int optimizer;      // PCR_LSQ_LM (0) | PCR_LSQ_GN (1)
  int max_iters;
  int lm_max_iters;
  int pad;
  double rot_eps, trans_eps, lm_init_lambda;
};

enum { VG_LINEARIZE = 0, VG_TRIAL = 1, VG_FINISHED = 2 };

struct VgicpState {
  VgicpEvalParams next;  // the evaluation this scan is waiting for
  int phase, pend;       // pend: 1 = `next` has to be evaluated
  int it, li;            // outer iteration, LM inner iteration
  int converged, nr_iterations;
  int n_linearize, n_error;
  long long total_corr, last_corr;
  double last_cost;
  double x0[16], xi[16], delta[16];
  double lm_lambda, nu, y0;
  double H[36], b[6], d[6];
};

namespace vgicp_logic {

PCR_HM void make_delta(const double* d, double* D) {  // delta.linear = so3_exp(d[0:3]), delta.translation = d[3:6]; column-major
  double R[9];
  hm::so3_exp_matrix(d, R);
  for (int i = 0; i < 16; i++) D[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) D[c * 4 + r] = R[r * 3 + c];
    D[12 + r] = d[3 + r];
  }
}

PCR_HM bool lsq_converged(const double* D, double rot_eps, double trans_eps) {  // :160-172
  double m = 0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) m = hm::std_max(m, 1.0 / rot_eps * fabs(D[c * 4 + r] - (r == c ? 1.0 : 0.0)));
    m = hm::std_max(m, 1.0 / trans_eps * fabs(D[12 + r]));
  }
  return m < 1;
}

PCR_HM void request(VgicpState& st, const double* T0, const double* Ti, int want_hb, int phase) {
  for (int i = 0; i < 16; i++) { st.next.T0[i] = T0[i]; st.next.Ti[i] = Ti[i]; }
  st.next.want_hb = want_hb;
  st.phase = phase;
  st.pend = 1;
}

PCR_HM void finish(VgicpState& st, bool conv) {
  st.converged = conv ? 1 : 0;
  st.phase = VG_FINISHED;
  st.pend = 0;
}

PCR_HM void start(VgicpState& st, const double* Tguess, int scan, const VgicpCfg& cfg) {
  for (int i = 0; i < 16; i++) st.x0[i] = double(static_cast<float>(Tguess[i]));  // VgicpRegister.cpp:36 cast<float>, lsq :54 cast<double>
  st.lm_lambda = -1.0;
  st.nu = 2.0;
  st.it = 0; st.li = 0;
  st.converged = 0; st.nr_iterations = 0;
  st.n_linearize = 0; st.n_error = 0; st.total_corr = 0; st.last_corr = 0; st.last_cost = 0.0;
  st.next.scan = scan;
  st.next.pad[0] = st.next.pad[1] = 0;
  if (cfg.max_iters <= 0) { finish(st, false); return; }
  request(st, st.x0, st.x0, 1, VG_LINEARIZE);
}

// one Levenberg-Marquardt trial (:121-150): solve (H + lambda I) d = -b, xi = delta * x0, ask for the error at xi
PCR_HM void try_lm(VgicpState& st) {
  double A[36], nb[6];
  for (int a = 0; a < 36; a++) A[a] = st.H[a];
  for (int a = 0; a < 6; a++) { A[a * 6 + a] += st.lm_lambda; nb[a] = -st.b[a]; }
  ldlt6_solve(A, nb, st.d);
  make_delta(st.d, st.delta);
  mat4_mul(st.delta, st.x0, st.xi);
  request(st, st.x0, st.xi, 0, VG_TRIAL);
}

// after an accepted step: convergence test, next outer iteration or the end (:57-79)
PCR_HM void after_step(VgicpState& st, const VgicpCfg& cfg) {
  const bool conv = lsq_converged(st.delta, cfg.rot_eps, cfg.trans_eps);
  st.it++;
  if (conv || st.it >= cfg.max_iters) { finish(st, conv); return; }
  request(st, st.x0, st.x0, 1, VG_LINEARIZE);
}

// v: the sums of the pending evaluation: cost, H upper 21 (row-major order r <= c), b[6], correspondences
PCR_HM void on_result(VgicpState& st, const double* v, const VgicpCfg& cfg) {
  st.total_corr += (long long)(v[28] + 0.5);
  if (st.phase == VG_LINEARIZE) {
    st.nr_iterations = st.it;
    st.n_linearize++;
    st.y0 = v[0];
    st.last_cost = v[0];
    st.last_corr = (long long)(v[28] + 0.5);
    int k = 1;
    for (int a = 0; a < 6; a++)
      for (int c = a; c < 6; c++) { st.H[a * 6 + c] = v[k]; st.H[c * 6 + a] = v[k]; k++; }
    for (int a = 0; a < 6; a++) st.b[a] = v[22 + a];
    if (cfg.optimizer == 1) {  // Gauss-Newton (:103-116)
      double nb[6], x1[16];
      for (int a = 0; a < 6; a++) nb[a] = -st.b[a];
      ldlt6_solve(st.H, nb, st.d);
      make_delta(st.d, st.delta);
      mat4_mul(st.delta, st.x0, x1);
      for (int i = 0; i < 16; i++) st.x0[i] = x1[i];
      after_step(st, cfg);
      return;
    }
    if (st.lm_lambda < 0.0) {  // :122-124
      double mx = 0;
      for (int a = 0; a < 6; a++) mx = hm::std_max(mx, fabs(st.H[a * 6 + a]));
      st.lm_lambda = cfg.lm_init_lambda * mx;
    }
    st.nu = 2.0;
    st.li = 0;
    if (cfg.lm_max_iters <= 0) { finish(st, false); return; }  // "lm not converged!!" (:69-72)
    try_lm(st);
    return;
  }
  if (st.phase == VG_TRIAL) {
    st.n_error++;
    const double yi = v[0];
    double den = 0;
    for (int a = 0; a < 6; a++) den += st.d[a] * (st.lm_lambda * st.d[a] - st.b[a]);
    const double rho = (st.y0 - yi) / den;
    if (rho < 0) {
      if (lsq_converged(st.delta, cfg.rot_eps, cfg.trans_eps)) { after_step(st, cfg); return; }  // :139-141: accepted as converged, x0 kept
      st.lm_lambda = st.nu * st.lm_lambda;
      st.nu = 2 * st.nu;
      st.li++;
      if (st.li >= cfg.lm_max_iters) { finish(st, false); return; }  // every trial rejected: "lm not converged!!"
      try_lm(st);
      return;
    }
    for (int i = 0; i < 16; i++) st.x0[i] = st.xi[i];
    const double t = 2 * rho - 1;
    st.lm_lambda = st.lm_lambda * hm::std_max(1.0 / 3.0, 1 - t * t * t);  // std::pow(2 rho - 1, 3)
    after_step(st, cfg);
  }
}

}  // namespace vgicp_logic
}  // namespace pcr
