// pclomp NDT control flow — Newton direction + More-Thuente line search (ndt_omp_impl.hpp:81-171, 773-932) — as a
// per-scan state machine that advances by ONE evaluation result at a time. It runs on the DEVICE, in the tail of the
// evaluation kernels (one thread of the last block of a request, ndt.cu), so that a registration needs no host round
// trip between evaluations; the same code compiled for the host is driven by the CPU oracle's derivative evaluations in
// tests/test_ndt_logic.py (iteration / evaluation counts and poses against the oracle's own align, no GPU needed).
#pragma once
#include "host_math.hpp"
#include <stdint.h>

namespace pcr {

struct NdtEvalParams {  // per scan, per evaluation
  float Tf[16];         // cloud transform (column-major float)
  float j_ang[8][3];
  float h_ang[15][3];
  double j_ang_d[8][3];
  double h_ang_d[15][3];
  int compute_hessian;
  int kind;             // 0 = computeDerivatives (float path), 1 = computeHessian (double path)
  int scan;             // which scan of the batch
  int pad;
};

struct NdtCfg {  // the tunables the control flow reads (pcr_params)
  double step_size, trans_eps;
  int max_iters;
  int trace;  // PCR_NDT_TAIL_TRACE=1: the tail of a request stamps %globaltimer into NdtCounters::t_tail (single-scan diagnostics)
};

enum { NDT_PEND_NONE = 0, NDT_PEND_FLOAT = 1, NDT_PEND_DOUBLE = 2 };
enum { NDT_INIT_EVAL = 0, NDT_LS_FIRST = 1, NDT_LS_LOOP = 2, NDT_LS_HESSIAN = 3, NDT_FINISHED = 4 };

struct NdtScanState {
  NdtEvalParams next;    // the evaluation this scan is waiting for
  int phase;
  int pend;              // NDT_PEND_* of `next`
  int nr_iterations;
  int converged;
  int step_iterations;
  int interval_converged, open_interval;
  int n_evals, n_hess;
  int fill_matrix;       // the pending request's transform has to be rebuilt from eval_p (a line-search trial), see fill_request
  long long n_pairs;
  double eval_p[6];      // transform vector of the pending evaluation
  double p[6];
  double score;
  double g[6];
  double H[36];
  float final_T[16];
  // line search
  double step_dir[6], phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, x_t[6], phi_t, d_phi_t, psi_t, d_psi_t;
};

namespace ndt_logic {

constexpr double kMu = 1.e-4, kNu = 0.9;   // ndt_omp_impl.hpp:800-802
constexpr int kMaxStepIterations = 10;      // :786

// The control flow only DECIDES the next evaluation (its transform vector, kind, flags); the trigonometry behind its float
// pose matrix and its angular derivative tables is done by fill_request — serially here (host, tests), by parallel lanes in
// the kernel tail (ndt.cu: ndt_fill_request_warp, same functions on the same values).
PCR_HM void request(NdtScanState& st, const double* p, int kind, int hess, int rebuild_matrix) {
  for (int i = 0; i < 6; i++) st.eval_p[i] = p[i];
  st.fill_matrix = rebuild_matrix;
  st.next.compute_hessian = hess;
  st.next.kind = kind;
  st.pend = kind == 0 ? NDT_PEND_FLOAT : NDT_PEND_DOUBLE;
}

PCR_HM void fill_request(NdtScanState& st) {
  if (st.pend == NDT_PEND_NONE) return;
  if (st.fill_matrix) hm::ndt_pose_matrix_f32(st.eval_p, st.final_T);  // :827-830 / :866-869 final_transformation_ of the trial
  for (int i = 0; i < 16; i++) st.next.Tf[i] = st.final_T[i];
  hm::ndt_angle_tables(st.eval_p, st.next.j_ang, st.next.h_ang, st.next.j_ang_d, st.next.h_ang_d);
}

PCR_HM void finish(NdtScanState& st) {
  st.phase = NDT_FINISHED;
  st.pend = NDT_PEND_NONE;
}

PCR_HM void start(NdtScanState& st, const double* Tguess, int scan) {
  float guess[16];
  bool ident = true;
  for (int i = 0; i < 16; i++) {
    guess[i] = static_cast<float>(Tguess[i]);  // NdtRegister.cpp:27 res.matrix().cast<float>()
    if (guess[i] != ((i % 5 == 0) ? 1.f : 0.f)) ident = false;
  }
  for (int i = 0; i < 16; i++) st.final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  if (!ident)
    for (int i = 0; i < 16; i++) st.final_T[i] = guess[i];  // ndt_omp_impl.hpp:95-101
  float Rm[9], eul[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Rm[r * 3 + c] = st.final_T[c * 4 + r];
  hm::euler_xyz_f32(Rm, eul);  // :103-111
  st.p[0] = st.final_T[12]; st.p[1] = st.final_T[13]; st.p[2] = st.final_T[14];
  st.p[3] = eul[0]; st.p[4] = eul[1]; st.p[5] = eul[2];
  st.nr_iterations = 0;
  st.converged = 0;
  st.step_iterations = 0;
  st.interval_converged = 0;
  st.open_interval = 1;
  st.n_evals = 0; st.n_hess = 0; st.n_pairs = 0;
  st.score = 0.0;
  st.phase = NDT_INIT_EVAL;
  st.next.scan = scan;
  st.next.pad = 0;
  request(st, st.p, 0, 1, 0);  // the first evaluation runs at the guess itself (:95-101), not at a matrix rebuilt from p
}

PCR_HM void take(NdtScanState& st, const double* v, bool with_hessian) {
  st.score = v[0];
  for (int i = 0; i < 6; i++) st.g[i] = v[1 + i];
  int k = 7;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) { st.H[a * 6 + b] = with_hessian ? v[k] : 0.0; st.H[b * 6 + a] = st.H[a * 6 + b]; k++; }
}

PCR_HM void set_trial(NdtScanState& st, const NdtCfg& cfg) {
  st.a_t = hm::std_min(st.a_t, cfg.step_size);      // :822-823 (NaN trial values pass through, as in the reference)
  st.a_t = hm::std_max(st.a_t, cfg.trans_eps / 2);
  for (int i = 0; i < 6; i++) st.x_t[i] = st.p[i] + st.step_dir[i] * st.a_t;
}

PCR_HM void begin_outer(NdtScanState& st, const NdtCfg& cfg);

PCR_HM void end_outer(NdtScanState& st, const NdtCfg& cfg, double a) {
  for (int i = 0; i < 6; i++) st.p[i] += st.step_dir[i] * a;
  if (st.nr_iterations > cfg.max_iters || (st.nr_iterations && (fabs(a) < cfg.trans_eps))) st.converged = 1;  // :158-162
  st.nr_iterations++;
  if (st.converged) { finish(st); return; }
  begin_outer(st, cfg);
}

PCR_HM void begin_outer(NdtScanState& st, const NdtCfg& cfg) {
  double b[6], delta_p[6];
  for (int i = 0; i < 6; i++) b[i] = -st.g[i];
  hm::solve6_newton(st.H, b, delta_p);  // :127-129
  double nrm = 0;
  for (int i = 0; i < 6; i++) nrm += delta_p[i] * delta_p[i];
  nrm = sqrt(nrm);
  if (nrm == 0 || nrm != nrm) {  // :134-139
    st.converged = (nrm == nrm) ? 1 : 0;
    finish(st);
    return;
  }
  for (int i = 0; i < 6; i++) st.step_dir[i] = delta_p[i] / nrm;
  // computeStepLengthMT :773-
  st.phi_0 = -st.score;
  double d = 0;
  for (int i = 0; i < 6; i++) d += st.g[i] * st.step_dir[i];
  st.d_phi_0 = -d;
  if (st.d_phi_0 >= 0) {
    if (st.d_phi_0 == 0) {
      // a zero directional derivative returns step length 0 (:791-793); end_outer then either converges or starts the
      // next outer iteration from the same derivatives — bounded by max_iters, written as a loop instead of recursion
      for (;;) {
        if (st.nr_iterations > cfg.max_iters || st.nr_iterations) st.converged = 1;  // |0| < trans_eps
        st.nr_iterations++;
        if (st.converged) { finish(st); return; }
      }
    }
    st.d_phi_0 *= -1;
    for (int i = 0; i < 6; i++) st.step_dir[i] *= -1;
  }
  st.step_iterations = 0;
  st.a_l = 0; st.a_u = 0;
  st.f_l = hm::mt_psi(st.a_l, st.phi_0, st.phi_0, st.d_phi_0, kMu);
  st.g_l = hm::mt_dpsi(st.d_phi_0, st.d_phi_0, kMu);
  st.f_u = hm::mt_psi(st.a_u, st.phi_0, st.phi_0, st.d_phi_0, kMu);
  st.g_u = hm::mt_dpsi(st.d_phi_0, st.d_phi_0, kMu);
  st.interval_converged = (cfg.step_size - cfg.trans_eps / 2) < 0 ? 1 : 0;
  st.open_interval = 1;
  st.a_t = nrm;
  set_trial(st, cfg);
  st.phase = NDT_LS_FIRST;
  request(st, st.x_t, 0, 1, 1);
}

PCR_HM void after_eval(NdtScanState& st) {
  st.phi_t = -st.score;
  double d = 0;
  for (int i = 0; i < 6; i++) d += st.g[i] * st.step_dir[i];
  st.d_phi_t = -d;
  st.psi_t = hm::mt_psi(st.a_t, st.phi_t, st.phi_0, st.d_phi_0, kMu);
  st.d_psi_t = hm::mt_dpsi(st.d_phi_t, st.d_phi_0, kMu);
}

PCR_HM void ls_continue(NdtScanState& st, const NdtCfg& cfg) {
  if (!st.interval_converged && st.step_iterations < kMaxStepIterations && !(st.psi_t <= 0 && st.d_phi_t <= -kNu * st.d_phi_0)) {
    if (st.open_interval) st.a_t = hm::mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.psi_t, st.d_psi_t);
    else st.a_t = hm::mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.phi_t, st.d_phi_t);
    set_trial(st, cfg);
    st.phase = NDT_LS_LOOP;
    request(st, st.x_t, 0, 0, 1);
    return;
  }
  if (st.step_iterations) {  // :928-929 computeHessian
    st.phase = NDT_LS_HESSIAN;
    request(st, st.x_t, 1, 1, 0);  // same x_t as the last trial: final_T already belongs to it
    return;
  }
  end_outer(st, cfg, st.a_t);
}

// v: the 29 sums of the evaluation that was pending (score, g[6], H upper 21, pairs)
PCR_HM void on_result(NdtScanState& st, const double* v, const NdtCfg& cfg) {
  st.n_pairs += (long long)(v[28] + 0.5);
  switch (st.phase) {
    case NDT_INIT_EVAL:
      st.n_evals++;
      take(st, v, true);
      begin_outer(st, cfg);
      break;
    case NDT_LS_FIRST:
      st.n_evals++;
      take(st, v, true);
      after_eval(st);
      ls_continue(st, cfg);
      break;
    case NDT_LS_LOOP: {
      st.n_evals++;
      take(st, v, false);
      after_eval(st);
      if (st.open_interval && (st.psi_t <= 0 && st.d_psi_t >= 0)) {
        st.open_interval = 0;
        st.f_l = st.f_l + st.phi_0 - kMu * st.d_phi_0 * st.a_l;
        st.g_l = st.g_l + kMu * st.d_phi_0;
        st.f_u = st.f_u + st.phi_0 - kMu * st.d_phi_0 * st.a_u;
        st.g_u = st.g_u + kMu * st.d_phi_0;
      }
      bool ic;
      if (st.open_interval) ic = hm::mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.psi_t, st.d_psi_t);
      else ic = hm::mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.phi_t, st.d_phi_t);
      st.interval_converged = ic ? 1 : 0;
      st.step_iterations++;
      ls_continue(st, cfg);
      break;
    }
    case NDT_LS_HESSIAN: {
      st.n_hess++;
      int k = 7;
      for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) { st.H[a * 6 + b] = v[k]; st.H[b * 6 + a] = v[k]; k++; }
      end_outer(st, cfg, st.a_t);
      break;
    }
    default: break;
  }
}

}  // namespace ndt_logic
}  // namespace pcr
