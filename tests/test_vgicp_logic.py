"""The fast_gicp Levenberg-Marquardt / Gauss-Newton state machine that vgicp_eval_kernel runs in its tail
(simpleslam_b200/csrc/vgicp_logic.cuh), compiled for the HOST and driven on the CPU by the oracle's linearize /
compute_error: it has to ask for exactly the evaluations LsqRegistration::computeTransformation / step_lm / step_gn make
(third_parties/pclomp/src/lsq_registration_impl.hpp:53-172) — same number of outer iterations, linearisations and LM
trials, same final pose as the oracle's own align. No GPU: pins the control flow otherwise only reachable in a kernel."""
import ctypes
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import synth
from test_product_linalg import shim, _p  # noqa: F401  (fixture: builds tests/cpp/libhostmath_shim.so)


def _run_state_machine(shim, ovg, src, covs, Tg, optimizer="LM", max_iters=64, lm_max_iters=10, rot_eps=2e-3, trans_eps=5e-4, lm_init_lambda=1e-9):
    shim.shim_vgicp_state_size.restype = ctypes.c_size_t
    st = np.zeros(shim.shim_vgicp_state_size() + 64, np.uint8)
    cfg = (0 if optimizer == "LM" else 1, int(max_iters), int(lm_max_iters), ctypes.c_double(rot_eps), ctypes.c_double(trans_eps), ctypes.c_double(lm_init_lambda))
    Tc = np.ascontiguousarray(np.asarray(Tg, dtype=np.float64).T).reshape(16).copy()
    shim.shim_vgicp_start(_p(st), _p(Tc), *cfg)
    pend, want = ctypes.c_int(0), ctypes.c_int(0)
    T0, Ti = np.empty(16), np.empty(16)
    rounds = 0
    while True:
        shim.shim_vgicp_pending(_p(st), ctypes.byref(pend), ctypes.byref(want), _p(T0), _p(Ti))
        if pend.value == 0:
            break
        v = np.zeros(29)
        if want.value:
            assert np.array_equal(T0, Ti)  # a linearisation is evaluated where it is linearised
            lin = ovg.linearize(src, covs, T0.reshape(4, 4).T)
            v[0], v[1:22], v[22:28], v[28] = lin["cost"], lin["H"][np.triu_indices(6)], lin["b"], lin["n"]
        else:
            v[0] = ovg.error(src, covs, T0.reshape(4, 4).T, Ti.reshape(4, 4).T)
        shim.shim_vgicp_on_result(_p(st), _p(v), *cfg)
        rounds += 1
        assert rounds < 2000
    T = np.empty(16)
    conv, nit, nlin, nerr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    shim.shim_vgicp_result(_p(st), _p(T), ctypes.byref(conv), ctypes.byref(nit), ctypes.byref(nlin), ctypes.byref(nerr))
    return dict(T=T.reshape(4, 4).T.copy(), converged=bool(conv.value), nr_iterations=nit.value, n_linearize=nlin.value, n_error_evals=nerr.value)


@pytest.fixture(scope="module")
def case():
    c = data.vgicp_case()
    src = c["src"][::3]  # the control flow does not depend on the cloud size; keep the CPU suite short
    return dict(c, src=src, covs=orc.gicp_covariances(src, 20), ovg=orc.Vgicp(c["dst"], 1.0, 20))


@pytest.mark.parametrize("optimizer", ["LM", "GN"])
def test_state_machine_follows_the_oracle(shim, case, optimizer):
    rng = np.random.RandomState(3)
    guesses = [case["T_guess"], case["T_true"], np.eye(4)]
    for _ in range(2):
        pert = np.concatenate([rng.uniform(-0.4, 0.4, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-3, 3, 3)) * [0.3, 0.3, 1]])
        guesses.append(case["T_true"] @ synth.se3_exp(pert))
    for Tg in guesses:
        o = case["ovg"].align(case["src"], Tg, src_covs=case["covs"], optimizer=optimizer)
        r = _run_state_machine(shim, case["ovg"], case["src"], case["covs"], Tg, optimizer=optimizer)
        assert (r["converged"], r["nr_iterations"], r["n_linearize"], r["n_error_evals"]) == \
               (o["converged"], o["nr_iterations"], o["n_linearize"], o["n_error_evals"])
        dt, dr = data.pose_err(r["T"], o["T"])
        assert dt < 1e-6 and dr < 1e-6, (dt, dr)  # the oracle hands the pose back through float, like VgicpRegister.cpp:40


def test_state_machine_limits_and_degenerate_inputs(shim, case):
    ovg, src, covs = case["ovg"], case["src"], case["covs"]
    for kw in (dict(max_iterations=2), dict(rot_eps=1e-5, trans_eps=1e-6), dict(max_iterations=1, optimizer="GN")):
        o = ovg.align(src, case["T_guess"], src_covs=covs, **kw)
        r = _run_state_machine(shim, ovg, src, covs, case["T_guess"], optimizer=kw.get("optimizer", "LM"), max_iters=kw.get("max_iterations", 64),
                               rot_eps=kw.get("rot_eps", 2e-3), trans_eps=kw.get("trans_eps", 5e-4))
        assert (r["converged"], r["nr_iterations"], r["n_linearize"], r["n_error_evals"]) == \
               (o["converged"], o["nr_iterations"], o["n_linearize"], o["n_error_evals"]), kw
        dt, dr = data.pose_err(r["T"], o["T"])
        assert dt < 1e-6 and dr < 1e-6
    # a scan with no voxel under it: H = b = 0 -> the LDLT of lambda*I = 0 gives a non-finite step; whatever the oracle does, follow it
    far = src.copy()
    far[:, :3] += 5000.0
    o = ovg.align(far, case["T_guess"], src_covs=covs)
    r = _run_state_machine(shim, ovg, far, covs, case["T_guess"])
    assert (r["converged"], r["nr_iterations"], r["n_linearize"], r["n_error_evals"]) == (o["converged"], o["nr_iterations"], o["n_linearize"], o["n_error_evals"])
    # no iterations allowed: nothing is evaluated, the (float-rounded) guess comes back
    r = _run_state_machine(shim, ovg, src, covs, case["T_guess"], max_iters=0)
    assert not r["converged"] and r["n_linearize"] == 0
    assert np.allclose(r["T"], case["T_guess"].astype(np.float32).astype(np.float64), atol=0)
