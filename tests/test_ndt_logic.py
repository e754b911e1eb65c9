"""The NDT Newton / More-Thuente state machine the evaluation kernels run in their tails (simpleslam_b200/csrc/ndt_logic.cuh),
compiled for the HOST and driven on the CPU by the oracle's derivative evaluations: it must walk through exactly the
sequence of evaluations pclomp's computeTransformation / computeStepLengthMT make (ndt_omp_impl.hpp:81-171, 773-932) —
same number of outer iterations, derivative and Hessian evaluations, same final pose as the oracle's own align.
No GPU: this pins the control flow that is otherwise only reachable inside a kernel."""
import ctypes
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import synth
from test_product_linalg import shim, _p  # noqa: F401  (fixture: builds tests/cpp/libhostmath_shim.so)


def _run_state_machine(shim, ondt, src, Tg, search="DIRECT7", step_size=0.1, trans_eps=0.1, max_iters=35):
    shim.shim_ndt_state_size.restype = ctypes.c_size_t
    st = np.zeros(shim.shim_ndt_state_size() + 64, np.uint8)
    Tc = np.ascontiguousarray(np.asarray(Tg, dtype=np.float64).T).reshape(16).copy()
    shim.shim_ndt_start(_p(st), _p(Tc))
    pend, hess = ctypes.c_int(0), ctypes.c_int(0)
    p, Tf = np.empty(6), np.empty(16, np.float32)
    n_rounds = 0
    while True:
        shim.shim_ndt_pending(_p(st), ctypes.byref(pend), ctypes.byref(hess), _p(p), _p(Tf))
        if pend.value == 0:
            break
        v = np.zeros(29)
        if pend.value == 1:
            d = ondt.derivatives(src, p, search=search, compute_hessian=bool(hess.value), Tf=Tf.reshape(4, 4).T)
            v[0], v[1:7] = d["score"], d["g"]
            v[7:28] = d["H"][np.triu_indices(6)]
            v[28] = float((d["nb"] > 0).sum())
        else:
            v[7:28] = ondt.hessian(src, p, search=search)[np.triu_indices(6)]
        shim.shim_ndt_on_result(_p(st), _p(v), ctypes.c_double(step_size), ctypes.c_double(trans_eps), int(max_iters))
        n_rounds += 1
        assert n_rounds < 600
    fT = np.empty(16, np.float32)
    conv, nit, nev, nh, score = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
    shim.shim_ndt_result(_p(st), _p(fT), ctypes.byref(conv), ctypes.byref(nit), ctypes.byref(nev), ctypes.byref(nh), ctypes.byref(score))
    return dict(T=fT.reshape(4, 4).T.astype(np.float64), converged=bool(conv.value), nr_iterations=nit.value, n_evals=nev.value, n_hess=nh.value)


def test_state_machine_follows_the_oracle(shim):
    case = data.ndt_case()
    ondt = orc.Ndt(case["dst"], 1.0)
    rng = np.random.RandomState(5)
    guesses = [case["T_guess"], case["T_true"], np.eye(4)]
    for _ in range(4):
        pert = np.concatenate([rng.uniform(-0.5, 0.5, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-3, 3, 3)) * [0.3, 0.3, 1]])
        guesses.append(case["T_true"] @ synth.se3_exp(pert))
    for Tg in guesses:
        o = ondt.align(case["src"], Tg)
        r = _run_state_machine(shim, ondt, case["src"], Tg)
        assert r["converged"] == o["converged"] and r["nr_iterations"] == o["nr_iterations"]
        assert r["n_evals"] == o["n_derivative_evals"] and r["n_hess"] == o["n_hessian_evals"]
        dt, dr = data.pose_err(r["T"], o["T"])
        assert dt < 1e-5 and dr < 1e-5, (dt, dr)


def test_state_machine_options_and_degenerate_inputs(shim):
    case = data.ndt_case()
    ondt = orc.Ndt(case["dst"], 1.0)
    for kw in (dict(max_iterations=3), dict(trans_eps=0.01, step_size=0.3), dict(search="DIRECT1")):
        o = ondt.align(case["src"], case["T_guess"], **{k: v for k, v in kw.items()})
        r = _run_state_machine(shim, ondt, case["src"], case["T_guess"], search=kw.get("search", "DIRECT7"), step_size=kw.get("step_size", 0.1),
                               trans_eps=kw.get("trans_eps", 0.1), max_iters=kw.get("max_iterations", 35))
        assert r["converged"] == o["converged"] and r["nr_iterations"] == o["nr_iterations"] and r["n_evals"] == o["n_derivative_evals"], kw
        dt, dr = data.pose_err(r["T"], o["T"])
        assert dt < 1e-5 and dr < 1e-5
    # a scan far away from the map: all sums zero -> zero Newton step -> finished at once, like the oracle
    far = case["src"].copy()
    far[:, :3] += 5000.0
    o = ondt.align(far, case["T_guess"])
    r = _run_state_machine(shim, ondt, far, case["T_guess"])
    assert r["converged"] == o["converged"] and r["nr_iterations"] == o["nr_iterations"] and r["n_evals"] == o["n_derivative_evals"]


def test_newton_solve_fast_path_equals_svd(shim):
    rng = np.random.RandomState(2)
    for _ in range(200):
        J = rng.randn(30, 6) * rng.uniform(0.01, 100, size=6)
        H = np.ascontiguousarray(-(J.T @ J) + 1e-6 * rng.randn(6, 6))
        g = rng.randn(6)
        a, b = np.empty(6), np.empty(6)
        shim.shim_solve6_newton(_p(H), _p(g), _p(a))
        shim.shim_svd6(_p(H), _p(g), _p(b))
        assert np.allclose(a, b, rtol=1e-6, atol=1e-12 * np.abs(b).max())
    # rank deficient: falls back to the truncated SVD (minimum-norm solution), zero matrix -> zero step
    Q = np.linalg.qr(rng.randn(6, 6))[0]
    S = np.ascontiguousarray(Q @ np.diag([5, 3, 2, 1, 0, 0]) @ Q.T)
    g = rng.randn(6)
    a, b = np.empty(6), np.empty(6)
    shim.shim_solve6_newton(_p(S), _p(g), _p(a))
    shim.shim_svd6(_p(S), _p(g), _p(b))
    assert np.allclose(a, b, rtol=1e-9, atol=1e-12) and np.allclose(a, np.linalg.pinv(S) @ g, rtol=1e-8, atol=1e-10)
    Z = np.zeros((6, 6))
    shim.shim_solve6_newton(_p(Z), _p(g), _p(a))
    assert np.all(a == 0)
