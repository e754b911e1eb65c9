"""Adversarial known-answer cases for the Eigen-dependent pieces the oracle restates WITHOUT being able to link Eigen here
(parity stays 'unpinned'): each case states the result Eigen's documented algorithm gives and checks the
oracle AND the product's host-compiled code against it.
  * ColPivHouseholderQR (Eigen/src/QR/ColPivHouseholderQR.h): pivot = column of largest remaining norm; a pivot is
    'nonzero' iff |R_kk| > threshold * maxpivot with threshold = eps * diagonalSize (= 3 here); solve() zeroes the
    coefficients of the non-pivot columns (basic solution).
  * SelfAdjointEigenSolver<Matrix3d>: eigenvalues ascending, eigenvectors orthonormal; repeated eigenvalues give an
    arbitrary orthonormal basis of the eigenspace -> only invariants are checked there.
  * Matrix3f::eulerAngles(0, 1, 2) (Eigen 3.3 Geometry/EulerAngles.h): res[0] = atan2(m(1,2), m(2,2)); c2 = |(m(0,0), m(0,1))|;
    if res[0] > 0 { res[0] -= pi; res[1] = atan2(-m(0,2), -c2) } else res[1] = atan2(-m(0,2), c2); res[2] from the
    rotated rows; result negated. At gimbal lock (c2 = 0) the formulas still apply and give a finite answer.
  * JacobiSVD::solve: singular values <= eps * diagSize * sigma_max are dropped (minimum-norm solution).
"""
import ctypes
import numpy as np
import pytest
from oracle import pyoracle as orc
from test_product_linalg import shim, _p  # noqa: F401


def _cpqr_both(shim, A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    xo, xp = np.empty(3), np.empty(3)
    rank = orc.lib().orc_test_cpqr5x3(_p(A.copy()), _p(b.copy()), _p(xo))
    shim.shim_cpqr5x3(_p(A.copy()), _p(b.copy()), _p(xp))
    return rank, xo, xp


def test_cpqr_exact_rank_two_keeps_two_pivots_and_zeroes_the_third(shim):
    # columns: c0 = e-direction of norm 10, c1 of norm 2, c2 = 0.5 * c0 exactly -> pivots: c0 (largest), then c1; c2's
    # remainder after deflation is exactly 0 -> not a nonzero pivot -> its coefficient is 0 in Eigen's solve()
    c0 = np.array([6.0, 0.0, 8.0, 0.0, 0.0])
    c1 = np.array([0.0, 2.0, 0.0, 0.0, 0.0])
    A = np.stack([c0, c1, 0.5 * c0], axis=1)
    b = np.array([3.0, 4.0, 4.0, 7.0, -1.0])
    rank, xo, xp = _cpqr_both(shim, A, b)
    assert rank == 2
    # least squares on the two pivot columns: x0 = c0.b / |c0|^2 = (18 + 32) / 100 = 0.5, x1 = 4 / 2 = 2, x2 = 0
    for x in (xo, xp):
        assert np.allclose(x, [0.5, 2.0, 0.0], rtol=0, atol=1e-14)


def test_cpqr_pivot_order_changes_the_basic_solution(shim):
    # same column space, but now the THIRD column is the long one: Eigen pivots on it first, so the basic solution puts
    # the weight on column 2 and zeroes column 0
    c0 = np.array([6.0, 0.0, 8.0, 0.0, 0.0])
    c1 = np.array([0.0, 2.0, 0.0, 0.0, 0.0])
    A = np.stack([0.5 * c0, c1, c0], axis=1)
    b = np.array([3.0, 4.0, 4.0, 7.0, -1.0])
    rank, xo, xp = _cpqr_both(shim, A, b)
    assert rank == 2
    for x in (xo, xp):
        assert np.allclose(x, [0.0, 2.0, 0.5], rtol=0, atol=1e-14)


def test_cpqr_threshold_band(shim):
    # third column = 0.5 * c0 + delta * e4: |R_22| = delta. Eigen keeps the pivot iff delta > 3 * eps * maxpivot (maxpivot = 10).
    c0 = np.array([6.0, 0.0, 8.0, 0.0, 0.0])
    c1 = np.array([0.0, 2.0, 0.0, 0.0, 0.0])
    e4 = np.array([0.0, 0.0, 0.0, 1.0, 0.0])
    b = np.array([3.0, 4.0, 4.0, 7.0, -1.0])
    eps = np.finfo(np.float64).eps
    thr = 3 * eps * 10.0
    for delta, want in ((100 * thr, 3), (4 * thr, 3), (0.25 * thr, 2), (0.0, 2)):
        A = np.stack([c0, c1, 0.5 * c0 + delta * e4], axis=1)
        rank, xo, xp = _cpqr_both(shim, A, b)
        assert rank == want, (delta / thr, rank)
        assert np.allclose(xo, xp, rtol=1e-9, atol=1e-12), "product and oracle take the same side of the threshold"
        if want == 2:
            assert xo[2] == 0.0 and xp[2] == 0.0


def test_cpqr_plane_far_from_origin_is_well_conditioned_enough(shim):
    # the LOAM use: five points of a plane 400 m from the origin, spread 0.4 m: n.p + 1 = 0 scaled -> x = -n / d
    rng = np.random.RandomState(0)
    n = np.array([0.6, 0.0, 0.8])
    d = 400.0
    for _ in range(100):
        uv = rng.uniform(-0.2, 0.2, (5, 2))
        t1, t2 = np.array([0.8, 0.0, -0.6]), np.array([0.0, 1.0, 0.0])
        P = d * n + uv[:, :1] * t1 + uv[:, 1:] * t2
        rank, xo, xp = _cpqr_both(shim, P, -np.ones(5))
        assert rank == 3
        for x in (xo, xp):
            assert np.allclose(x, -n / d, rtol=1e-7, atol=1e-12)
        assert np.allclose(xo, xp, rtol=1e-10, atol=1e-14)


def test_eig3_repeated_and_zero_eigenvalues(shim):
    cases = [np.diag([2.0, 2.0, 5.0]), np.diag([3.0, 3.0, 3.0]), np.diag([0.0, 0.0, 4.0]), np.zeros((3, 3))]
    R = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])  # exact rotation
    cases += [R @ c @ R.T for c in cases[:3]]
    for A in cases:
        A = np.ascontiguousarray(0.5 * (A + A.T))
        for fn in (orc.lib().orc_test_eig3, shim.shim_eig_sym3):
            w, V = np.empty(3), np.empty(9)
            fn(_p(A), _p(w), _p(V))
            V = V.reshape(3, 3)
            assert np.all(np.diff(w) >= -1e-15), "ascending like SelfAdjointEigenSolver"
            assert np.allclose(w, np.linalg.eigvalsh(A), rtol=0, atol=1e-13)
            assert np.allclose(V.T @ V, np.eye(3), atol=1e-12), "orthonormal basis even inside a repeated eigenspace"
            assert np.allclose(A @ V, V * w, atol=1e-12)


def _euler_expected(R):
    """Matrix3f::eulerAngles(0,1,2) of Eigen 3.3, written out in float32"""
    f = np.float32
    m = R.astype(np.float32)
    r0 = f(np.arctan2(m[1, 2], m[2, 2]))
    c2 = f(np.sqrt(f(m[0, 0] * m[0, 0]) + f(m[0, 1] * m[0, 1])))
    if r0 > 0:
        r0 = f(r0 - f(np.pi))
        r1 = f(np.arctan2(-m[0, 2], -c2))
    else:
        r1 = f(np.arctan2(-m[0, 2], c2))
    s1, c1 = f(np.sin(r0)), f(np.cos(r0))
    r2 = f(np.arctan2(f(s1 * m[2, 0]) - f(c1 * m[1, 0]), f(c1 * m[1, 1]) - f(s1 * m[2, 1])))
    return np.array([-r0, -r1, -r2], np.float32)


def test_euler_angles_gimbal_lock_and_branches(shim):
    from scipy.spatial.transform import Rotation
    cases = [Rotation.from_euler("XYZ", a).as_matrix() for a in
             ([0.3, np.pi / 2, 0.2], [0.3, -np.pi / 2, 0.2], [0.0, np.pi / 2, 0.0], [2.5, 0.4, -1.0], [-2.5, -0.4, 1.0], [np.pi, 0.0, 0.0],
              [0.0, 0.0, np.pi], [1e-7, 1e-7, 1e-7], [0.0, 0.0, 0.0])]
    for R in cases:
        Rf = np.ascontiguousarray(R, dtype=np.float32)
        exp = _euler_expected(Rf)
        eo, ep = np.empty(3, np.float32), np.empty(3, np.float32)
        orc.lib().orc_euler_xyz_f32(_p(Rf.reshape(9).copy()), _p(eo))
        shim.shim_euler(_p(Rf.reshape(9).copy()), _p(ep))
        assert np.all(np.isfinite(eo)) and np.all(np.isfinite(ep))
        assert np.allclose(eo, exp, atol=2e-6) and np.allclose(ep, exp, atol=2e-6), (R, eo, ep, exp)
        # and the angles reproduce the rotation wherever it is not degenerate (|cos pitch| > 1e-3)
        if abs(np.cos(exp[1])) > 1e-3:
            back = Rotation.from_euler("XYZ", ep.astype(np.float64)).as_matrix()
            assert np.allclose(back, R, atol=2e-6)


def test_svd_solve_singular_and_threshold(shim):
    rng = np.random.RandomState(7)
    Q = np.linalg.qr(rng.randn(6, 6))[0]
    g = rng.randn(6)
    eps = np.finfo(np.float64).eps
    for tail, kept in (([1e-3, 1e-6], 6), ([1e-3, 0.0], 5), ([0.0, 0.0], 4), ([1e-3, 0.1 * 6 * eps * 5.0], 5)):
        s = np.array([5.0, 3.0, 2.0, 1.0] + tail)
        A = np.ascontiguousarray(Q @ np.diag(s) @ Q.T)
        want = Q[:, :kept] @ ((Q[:, :kept].T @ g) / s[:kept])   # Eigen: singular values <= eps * 6 * sigma_max are dropped
        for fn in (orc.lib().orc_test_svd6, shim.shim_svd6, shim.shim_solve6_newton):
            x = np.empty(6)
            fn(_p(A), _p(g.copy()), _p(x))
            assert np.allclose(x, want, rtol=1e-6, atol=1e-9 * np.abs(want).max()), (tail, fn)
    Z = np.zeros((6, 6))
    for fn in (orc.lib().orc_test_svd6, shim.shim_svd6, shim.shim_solve6_newton):
        x = np.ones(6)
        fn(_p(Z), _p(g.copy()), _p(x))
        assert np.all(x == 0), "zero matrix: every singular value is dropped, the Newton step is zero (NDT then stops)"


def test_loam_gate_ambiguous_band_is_empty_on_the_test_data():
    """how close do the parity cases come to a decision threshold? (5th-neighbour gate d2 < 1, plane threshold 0.2, weight
    threshold 0.1): counts of queries within 1e-9 (relative) of a threshold — a status mismatch there would not be a bug"""
    import data
    case = data.loam_case()
    o = orc.loam_linearize(case["src"], case["dst"], case["T_guess"])
    idx = o["knn_idx"]
    gate = o["status"] >= 1
    q = (case["src"][:, :3].astype(np.float64) @ case["T_guess"][:3, :3].T + case["T_guess"][:3, 3]).astype(np.float32).astype(np.float64)
    d5 = np.sum((q[gate] - case["dst"][idx[gate][:, 4], :3].astype(np.float64)) ** 2, axis=1)
    assert (d5 < 1.0).all()
    assert int(np.sum(np.abs(d5 - 1.0) < 1e-9)) == 0
