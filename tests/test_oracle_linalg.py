"""The oracle's dense linear algebra restatements (Eigen algorithms) against numpy — CPU only."""
import ctypes
import numpy as np
from oracle import pyoracle as orc


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def test_cpqr_full_rank_matches_lstsq():
    rng = np.random.RandomState(0)
    for _ in range(200):
        A = rng.randn(5, 3) * rng.uniform(0.1, 50)
        b = -np.ones(5)
        x = np.empty(3)
        r = orc.lib().orc_test_cpqr5x3(_p(np.ascontiguousarray(A)), _p(b), _p(x))
        assert r == 3
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(x, ref, rtol=1e-9, atol=1e-12)


def test_cpqr_rank_deficient_gives_basic_solution():
    # five collinear points (rank 2 after pivoting, or 1): Eigen returns the *basic* solution: at least one zero coefficient,
    # and it is still a least-squares minimiser
    rng = np.random.RandomState(1)
    for _ in range(50):
        d = rng.randn(3)
        o = rng.randn(3)
        A = np.stack([o * 0 + d * t for t in rng.randn(5)])  # rank 1
        b = -np.ones(5)
        x = np.empty(3)
        r = orc.lib().orc_test_cpqr5x3(_p(np.ascontiguousarray(A)), _p(b), _p(x))
        # Eigen's pivot threshold sits right at the rounding noise of the deflated columns, so 1 or 2 pivots may be kept
        assert r in (1, 2)
        assert (x == 0).sum() >= 3 - r
        if r == 1:
            ref = np.linalg.lstsq(A, b, rcond=None)[0]
            assert np.isclose(np.linalg.norm(A @ x - b), np.linalg.norm(A @ ref - b), rtol=1e-9)


def test_ldlt_and_svd_solve():
    rng = np.random.RandomState(2)
    for _ in range(100):
        J = rng.randn(40, 6) * rng.uniform(0.1, 10, size=6)
        A = J.T @ J
        b = rng.randn(6)
        x = np.empty(6)
        orc.lib().orc_test_ldlt6(_p(np.ascontiguousarray(A)), _p(b), _p(x))
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-12)
        # general (indefinite, slightly non-symmetric) matrix through the SVD solve
        G = -A + 1e-3 * rng.randn(6, 6)
        orc.lib().orc_test_svd6(_p(np.ascontiguousarray(G)), _p(b), _p(x))
        assert np.allclose(x, np.linalg.solve(G, b), rtol=1e-7, atol=1e-12)


def test_svd_solve_rank_truncation():
    # singular matrix: Eigen's JacobiSVD::solve drops singular values below diagSize*eps*sigma_max -> minimum-norm solution
    rng = np.random.RandomState(3)
    Q = np.linalg.qr(rng.randn(6, 6))[0]
    s = np.array([5.0, 3.0, 2.0, 1.0, 0.5, 0.0])
    A = Q @ np.diag(s) @ Q.T
    b = rng.randn(6)
    x = np.empty(6)
    orc.lib().orc_test_svd6(_p(np.ascontiguousarray(A)), _p(b), _p(x))
    assert np.allclose(x, np.linalg.pinv(A) @ b, rtol=1e-8, atol=1e-10)


def test_eig3_ascending_orthonormal():
    rng = np.random.RandomState(4)
    for _ in range(100):
        B = rng.randn(3, 3)
        A = B @ B.T * rng.uniform(1e-3, 1e3)
        w = np.empty(3)
        V = np.empty(9)
        orc.lib().orc_test_eig3(_p(np.ascontiguousarray(A)), _p(w), _p(V))
        V = V.reshape(3, 3)
        assert np.all(np.diff(w) >= 0)
        assert np.allclose(w, np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-14)
        assert np.allclose(V @ np.diag(w) @ V.T, A, rtol=1e-10, atol=1e-12 * np.abs(A).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-12)


def test_se3_exp_and_t2se3():
    from scipy.linalg import expm
    rng = np.random.RandomState(5)
    for _ in range(50):
        x = rng.randn(6) * 0.3
        T = orc.se3_exp(x)
        wx = np.array([[0, -x[5], x[4]], [x[5], 0, -x[3]], [-x[4], x[3], 0]])
        xi = np.zeros((4, 4))
        xi[:3, :3] = wx
        xi[:3, 3] = x[:3]
        assert np.allclose(T, expm(xi), atol=1e-12)
    # tiny rotation: translation passes through, rotation = identity (manifolds.hpp:41-44)
    T = orc.se3_exp([0.1, 0.2, 0.3, 1e-8, 0, 0])
    assert np.array_equal(T[:3, :3], np.eye(3)) and np.allclose(T[:3, 3], [0.1, 0.2, 0.3])
    # T2SE3 re-orthonormalises
    T = orc.se3_exp([0, 0, 0, 0.3, -0.2, 0.9])
    Tn = T.copy()
    Tn[:3, :3] += 1e-6 * rng.randn(3, 3)
    R = orc.t2se3(Tn)[:3, :3]
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-14)
    assert np.allclose(R, T[:3, :3], atol=1e-5)
