"""The PRODUCT's small dense linear algebra (dev_linalg.cuh: the __host__ __device__ routines the kernels run — 5x3
column-pivoted QR, 6x6 LDLT, 3x3 eigen / inverse, SE(3) exp, T2SE3 — and host_math.hpp: 6x6 Jacobi-SVD solve, NDT pose /
Euler / angle tables, so3 exp) compiled for the HOST with nvcc and checked on the CPU against numpy / scipy and against
the oracle's independent restatement. No GPU and no product code path through the oracle: this is a unit test."""
import ctypes
import os
import subprocess
import numpy as np
import pytest
from scipy.linalg import expm
from scipy.spatial.transform import Rotation
from oracle import pyoracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "cpp", "libhostmath_shim.so")
SRC = os.path.join(ROOT, "tests", "cpp", "hostmath_shim.cu")


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def shim():
    deps = [SRC, os.path.join(ROOT, "simpleslam_b200", "csrc", "dev_linalg.cuh"), os.path.join(ROOT, "simpleslam_b200", "csrc", "host_math.hpp"),
            os.path.join(ROOT, "simpleslam_b200", "csrc", "ndt_logic.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
        cmd = [nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, SRC]
        if os.path.exists("/usr/bin/g++"):
            cmd += ["-ccbin", "/usr/bin/g++"]
        subprocess.check_call(cmd)
    return ctypes.CDLL(SO)


def test_cpqr_ldlt_svd(shim):
    rng = np.random.RandomState(0)
    for _ in range(300):
        A = np.ascontiguousarray(rng.randn(5, 3) * rng.uniform(0.1, 300) + rng.randn(3) * rng.uniform(0, 300))  # plane points far from the origin
        b = -np.ones(5)
        x, xo = np.empty(3), np.empty(3)
        shim.shim_cpqr5x3(_p(A), _p(b), _p(x))
        orc.lib().orc_test_cpqr5x3(_p(A), _p(b), _p(xo))
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(x, ref, rtol=1e-7, atol=1e-10)
        assert np.allclose(x, xo, rtol=1e-12, atol=1e-15), "product and oracle follow the same Eigen algorithm"
        J = rng.randn(40, 6) * rng.uniform(0.1, 10, size=6)
        H = np.ascontiguousarray(J.T @ J)
        g = rng.randn(6)
        y = np.empty(6)
        shim.shim_ldlt6(_p(H), _p(g), _p(y))
        assert np.allclose(y, np.linalg.solve(H, g), rtol=1e-8, atol=1e-12)
        G = np.ascontiguousarray(-H + 1e-3 * rng.randn(6, 6))
        shim.shim_svd6(_p(G), _p(g), _p(y))
        assert np.allclose(y, np.linalg.solve(G, g), rtol=1e-7, atol=1e-12)
    # rank-deficient 6x6: minimum-norm solution like Eigen::JacobiSVD::solve
    Q = np.linalg.qr(rng.randn(6, 6))[0]
    S = np.ascontiguousarray(Q @ np.diag([5, 3, 2, 1, 0, 0]) @ Q.T)
    g = rng.randn(6)
    y = np.empty(6)
    shim.shim_svd6(_p(S), _p(g), _p(y))
    assert np.allclose(y, np.linalg.pinv(S) @ g, rtol=1e-8, atol=1e-10)


def test_eig_inv_se3(shim):
    rng = np.random.RandomState(1)
    for _ in range(300):
        B = rng.randn(3, 3) * rng.uniform(0.01, 10)
        A = np.ascontiguousarray(B @ B.T + np.diag(rng.uniform(0, 1e-3, 3)))
        w, V = np.empty(3), np.empty(9)
        shim.shim_eig_sym3(_p(A), _p(w), _p(V))
        V = V.reshape(3, 3)
        assert np.all(np.diff(w) >= 0) and np.allclose(w, np.linalg.eigvalsh(A), rtol=1e-9, atol=1e-12 * np.abs(A).max())
        assert np.allclose(V @ np.diag(w) @ V.T, A, rtol=1e-9, atol=1e-11 * np.abs(A).max()) and np.allclose(V.T @ V, np.eye(3), atol=1e-10)
        O = np.empty(9)
        shim.shim_inv3(_p(A), _p(O))
        assert np.allclose(O.reshape(3, 3) @ A, np.eye(3), atol=1e-6)
        k = np.concatenate([rng.randn(3), rng.randn(3) * rng.choice([1e-9, 0.01, 1.0, 3.0])])
        E, Eo = np.empty(16), np.empty(16)
        shim.shim_se3_exp(_p(k), _p(E))
        orc.lib().orc_se3_exp(_p(k), _p(Eo))
        wx = np.array([[0, -k[5], k[4]], [k[5], 0, -k[3]], [-k[4], k[3], 0]])
        tw = np.zeros((4, 4)); tw[:3, :3] = wx; tw[:3, 3] = k[:3]
        theta = np.linalg.norm(k[3:])
        tol = 1e-9 if theta >= 1e-6 else 2 * theta + 1e-12   # below 1e-6 rad the reference returns R = I, t = rho (manifolds.hpp:40-44)
        assert np.allclose(E.reshape(4, 4).T, expm(tw), rtol=1e-9, atol=tol)      # manifolds::exp == the SE(3) exponential
        assert np.allclose(E, Eo, rtol=1e-13, atol=1e-15)
        T = E.copy()
        T[:12] += 1e-6 * rng.randn(12)                                            # slightly non-orthonormal rotation
        shim.shim_t2se3(_p(T))
        R = T.reshape(4, 4).T[:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) > 0.999
        om = k[3:]
        R9 = np.empty(9)
        shim.shim_so3_exp(_p(om), _p(R9))
        assert np.allclose(R9.reshape(3, 3), Rotation.from_rotvec(om).as_matrix(), atol=1e-9)


def test_ndt_pose_euler_and_angle_tables(shim):
    rng = np.random.RandomState(2)
    for _ in range(200):
        p = np.concatenate([rng.randn(3) * 50, rng.uniform(-1.4, 1.4, 3)])
        M = np.empty(16, np.float32)
        shim.shim_ndt_pose(_p(p), _p(M))
        Mm = M.reshape(4, 4).T
        ref = Rotation.from_euler("XYZ", p[3:]).as_matrix()       # Rx * Ry * Rz (intrinsic), ndt_omp_impl.hpp:827-830
        assert np.allclose(Mm[:3, :3], ref, atol=5e-6) and np.allclose(Mm[:3, 3], p[:3].astype(np.float32))
        e = np.empty(3, np.float32)
        shim.shim_euler(_p(np.ascontiguousarray(Mm[:3, :3].astype(np.float32))), _p(e))
        # Eigen's eulerAngles(0, 1, 2) keeps the first angle in [0, pi]: for rx < 0 it returns the equivalent triple
        # (rx + pi, pi - ry, rz + pi); the round trip is therefore checked on the rotation it encodes
        assert -1e-6 <= e[0] <= np.pi + 1e-6
        M2 = np.empty(16, np.float32)
        shim.shim_ndt_pose(_p(np.concatenate([p[:3], e.astype(np.float64)])), _p(M2))
        assert np.allclose(M2.reshape(4, 4).T[:3, :3], Mm[:3, :3], atol=2e-5)
        if p[3] > 1e-3:
            assert np.allclose(e, p[3:], atol=2e-5)
        # angular Jacobian rows are the derivatives of R(p) x with respect to the three angles
        jf = np.empty((8, 3), np.float32); hf = np.empty((15, 3), np.float32); jd = np.empty((8, 3)); hd = np.empty((15, 3))
        shim.shim_angle_tables(_p(p), _p(jf), _p(hf), _p(jd), _p(hd))
        x = rng.randn(3)
        h = 1e-6
        def Rx(q):
            return Rotation.from_euler("XYZ", q).as_matrix() @ x
        d = [(Rx(p[3:] + h * np.eye(3)[a]) - Rx(p[3:] - h * np.eye(3)[a])) / (2 * h) for a in range(3)]
        assert np.allclose([jd[0] @ x, jd[1] @ x], d[0][1:], atol=1e-6) and abs(d[0][0]) < 1e-6     # d/drx: rows 1, 2
        assert np.allclose([jd[2] @ x, jd[3] @ x, jd[4] @ x], d[1], atol=1e-6)                      # d/dry
        assert np.allclose([jd[5] @ x, jd[6] @ x, jd[7] @ x], d[2], atol=1e-6)                      # d/drz
        assert np.allclose(jf, jd.astype(np.float32)) and hd.shape == (15, 3)
        # float-path quirk of the reference: row 6 (d1) of h_ang has +sy in the float table, -sy in the double one (:361, :383)
        assert np.isclose(hf[6, 2], -hd[6, 2].astype(np.float32)) or abs(hd[6, 2]) < 1e-7


def test_host_packer_matches_the_device_pack_layout(shim):
    """hostpack.hpp: `cores` host threads turn AoS records into (x, y, z, intensity | 0) float4 records — the host twin of
    pack_kernel — for every record layout the C ABI accepts, any thread count, ragged sizes, repeated use of one pool."""
    rng = np.random.RandomState(0)
    for width, n in ((8, 100003), (8, 5), (4, 70001), (5, 33333), (3, 4097), (6, 12345)):
        rec = rng.randn(n, width).astype(np.float32)
        want = np.zeros((n, 4), np.float32)
        want[:, :3] = rec[:, :3]
        if width == 8 or width >= 5:
            want[:, 3] = rec[:, 4]
        for threads in (1, 3, 8):
            out = np.full((n, 4), np.nan, np.float32)
            shim.shim_host_pack(_p(rec), ctypes.c_size_t(n), ctypes.c_size_t(width * 4), threads, 3, _p(out))
            assert np.array_equal(out, want), (width, n, threads)
