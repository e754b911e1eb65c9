"""Closed-form self-checks of the CPU oracle (SURVEY §8c plan, item iii): since the reference ships no fixtures and its
libraries cannot be built here, every analytic piece of the restatement is cross-checked against something independent —
scipy's kd-tree for the exact k-NN, central finite differences of the oracle's own cost functions for the NDT gradient /
Hessian tables (Magnusson eq. 6.12 / 6.13 as coded in ndt_omp_impl.hpp), the VGICP linearisation and the LOAM point-to-plane
Jacobian [I | -p^]."""
import numpy as np
from scipy.spatial import cKDTree
import data
from oracle import pyoracle as orc
from simpleslam_b200 import synth


def test_knn_against_scipy_kdtree():
    rng = np.random.RandomState(3)
    pts = data.xyzi((rng.uniform(-30, 30, (20000, 3)) * [1, 1, 0.1]).astype(np.float32))
    q = rng.uniform(-30, 30, (400, 3)) * [1, 1, 0.1]
    tree = cKDTree(pts[:, :3].astype(np.float64))
    for k in (1, 5, 20):
        idx, d2 = orc.knn(pts, q, k, metric_float=False, cell=1.0)
        dd, ii = tree.query(q, k=k)
        dd, ii = dd.reshape(len(q), k), ii.reshape(len(q), k)
        assert np.allclose(np.sqrt(d2), dd, rtol=1e-12, atol=1e-12)
        distinct = np.all(np.diff(d2, axis=1) > 0, axis=1) if k > 1 else np.ones(len(q), bool)
        assert distinct.mean() > 0.99 and np.array_equal(idx[distinct], ii[distinct])


def _ndt_case():
    rng = np.random.RandomState(5)
    tgt = rng.uniform(-5, 5, (30000, 3)) * [1, 1, 0.3]
    tgt[:, 2] += 0.05 * np.sin(tgt[:, 0]) + 0.03 * tgt[:, 1]      # gently curved slab: anisotropic voxel covariances
    centres = np.stack(np.meshgrid(np.arange(-4, 4) + 0.5, np.arange(-4, 4) + 0.5, [-0.5, 0.5], indexing="ij"), -1).reshape(-1, 3)
    src = centres + rng.uniform(-0.15, 0.15, centres.shape)        # every point stays well inside its voxel under the test poses
    return data.xyzi(tgt.astype(np.float32)), data.xyzi(src.astype(np.float32))


def test_ndt_gradient_and_hessian_match_finite_differences():
    tgt, src = _ndt_case()
    ndt = orc.Ndt(tgt, 1.0)
    p = np.array([0.04, -0.03, 0.02, 0.006, -0.004, 0.008])
    base = ndt.derivatives(src, p)
    assert base["nb"].min() >= 1
    h = 2e-4
    g_fd, H_fd = np.zeros(6), np.zeros((6, 6))
    for i in range(6):
        e = np.zeros(6); e[i] = h
        a, b = ndt.derivatives(src, p + e), ndt.derivatives(src, p - e)
        g_fd[i] = (a["score"] - b["score"]) / (2 * h)
        H_fd[:, i] = (a["g"] - b["g"]) / (2 * h)
    # float32 pair math: the score carries ~1e-7 relative noise, i.e. ~1e-3 of the gradient scale after division by 2h
    assert np.linalg.norm(g_fd - base["g"]) <= 0.02 * np.linalg.norm(base["g"]), (g_fd, base["g"])
    assert np.linalg.norm(H_fd - base["H"]) <= 0.02 * np.linalg.norm(base["H"])
    assert np.allclose(base["H"], base["H"].T, rtol=1e-6, atol=1e-6 * np.abs(base["H"]).max())


def _delta(d):
    """LsqRegistration: delta.linear = so3_exp(d[0:3]), delta.translation = d[3:6]"""
    D = np.eye(4)
    D[:3, :3] = synth.se3_exp(np.concatenate([np.zeros(3), d[:3]]))[:3, :3]
    D[:3, 3] = d[3:]
    return D


def test_vgicp_linearisation_matches_finite_differences():
    case = data.vgicp_case()
    dst, src = case["dst"], case["src"][::3]
    v = orc.Vgicp(dst, 1.0, 20)
    covs = orc.gicp_covariances(src, 20)
    T0 = case["T_guess"].astype(np.float32).astype(np.float64)
    lin = v.linearize(src, covs, T0)
    assert lin["n"] > 1000
    h = 1e-5
    g_fd = np.zeros(6)
    for i in range(6):
        e = np.zeros(6); e[i] = h
        g_fd[i] = (v.error(src, covs, T0, _delta(e) @ T0) - v.error(src, covs, T0, _delta(-e) @ T0)) / (2 * h)
    # cost = sum w e^T M e with correspondences and M frozen at T0: gradient = 2 b, Gauss-Newton Hessian = J^T M J
    assert np.linalg.norm(g_fd - 2 * lin["b"]) <= 1e-5 * np.linalg.norm(2 * lin["b"]) + 1e-6
    H_fd = np.zeros((6, 6))
    hh = 1e-3
    c0 = v.error(src, covs, T0, T0)
    assert abs(c0 - lin["cost"]) <= 1e-9 * abs(c0)
    for i in range(6):
        e = np.zeros(6); e[i] = hh
        H_fd[i, i] = (v.error(src, covs, T0, _delta(e) @ T0) - 2 * c0 + v.error(src, covs, T0, _delta(-e) @ T0)) / hh ** 2
    # translations enter e linearly: exact; rotations differ by the (small) residual-curvature term Gauss-Newton drops
    assert np.allclose(np.diag(H_fd)[3:], 2 * np.diag(lin["H"])[3:], rtol=1e-4)
    assert np.allclose(np.diag(H_fd)[:3], 2 * np.diag(lin["H"])[:3], rtol=3e-2)


def test_loam_jacobian_is_the_point_to_plane_derivative():
    case = data.loam_case()
    src, dst, T = case["src"], case["dst"], case["T_true"]
    base = orc.loam_linearize(src, dst, T)
    acc = base["status"] == 3
    s = np.linalg.norm(base["J"][:, :3], axis=1)          # J = s * [n, p x n] with |n| = 1
    d0 = base["resid"] / np.where(acc, s, 1.0)
    h = 1e-4
    worst = 0.0
    for i in range(6):
        e = np.zeros(6); e[i] = h
        pert = orc.loam_linearize(src, dst, synth.se3_exp(e) @ T)
        same = np.all(np.sort(pert["knn_idx"], axis=1) == np.sort(base["knn_idx"], axis=1), axis=1)   # same five neighbours
        both = acc & (pert["status"] == 3) & same
        assert both.sum() > 0.7 * acc.sum()
        sp = np.linalg.norm(pert["J"][:, :3], axis=1)
        d1 = pert["resid"] / np.where(both, sp, 1.0)
        fd = (d1 - d0)[both] / h
        an = (base["J"][:, i] / np.where(acc, s, 1.0))[both]
        # the map-frame point is rounded to float before the distance is taken: ~1e-5 m of quantisation over h = 1e-4
        err = np.abs(fd - an)
        assert np.median(err) < 0.02 and np.percentile(err, 95) < 0.25, (i, np.median(err), np.percentile(err, 95))
        worst = max(worst, float(np.median(err)))
    assert worst < 0.02


def test_voxel_downsample_against_numpy():
    """pcl::VoxelGrid semantics restated a second time in numpy: float32 key math, keys ascending, float32 centroid sums in
    ascending original index (cumsum is sequential)"""
    rng = np.random.RandomState(9)
    pts = data.xyzi((rng.uniform(-20, 20, (5000, 3)) * [1, 1, 0.2]).astype(np.float32), rng.rand(5000).astype(np.float32))
    leaf = np.float32(0.5)
    o = orc.voxel_downsample(pts, float(leaf))
    inv = np.float32(1.0) / leaf
    xyz = pts[:, :3]
    mn, mx = xyz.min(0), xyz.max(0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int32)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    assert np.array_equal(key, o["keys"])
    order = np.argsort(key, kind="stable")
    uk, start, cnt = np.unique(key[order], return_index=True, return_counts=True)
    assert np.array_equal(uk, o["out_keys"]) and np.array_equal(cnt, o["counts"])
    for v in rng.choice(len(uk), 200, replace=False):
        sel = order[start[v]: start[v] + cnt[v]]
        c = np.cumsum(pts[sel][:, [0, 1, 2, 4]], axis=0, dtype=np.float32)[-1] / np.float32(cnt[v])
        assert np.array_equal(c.view(np.uint32), o["points"][v][[0, 1, 2, 4]].view(np.uint32))
        assert o["points"][v][3] == 1.0


def test_ndt_leaves_against_numpy():
    tgt, _ = _ndt_case()
    L = orc.Ndt(tgt, 1.0).leaves()
    xyz = tgt[:, :3].astype(np.float64)
    ijk = np.floor(tgt[:, :3] * np.float32(1.0)).astype(np.int64) - L["min_b"]
    key = ijk[:, 0] + ijk[:, 1] * L["div_b"][0] + ijk[:, 2] * L["div_b"][0] * L["div_b"][1]
    checked = 0
    for j in np.random.RandomState(0).choice(len(L["keys"]), 60, replace=False):
        p = xyz[key == L["keys"][j]]
        n = len(p)
        if L["npts"][j] < 6:
            assert n == max(L["npts"][j], n)
            continue
        assert n == L["npts"][j]
        mean = p.mean(0)
        assert np.allclose(mean, L["mean"][j], rtol=0, atol=1e-12)
        # single-pass form of voxel_grid_covariance_omp_impl.hpp:329-330 with cov_ started at Identity
        cov = ((np.eye(3) + p.T @ p) - 2 * np.outer(p.sum(0), mean)) / n + np.outer(mean, mean)
        cov *= (n - 1.0) / n
        w, V = np.linalg.eigh(cov)
        if w[0] < 0.01 * w[2]:
            w[0] = 0.01 * w[2]
            if w[1] < 0.01 * w[2]:
                w[1] = 0.01 * w[2]
            cov = V @ np.diag(w) @ np.linalg.inv(V)
        assert np.allclose(cov, L["cov"][j], rtol=1e-9, atol=1e-12)
        assert np.allclose(np.linalg.inv(cov), L["icov"][j], rtol=1e-7, atol=1e-9)
        checked += 1
    assert checked > 20


def test_gicp_covariances_against_numpy():
    rng = np.random.RandomState(4)
    pts = data.xyzi((rng.uniform(-10, 10, (4000, 3)) * [1, 1, 0.05]).astype(np.float32))
    covs, idx = orc.gicp_covariances(pts, 20, want_idx=True)
    xyz = pts[:, :3].astype(np.float64)
    tree = cKDTree(xyz)
    for i in rng.choice(len(pts), 150, replace=False):
        _, nn = tree.query(xyz[i], k=20)
        if set(nn.tolist()) != set(idx[i].tolist()):
            continue  # float-metric tie at the 20th neighbour
        nb = xyz[idx[i]]
        d = nb - nb.mean(0)
        c = d.T @ d / 20.0
        U, _, Vt = np.linalg.svd(c)
        ref = U @ np.diag([1.0, 1.0, 1e-3]) @ Vt        # PLANE regularisation (fast_gicp_impl.hpp:272-293)
        assert np.allclose(covs[i], ref, rtol=1e-6, atol=1e-9), i
