"""Parity at BASELINE.json's full sizes (the workloads bench.py measures): the oracle still finishes in seconds on
C1-C3, so the CUDA path is compared with it directly (poses <= 1e-4 m / 1e-4 rad, same convergence flag and iteration
counts), plus size-independent properties: voxel-downsample membership / idempotence on millions of points,
bit-reproducibility of a call, registration towards the known synthetic truth, and batch == singles on the C4-style map."""
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi, workloads

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-4, 1e-4  # north_star: final poses within 1e-4 m / 1e-4 rad


@pytest.fixture(scope="module")
def ctx_any():
    c = capi.Context(capi.PCR_LOAM)
    yield c
    c.close()


def test_c1_loam_full_size(ctx_any):
    wl = workloads.c1_loam(lambda p, leaf: ctx_any.voxel_downsample(p, leaf), 3)
    assert 150_000 < len(wl["dst"]) < 260_000  # "~200k-point local submap"
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(wl["dst"])
    for s, Tg, Tt in zip(wl["scans"], wl["guesses"], wl["truths"]):
        o = orc.loam_align(s, wl["dst"], Tg, threads=8)
        T, conv = c.align(s, Tg)
        assert conv == o["converged"] and c.stats()["iterations"] == len(o["iters"])
        dt, dr = data.pose_err(T, o["T"])
        assert dt < TOL_T and dr < TOL_R, (dt, dr)
        T2, _ = c.align(s, Tg)
        assert np.array_equal(T, T2), "a registration must be bit-reproducible"
        et, er = data.pose_err(T, Tt)
        assert et < 0.05 and er < 5e-3, "registration should land on the synthetic truth"
    # first linearisation at full size: gate decisions and kNN indices bit-exact, normal equations 1e-6
    s, Tg = wl["scans"][0], wl["guesses"][0]
    g, o = c.loam_linearize(s, Tg), orc.loam_linearize(s, wl["dst"], Tg)
    gate = o["status"] >= 1
    assert np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["knn_idx"][gate], o["knn_idx"][gate].astype(np.int32))
    assert data.rel_err(g["JtJ"], o["JtJ"]) < 1e-6 and data.rel_err(g["JtE"], o["JtE"]) < 1e-6
    c.close()


def test_c2_ndt_full_size(ctx_any):
    wl = workloads.c2_ndt(lambda p, leaf: ctx_any.voxel_downsample(p, leaf), 2)
    assert 1_500_000 < len(wl["dst"]) < 3_000_000 and 100_000 < len(wl["scans"][0]) < 140_000
    # voxel downsample at map scale: the GPU output is exactly the oracle's (keys, membership, float32 centroids)
    sub = wl["dst"][: 400_000]
    assert np.array_equal(ctx_any.voxel_downsample(sub, 1.0).view(np.uint32), orc.voxel_downsample(sub, 1.0)["points"].view(np.uint32))
    again = ctx_any.voxel_downsample(wl["dst"], 0.2)  # idempotence up to centroid motion: never more voxels than points,
    assert len(again) <= len(wl["dst"]) and len(again) > 0.9 * len(wl["dst"])  # and a 0.2 m re-grid keeps almost all of them
    c = capi.Context(capi.PCR_NDT)
    c.set_target(wl["dst"])
    ondt = orc.Ndt(wl["dst"], 1.0)
    for s, Tg in zip(wl["scans"], wl["guesses"]):
        o = ondt.align(s, Tg, threads=8)
        T, conv = c.align(s, Tg)
        st = c.stats()
        assert conv == o["converged"] and st["iterations"] == o["nr_iterations"] and st["evaluations"] == o["n_derivative_evals"]
        dt, dr = data.pose_err(T, o["T"])
        assert dt < TOL_T and dr < TOL_R, (dt, dr)
    c.close()


def test_c3_vgicp_full_size():
    wl = workloads.c3_vgicp(1)
    p = wl["pairs"][0]
    assert len(p["src"]) > 150_000 and len(p["dst"]) > 150_000
    c = capi.Context(capi.PCR_VGICP)
    c.set_target(p["dst"])
    # exact 20-NN on the raw, density-skewed 128-beam scan: indices bit-exact on a sample of queries
    covs, idx = c.gicp_covariances(p["dst"], 20, want_idx=True)
    sample = np.random.RandomState(0).choice(len(p["dst"]), 3000, replace=False)
    oi, _ = orc.knn(p["dst"], p["dst"][sample, :3].astype(np.float64), 20, metric_float=True, cell=0.5, threads=8)
    assert np.array_equal(idx[sample], oi.astype(np.int32))
    c.set_target(p["dst"])
    T, conv = c.align(p["src"], p["T_guess"])
    o = orc.Vgicp(p["dst"], 1.0, 20, threads=8).align(p["src"], p["T_guess"], threads=8)
    dt, dr = data.pose_err(T, o["T"])
    assert conv == o["converged"] and dt < TOL_T and dr < TOL_R, (dt, dr)
    et, er = data.pose_err(T, p["T_true"])
    assert et < 0.05 and er < 5e-3
    # fitness (getFitnessScore): mean squared 1-NN distance of the aligned scan, against the oracle's exact 1-NN
    f = c.fitness()
    Tf = T.astype(np.float32)
    q = (p["src"][:, :3] @ Tf[:3, :3].T + Tf[:3, 3]).astype(np.float32)
    _, d2 = orc.knn(p["dst"], q.astype(np.float64), 1, metric_float=True, cell=1.0, threads=8)
    assert abs(d2.mean() - f) <= 1e-9 * f, (d2.mean(), f)  # same float 1-NN distances, FP64 mean
    c.close()


def test_c4_style_batch_on_dense_map(ctx_any):
    """dense (0.2 m) static map -> the half-gate cell grid + second ring path of the LOAM search; batch == singles"""
    wl = workloads.c4_batched("loam", lambda p, leaf: ctx_any.voxel_downsample(p, leaf), 6, tiles=1)
    assert len(wl["dst"]) > 800_000
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(wl["dst"])
    offs = np.concatenate([[0], np.cumsum([len(s) for s in wl["scans"]])])
    bT, bconv = c.batch_align(np.concatenate(wl["scans"]), offs, wl["guesses"])
    for s, Tg, T, conv in zip(wl["scans"], wl["guesses"], bT, bconv):
        sT, sconv = c.align(s, Tg)
        assert sconv == conv and np.allclose(sT, T, rtol=0, atol=1e-9)
    for k in (0, 3):
        o = orc.loam_align(wl["scans"][k], wl["dst"], wl["guesses"][k], threads=8)
        dt, dr = data.pose_err(bT[k], o["T"])
        assert bool(bconv[k]) == o["converged"] and dt < TOL_T and dr < TOL_R, (dt, dr)
    s, Tg = wl["scans"][1], wl["guesses"][1]
    g, o = c.loam_linearize(s, Tg), orc.loam_linearize(s, wl["dst"], Tg)
    gate = o["status"] >= 1
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["knn_idx"][gate], o["knn_idx"][gate].astype(np.int32))
    c.close()
