"""BASELINE config 4 at its real size: the ~20 M-point static map (4x4 tiles of 200 m), LOAM and NDT, a handful of
scans against the oracle (poses <= 1e-4 m / 1e-4 rad, same convergence / iteration counts), first linearisation of the
LOAM search on the full map bit-exact, and the batched call through the kernel variant the benchmark uses."""
import os
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi, workloads

pytestmark = pytest.mark.gpu
TOL_T, TOL_R = 1e-4, 1e-4


@pytest.fixture(scope="module")
def ds_ctx():
    c = capi.Context(capi.PCR_LOAM)
    yield c
    c.close()


def test_c4_loam_full_map(ds_ctx):
    wl = workloads.c4_batched("loam", lambda p, leaf: ds_ctx.voxel_downsample(p, leaf), 8)
    assert len(wl["dst"]) > 15_000_000
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(wl["dst"])
    offs = np.concatenate([[0], np.cumsum([len(s) for s in wl["scans"]])])
    saved = {k: os.environ.get(k) for k in ("PCR_LOAM_LPQ", "PCR_LOAM_TILE")}
    os.environ["PCR_LOAM_LPQ"], os.environ["PCR_LOAM_TILE"] = "1", "32"   # the variant 128-scan batches select by themselves
    try:
        bT, bconv = c.batch_align(np.concatenate(wl["scans"]), offs, wl["guesses"])
        assert c.loam_last_shape() == dict(lpq=1, tile=32, split=True)
        g = c.loam_linearize(wl["scans"][1], wl["guesses"][1])
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    for k in (0, 3, 7):
        o = orc.loam_align(wl["scans"][k], wl["dst"], wl["guesses"][k], threads=8)
        dt, dr = data.pose_err(bT[k], o["T"])
        assert bool(bconv[k]) == o["converged"] and dt < TOL_T and dr < TOL_R, (k, dt, dr)
    o = orc.loam_linearize(wl["scans"][1], wl["dst"], wl["guesses"][1], threads=8)
    gate = o["status"] >= 1
    assert gate.sum() > 500
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["knn_idx"][gate], o["knn_idx"][gate].astype(np.int32))
    assert data.rel_err(g["JtJ"], o["JtJ"]) < 1e-6 and data.rel_err(g["JtE"], o["JtE"]) < 1e-6
    c.close()


def test_c4_ndt_full_map(ds_ctx):
    wl = workloads.c4_batched("ndt", lambda p, leaf: ds_ctx.voxel_downsample(p, leaf), 4)
    assert len(wl["dst"]) > 15_000_000
    c = capi.Context(capi.PCR_NDT)
    c.set_target(wl["dst"])
    offs = np.concatenate([[0], np.cumsum([len(s) for s in wl["scans"]])])
    bT, bconv = c.batch_align(np.concatenate(wl["scans"]), offs, wl["guesses"])
    ondt = orc.Ndt(wl["dst"], 1.0)
    for k in (0, 2):
        o = ondt.align(wl["scans"][k], wl["guesses"][k], threads=8)
        dt, dr = data.pose_err(bT[k], o["T"])
        assert bool(bconv[k]) == o["converged"] and dt < TOL_T and dr < TOL_R, (k, dt, dr)
        T1, c1 = c.align(wl["scans"][k], wl["guesses"][k])
        st = c.stats()
        assert st["iterations"] == o["nr_iterations"] and st["evaluations"] == o["n_derivative_evals"]
        assert np.allclose(T1, bT[k], rtol=0, atol=1e-6)
    c.close()
