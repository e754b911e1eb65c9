"""Every LOAM kernel variant against the oracle — in particular the one the batched benchmarks run: one lane per query,
full 32-query tiles, the iteration split into the search kernel + the fit kernel (`pick_shape` selects it for >= ~150 k
queries; here it is forced through the PCR_LOAM_LPQ / PCR_LOAM_TILE knobs, and pcr_loam_last_shape proves which variant ran).
kNN indices and gate / plane / weight decisions bit-exact, per-iteration JtJ / JtE <= 1e-6 relative, final poses
<= 1e-4 m / 1e-4 rad (LoamRegister.cpp:99-223)."""
import os
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi, workloads

pytestmark = pytest.mark.gpu
TOL_REL, TOL_T, TOL_R = 1e-6, 1e-4, 1e-4

# (lanes per query, tile, split) -> environment
VARIANTS = {
    "lpq1_split": dict(PCR_LOAM_LPQ="1", PCR_LOAM_TILE="32", PCR_LOAM_SPLIT="1"),   # the benchmarked batch path
    "lpq1_split_staged": dict(PCR_LOAM_LPQ="1", PCR_LOAM_TILE="32", PCR_LOAM_SPLIT="1", PCR_LOAM_STAGE="1", PCR_LOAM_PREFETCH="1"),  # cp.async staging + L1 prefetch
    "lpq1_fused": dict(PCR_LOAM_LPQ="1", PCR_LOAM_TILE="32", PCR_LOAM_SPLIT="0"),
    "lpq2": dict(PCR_LOAM_LPQ="2", PCR_LOAM_TILE="32"),
    "lpq4": dict(PCR_LOAM_LPQ="4", PCR_LOAM_TILE="16"),
    "lpq8": dict(PCR_LOAM_LPQ="8", PCR_LOAM_TILE="4"),
}
KNOBS = ("PCR_LOAM_LPQ", "PCR_LOAM_TILE", "PCR_LOAM_SPLIT", "PCR_LOAM_STAGE", "PCR_LOAM_PREFETCH")


@pytest.fixture
def variant(request):
    saved = {k: os.environ.get(k) for k in KNOBS}
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(VARIANTS[request.param])
    yield request.param
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _expect_shape(ctx, name):
    sh = ctx.loam_last_shape()
    want = VARIANTS[name]
    assert sh["lpq"] == int(want["PCR_LOAM_LPQ"]) and sh["tile"] == int(want["PCR_LOAM_TILE"]), sh
    assert sh["split"] == (want.get("PCR_LOAM_SPLIT") == "1"), sh


def _check_linearize(ctx, src, dst, T, threads=8, normal_equations=True):
    o = orc.loam_linearize(src, dst, T, threads=threads)
    g = ctx.loam_linearize(src, T)
    gate = o["status"] >= 1
    assert np.array_equal(g["status"], o["status"]), "gate / plane / weight decisions differ"
    assert np.array_equal(g["knn_idx"][gate], o["knn_idx"][gate].astype(np.int32)), "kNN indices differ on accepted queries"
    assert (g["knn_idx"][~gate] == -1).all()
    assert g["n"] == o["n"]
    if normal_equations:
        assert data.rel_err(g["JtJ"], o["JtJ"]) < TOL_REL and data.rel_err(g["JtE"], o["JtE"]) < TOL_REL
    return int(gate.sum())


def _check_align_logs(ctx, src, dst, Tg, threads=8):
    o = orc.loam_align(src, dst, Tg, threads=threads)
    T, conv = ctx.align(src, Tg)
    logs = ctx.loam_logs()
    assert conv == o["converged"] and len(logs) == len(o["iters"])
    for gl, ol in zip(logs, o["iters"]):
        assert gl["n"] == ol["n"]
        assert data.rel_err(gl["JtJ"], ol["JtJ"]) < TOL_REL and data.rel_err(gl["JtE"], ol["JtE"]) < TOL_REL
        assert np.allclose(gl["x"], ol["x"], atol=1e-8)
    dt, dr = data.pose_err(T, o["T"])
    assert dt < TOL_T and dr < TOL_R, (dt, dr)


@pytest.mark.parametrize("variant", list(VARIANTS), indirect=True)
def test_variant_sparse_map(variant):
    """C1-like submap (cells = gate radius, one ring)"""
    case = data.loam_case()
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(case["dst"])
    assert _check_linearize(c, case["src"], case["dst"], case["T_guess"]) > 1000
    _expect_shape(c, variant)
    _check_linearize(c, case["src"], case["dst"], case["T_true"])
    _check_align_logs(c, case["src"], case["dst"], case["T_guess"])
    _expect_shape(c, variant)
    c.close()


@pytest.mark.parametrize("variant", list(VARIANTS), indirect=True)
def test_variant_quantised_ties(variant):
    """exact distance ties: the (d2, index) tie-break must match bit for bit in every variant. Only indices and decisions
    are compared here: on this lattice many 5-neighbourhoods are collinear, the plane fit is rank deficient with exactly
    tied column norms, and which basic solution the pivoted QR returns then hinges on the last bit of those norms
    (FMA contraction on the device, none in the gcc-built oracle) — the normal equations are compared on the
    non-degenerate cases of this file instead."""
    rng = np.random.RandomState(0)
    dst = data.xyzi((np.round(rng.uniform(-8, 8, (20000, 3)) / 0.25) * 0.25 * [1, 1, 0.05]).astype(np.float32))
    src = data.xyzi((np.round(rng.uniform(-7, 7, (3000, 3)) / 0.125) * 0.125 * [1, 1, 0.05]).astype(np.float32))
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(dst)
    assert _check_linearize(c, src, dst, np.eye(4), normal_equations=False) > 1000
    _expect_shape(c, variant)
    c.close()


@pytest.fixture(scope="module")
def dense():
    """C4-style dense map (0.2 m surfaces, one 200 m tile): half-gate cells, second ring on demand"""
    c = capi.Context(capi.PCR_LOAM)
    wl = workloads.c4_batched("loam", lambda p, leaf: c.voxel_downsample(p, leaf), 10, tiles=1)
    c.close()
    return wl


@pytest.mark.parametrize("variant", list(VARIANTS), indirect=True)
def test_variant_dense_map(variant, dense):
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(dense["dst"])
    for k in (0, 1):
        _check_linearize(c, dense["scans"][k], dense["dst"], dense["guesses"][k])
    _expect_shape(c, variant)
    _check_align_logs(c, dense["scans"][2], dense["dst"], dense["guesses"][2])
    c.close()


@pytest.mark.parametrize("variant", ["lpq1_split", "lpq2"], indirect=True)
@pytest.mark.parametrize("which", ["sparse", "dense"])
def test_variant_batch_of_scans(variant, which, dense):
    """a >= 8-scan batch through pcr_batch_align in the forced variant: every pose, convergence flag and scan 0's
    per-iteration normal equations against the oracle"""
    if which == "sparse":
        case = data.loam_case()
        rng = np.random.RandomState(11)
        dst = case["dst"]
        scans, guesses = [], []
        for k in range(9):
            keep = rng.rand(len(case["src"])) < rng.uniform(0.6, 1.0)
            scans.append(np.ascontiguousarray(case["src"][keep]))
            pert = np.concatenate([rng.uniform(-0.3, 0.3, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-2, 2, 3)) * [0.3, 0.3, 1]])
            guesses.append(case["T_true"] @ workloads.synth.se3_exp(pert))
    else:
        dst, scans, guesses = dense["dst"], dense["scans"][:10], dense["guesses"][:10]
    c = capi.Context(capi.PCR_LOAM)
    c.set_target(dst)
    offs = np.concatenate([[0], np.cumsum([len(s) for s in scans])])
    bT, bconv = c.batch_align(np.concatenate(scans), offs, guesses)
    _expect_shape(c, variant)
    logs = c.loam_logs()
    for k, (s, Tg) in enumerate(zip(scans, guesses)):
        o = orc.loam_align(s, dst, Tg, threads=8)
        dt, dr = data.pose_err(bT[k], o["T"])
        assert bool(bconv[k]) == o["converged"] and dt < TOL_T and dr < TOL_R, (k, dt, dr)
        if k == 0:
            assert len(logs) == len(o["iters"])
            for gl, ol in zip(logs, o["iters"]):
                assert gl["n"] == ol["n"]
                assert data.rel_err(gl["JtJ"], ol["JtJ"]) < TOL_REL and data.rel_err(gl["JtE"], ol["JtE"]) < TOL_REL
    c.close()
