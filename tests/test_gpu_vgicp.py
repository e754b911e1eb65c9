"""VGICP (V1-V7): CUDA through the C-ABI vs the oracle. k-NN indices exact, covariances / voxel statistics ~1e-9,
cost / H / b <= 1e-6 relative, final pose <= 1e-4 m / 1e-4 rad, fitness score 1e-9."""
import os
import numpy as np
import pytest
import data
from oracle import pyoracle as orc
from simpleslam_b200 import capi, registers

pytestmark = pytest.mark.gpu
TOL_REL = 1e-6
TOL_T, TOL_R = 1e-4, 1e-4


@pytest.fixture(scope="module")
def case():
    return data.vgicp_case()


@pytest.fixture(scope="module")
def ctx(case):
    c = capi.Context(capi.PCR_VGICP)
    c.set_target(case["dst"])
    yield c
    c.close()


def test_knn20_and_covariances_parity(ctx, case):
    oc, oi = orc.gicp_covariances(case["src"], 20, want_idx=True)
    gc, gi = ctx.gicp_covariances(case["src"], 20, want_idx=True)
    assert np.array_equal(gi, oi.astype(np.int32)), "20-NN indices differ"
    assert data.rel_err(gc, oc) < 1e-9
    # ties: quantised cloud
    rng = np.random.RandomState(0)
    q = data.xyzi((np.round(rng.uniform(-5, 5, (5000, 3)) / 0.25) * 0.25).astype(np.float32))
    oc, oi = orc.gicp_covariances(q, 20, want_idx=True)
    gc, gi = ctx.gicp_covariances(q, 20, want_idx=True)
    assert np.array_equal(gi, oi.astype(np.int32))


def test_voxelmap_parity(ctx, case):
    o = orc.Vgicp(case["dst"], 1.0, 20).voxels()
    g = ctx.vgicp_voxels()
    assert np.array_equal(g["coords"], o["coords"]) and np.array_equal(g["npts"], o["npts"])
    assert np.allclose(g["mean"], o["mean"], rtol=0, atol=1e-12)
    assert data.rel_err(g["cov"], o["cov"]) < 1e-9


def test_linearize_and_error_parity(ctx, case):
    scov = orc.gicp_covariances(case["src"], 20)
    ovg = orc.Vgicp(case["dst"], 1.0, 20)
    for T in (case["T_guess"], case["T_true"]):
        o = ovg.linearize(case["src"], scov, T)
        g = ctx.vgicp_evaluate(case["src"], T)
        assert g["n"] == o["n"] and o["n"] > 1000
        assert abs(g["cost"] - o["cost"]) <= TOL_REL * o["cost"]
        assert data.rel_err(g["H"], o["H"]) < TOL_REL and data.rel_err(g["b"], o["b"]) < TOL_REL
    Ti = case["T_guess"].copy()
    Ti[:3, 3] += [0.01, -0.02, 0.005]
    oe = ovg.error(case["src"], scov, case["T_guess"], Ti)
    ge = ctx.vgicp_evaluate(case["src"], case["T_guess"], Ti, want_hb=False)
    assert abs(ge["cost"] - oe) <= TOL_REL * oe


@pytest.mark.parametrize("optimizer", ["LM", "GN"])
def test_align_parity(case, optimizer):
    c = capi.Context(capi.PCR_VGICP, vgicp_optimizer=capi.PCR_LSQ_LM if optimizer == "LM" else capi.PCR_LSQ_GN,
                     vgicp_max_iters=64 if optimizer == "LM" else 20)
    c.set_target(case["dst"])
    o = orc.Vgicp(case["dst"], 1.0, 20).align(case["src"], case["T_guess"], optimizer=optimizer, max_iterations=64 if optimizer == "LM" else 20)
    T, conv = c.align(case["src"], case["T_guess"])
    st = c.stats()
    assert conv == o["converged"] and st["iterations"] == o["nr_iterations"]
    dt, dr = data.pose_err(T, o["T"])
    assert dt < TOL_T and dr < TOL_R
    dt, dr = data.pose_err(T, case["T_true"])
    assert dt < 0.05 and dr < 5e-3
    # V6 fitness
    of = orc.fitness(case["src"], case["dst"], o["T"])
    gf = c.fitness()
    assert abs(gf - of) <= 1e-6 * of
    c.close()


def test_golden():
    g = np.load(os.path.join(data.GOLDEN, "vgicp_small.npz"))
    src, dst = data.xyzi(g["src"]), data.xyzi(g["dst"])
    c = capi.Context(capi.PCR_VGICP)
    c.set_target(dst)
    gc, gi = c.gicp_covariances(src, 20, want_idx=True)
    assert np.array_equal(gi, g["src_knn"]) and data.rel_err(gc, g["src_covs"]) < 1e-9
    vx = c.vgicp_voxels()
    assert np.array_equal(vx["coords"], g["vox_coords"]) and np.array_equal(vx["npts"], g["vox_npts"])
    ev = c.vgicp_evaluate(src, g["T_guess"])
    assert ev["n"] == int(g["lin_n"]) and abs(ev["cost"] - float(g["lin_cost"])) <= TOL_REL * float(g["lin_cost"])
    assert data.rel_err(ev["H"], g["lin_H"]) < TOL_REL and data.rel_err(ev["b"], g["lin_b"]) < TOL_REL
    T, conv = c.align(src, g["T_guess"])
    assert conv == bool(g["converged"])
    dt, dr = data.pose_err(T, g["T_final"])
    assert dt < TOL_T and dr < TOL_R
    assert abs(c.fitness() - float(g["fitness"])) <= 1e-6 * float(g["fitness"])
    c.close()


def test_register_interface_lc_mode(case):
    """VgicpRegister::initForLC + scan2Map + getFitnessScore as used by LoopClosureManager.cpp:98-106"""
    reg = registers.make_register("vgicp")
    reg.initForLC()
    res = case["T_guess"].copy()
    ok = reg.scan2Map(case["src"], case["dst"], res)
    o = orc.Vgicp(case["dst"], 1.0, 20).align(case["src"], case["T_guess"], max_iterations=100, trans_eps=1e-6)
    assert ok == o["converged"]
    dt, dr = data.pose_err(res, o["T"])
    assert dt < TOL_T and dr < TOL_R
    fs = reg.getFitnessScore()
    assert abs(fs - orc.fitness(case["src"], case["dst"], o["T"])) <= 1e-6 * fs
