"""The oracle against the committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py) and
closed-form self checks — CPU only."""
import os
import numpy as np
import data
from oracle import pyoracle as orc
from simpleslam_b200 import synth


def _load(name):
    return np.load(os.path.join(data.GOLDEN, name))


def test_voxel_downsample_golden():
    g = _load("loam_small.npz")
    raw = data.xyzi(g["raw"][:, :3], g["raw"][:, 4])
    vd = orc.voxel_downsample(raw, float(g["vd_leaf"]))
    assert np.array_equal(vd["keys"], g["vd_keys"])
    assert np.array_equal(vd["out_keys"], g["vd_out_keys"])
    assert np.array_equal(vd["counts"], g["vd_counts"])
    assert np.array_equal(vd["points"][:, :5].view(np.uint32), g["vd_points"].view(np.uint32))
    # properties: keys ascending, counts sum to n, membership consistent
    assert np.all(np.diff(vd["out_keys"]) > 0)
    assert vd["counts"].sum() == len(raw)
    assert np.array_equal(np.unique(vd["keys"]), vd["out_keys"])


def test_voxel_downsample_edge_cases():
    empty = np.zeros((0, 8), np.float32)
    assert len(orc.voxel_downsample(empty, 0.5)["points"]) == 0
    one = data.xyzi(np.array([[1.5, -2.25, 0.125]], np.float32), [7.0])
    r = orc.voxel_downsample(one, 0.5)
    assert np.array_equal(r["points"][0, :5], np.array([1.5, -2.25, 0.125, 1.0, 7.0], np.float32))
    # PCL: grid overflow (dx*dy*dz > INT32_MAX) -> input returned unchanged
    big = data.xyzi(np.array([[0, 0, 0], [3000, 3000, 3000]], np.float32))
    r = orc.voxel_downsample(big, 0.5)
    assert r["overflow"] and len(r["points"]) == 2


def test_loam_golden():
    g = _load("loam_small.npz")
    src, dst = data.xyzi(g["src"]), data.xyzi(g["dst"])
    lin = orc.loam_linearize(src, dst, g["T_guess"])
    assert np.array_equal(lin["knn_idx"].astype(np.int32), g["lin_knn_idx"])
    assert np.array_equal(lin["status"], g["lin_status"])
    assert lin["n"] == int(g["lin_n"])
    assert data.rel_err(lin["JtJ"], g["lin_JtJ"]) < 1e-12 and data.rel_err(lin["JtE"], g["lin_JtE"]) < 1e-12
    al = orc.loam_align(src, dst, g["T_guess"])
    assert len(al["iters"]) == len(g["it_n"]) and al["converged"] == bool(g["converged"])
    for i, it in enumerate(al["iters"]):
        assert it["n"] == g["it_n"][i]
        assert data.rel_err(it["JtJ"], g["it_JtJ"][i]) < 1e-10
    assert np.allclose(al["T"], g["T_final"], atol=1e-12)
    # closed form: the synthetic truth is recovered
    dt, dr = data.pose_err(al["T"], g["T_true"])
    assert dt < 0.05 and dr < 2e-3


def test_loam_identity_on_exact_copy():
    """A scan that is an exact rigid copy of map points must give zero residuals at the true pose."""
    g = _load("loam_small.npz")
    dst = data.xyzi(g["dst"])
    T = g["T_true"]
    Ti = np.linalg.inv(T)
    sub = dst[::7]
    src = data.xyzi((sub[:, :3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32))
    lin = orc.loam_linearize(src, dst, T)
    assert lin["n"] > 100
    x = np.linalg.solve(lin["JtJ"], -lin["JtE"])
    assert np.linalg.norm(x[:3]) < 5e-2 and np.linalg.norm(x[3:]) < 2e-3


def test_ndt_golden():
    g = _load("ndt_small.npz")
    src, dst = data.xyzi(g["src"]), data.xyzi(g["dst"])
    ndt = orc.Ndt(dst, 1.0)
    lv = ndt.leaves()
    assert np.array_equal(lv["keys"], g["leaf_keys"]) and np.array_equal(lv["npts"], g["leaf_npts"])
    assert np.allclose(lv["mean"], g["leaf_mean"], rtol=0, atol=1e-12)
    assert data.rel_err(lv["icov"], g["leaf_icov"]) < 1e-12
    dv = ndt.derivatives(src, g["p0"])
    assert abs(dv["score"] - float(g["score"])) <= 1e-9 * abs(float(g["score"]))
    assert data.rel_err(dv["g"], g["g"]) < 1e-9 and data.rel_err(dv["H"], g["H"]) < 1e-9
    assert data.rel_err(ndt.hessian(src, g["p0"]), g["H_double"]) < 1e-12
    # float path and double path Hessians agree to float accuracy except the documented +sy/-sy quirk row (tiny here)
    assert data.rel_err(dv["H"], g["H_double"]) < 1e-3
    for k in (0, 4, 5):
        res = ndt.align(src, g["T_guess"][k])
        assert [res["nr_iterations"], res["n_derivative_evals"], res["n_hessian_evals"], int(res["converged"])] == list(g["meta"][k])
        assert np.allclose(res["T"], g["T_final"][k], atol=1e-9)


def test_ndt_more_thuente_inner_loop_is_exercised():
    """full-size case: at least one guess drives the line search into its inner loop -> computeHessian (double path)."""
    c = data.ndt_case()
    ndt = orc.Ndt(c["dst"], 1.0)
    rng = np.random.RandomState(0)
    hess = 0
    for k in range(9):
        pert = np.concatenate([rng.uniform(-0.8, 0.8, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-4, 4, 3)) * [0.3, 0.3, 1]])
        if k != 8:
            continue
        res = ndt.align(c["src"], c["T_true"] @ synth.se3_exp(pert))
        hess += res["n_hessian_evals"]
    assert hess >= 1


def test_vgicp_golden():
    g = _load("vgicp_small.npz")
    src, dst = data.xyzi(g["src"]), data.xyzi(g["dst"])
    scov, sidx = orc.gicp_covariances(src, 20, want_idx=True)
    assert np.array_equal(sidx.astype(np.int32), g["src_knn"])
    assert data.rel_err(scov, g["src_covs"]) < 1e-12
    # PLANE regularisation: eigenvalues (1e-3, 1, 1)
    w = np.linalg.eigvalsh(scov[:50])
    assert np.allclose(w, [1e-3, 1, 1], atol=1e-9)
    vg = orc.Vgicp(dst, 1.0, 20)
    vx = vg.voxels()
    assert np.array_equal(vx["coords"], g["vox_coords"]) and np.array_equal(vx["npts"], g["vox_npts"])
    assert np.allclose(vx["mean"], g["vox_mean"], atol=1e-12)
    lin = vg.linearize(src, scov, g["T_guess"])
    assert lin["n"] == int(g["lin_n"]) and abs(lin["cost"] - float(g["lin_cost"])) < 1e-9 * float(g["lin_cost"])
    assert data.rel_err(lin["H"], g["lin_H"]) < 1e-10 and data.rel_err(lin["b"], g["lin_b"]) < 1e-10
    err = vg.error(src, scov, g["T_guess"], g["Ti"])
    assert abs(err - float(g["err_cost"])) < 1e-9 * float(g["err_cost"])
    res = vg.align(src, g["T_guess"], src_covs=scov)
    assert np.allclose(res["T"], g["T_final"], atol=1e-9) and res["nr_iterations"] == int(g["nr_iterations"])
    assert abs(orc.fitness(src, dst, res["T"]) - float(g["fitness"])) < 1e-9
    dt, dr = data.pose_err(res["T"], g["T_true"])
    assert dt < 0.05 and dr < 5e-3


def test_vgicp_linearity_of_cost_in_weights():
    """size-independent property: evaluating at Ti == T0 through `error` equals the linearize cost."""
    g = _load("vgicp_small.npz")
    src, dst = data.xyzi(g["src"]), data.xyzi(g["dst"])
    vg = orc.Vgicp(dst, 1.0, 20)
    lin = vg.linearize(src, g["src_covs"], g["T_guess"])
    assert abs(vg.error(src, g["src_covs"], g["T_guess"], g["T_guess"]) - lin["cost"]) <= 1e-12 * lin["cost"]
