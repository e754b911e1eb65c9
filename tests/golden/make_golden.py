"""Generates tests/golden/*.npz from the CPU oracle (run here, where /root/reference is mounted, so the kNN part is
additionally pinned against the reference's own nanoflann built into oracle/_ref).

    python tests/golden/make_golden.py

PARITY UNPINNED by the reference's tests (it ships no fixtures): these vectors are the oracle's own
outputs, frozen so that (a) the oracle cannot drift silently and (b) the CUDA path is compared with a fixed answer.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as orc  # noqa: E402
import data  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sub(pts, n, seed):
    rng = np.random.RandomState(seed)
    idx = np.sort(rng.choice(len(pts), size=min(n, len(pts)), replace=False))
    return np.ascontiguousarray(pts[idx])


def main():
    orc.build()
    # ---- voxel downsample + LOAM -----------------------------------------------------------------
    c = data.loam_case()
    raw = sub(c["raw"], 6000, 0)
    vd = orc.voxel_downsample(raw, 0.5)
    src = sub(c["src"], 2500, 1)
    # keep the map dense around the scan so enough residuals survive
    d = c["dst"]
    ctr = c["T_true"][:3, 3]
    near = np.linalg.norm(d[:, :2] - ctr[:2], axis=1) < 45.0
    dst = np.ascontiguousarray(d[near])
    lin = orc.loam_linearize(src, dst, c["T_guess"])
    al = orc.loam_align(src, dst, c["T_guess"])
    q = (src[:, :3].astype(np.float64) @ c["T_guess"][:3, :3].T + c["T_guess"][:3, 3]).astype(np.float32).astype(np.float64)
    ref = orc.ref_knn(dst, q, 5)
    if ref is not None:
        gi, gd = orc.knn(dst, q, 5)
        assert np.array_equal(gd, ref[1]), "oracle kNN distances differ from the reference's nanoflann"
        ties = int((gi != ref[0]).any(1).sum())
        print("kNN vs reference nanoflann: %d/%d rows differ (ties only)" % (ties, len(q)))
    np.savez_compressed(
        os.path.join(OUT, "loam_small.npz"), raw=raw[:, :5], vd_leaf=np.float32(0.5), vd_keys=vd["keys"], vd_out_keys=vd["out_keys"],
        vd_counts=vd["counts"], vd_points=vd["points"][:, :5], src=src[:, :3], dst=dst[:, :3], T_guess=c["T_guess"], T_true=c["T_true"],
        lin_knn_idx=lin["knn_idx"].astype(np.int32), lin_status=lin["status"], lin_JtJ=lin["JtJ"], lin_JtE=lin["JtE"], lin_n=lin["n"],
        it_JtJ=np.stack([i["JtJ"] for i in al["iters"]]), it_JtE=np.stack([i["JtE"] for i in al["iters"]]),
        it_x=np.stack([i["x"] for i in al["iters"]]), it_n=np.array([i["n"] for i in al["iters"]]),
        it_T=np.stack([i["T_before"] for i in al["iters"]]), T_final=al["T"], converged=al["converged"])
    print("loam_small: src", src.shape, "dst", dst.shape, "iters", len(al["iters"]), "converged", al["converged"], "n", [i["n"] for i in al["iters"]])

    # ---- NDT ----------------------------------------------------------------------------------------
    c = data.ndt_case()
    src = sub(c["src"], 4000, 2)
    ctr = c["T_true"][:3, 3]
    d = c["dst"]
    near = np.linalg.norm(d[:, :2] - ctr[:2], axis=1) < 30.0
    dst = sub(np.ascontiguousarray(d[near]), 60000, 3)
    ndt = orc.Ndt(dst, 1.0)
    lv = ndt.leaves()
    # several initial guesses: short / long Newton runs and one that enters the More-Thuente inner loop (computeHessian)
    rng = np.random.RandomState(0)
    guesses, finals, meta = [], [], []
    for k in range(12):
        pert = np.concatenate([rng.uniform(-0.8, 0.8, 3) * [1, 1, 0.2], np.deg2rad(rng.uniform(-4, 4, 3)) * [0.3, 0.3, 1]])
        Tg = c["T_true"] @ data.synth.se3_exp(pert)
        res = ndt.align(src, Tg)
        guesses.append(Tg); finals.append(res["T"])
        meta.append([res["nr_iterations"], res["n_derivative_evals"], res["n_hessian_evals"], int(res["converged"])])
    meta = np.array(meta)
    p0 = np.array([c["T_guess"][0, 3], c["T_guess"][1, 3], c["T_guess"][2, 3], 0.01, -0.02, 0.33])
    dv = ndt.derivatives(src, p0)
    hs = ndt.hessian(src, p0)
    np.savez_compressed(
        os.path.join(OUT, "ndt_small.npz"), src=src[:, :3], dst=dst[:, :3], T_true=c["T_true"], leaf_keys=lv["keys"],
        leaf_npts=lv["npts"], leaf_mean=lv["mean"], leaf_icov=lv["icov"], p0=p0, score=dv["score"], g=dv["g"], H=dv["H"], H_double=hs,
        T_guess=np.stack(guesses), T_final=np.stack(finals), meta=meta)
    print("ndt_small: leaves", len(lv["keys"]), "valid", int((lv["npts"] >= 6).sum()), "meta [iters, evals, hess, conv]:")
    print(meta.T)

    # ---- VGICP --------------------------------------------------------------------------------------
    c = data.vgicp_case()
    src = sub(c["src"], 3000, 4)
    dst = sub(c["dst"], 6000, 5)
    scov, sidx = orc.gicp_covariances(src, 20, want_idx=True)
    vg = orc.Vgicp(dst, 1.0, 20)
    vx = vg.voxels()
    lin = vg.linearize(src, scov, c["T_guess"])
    Ti = c["T_guess"].copy()
    Ti[:3, 3] += [0.01, -0.02, 0.005]
    err = vg.error(src, scov, c["T_guess"], Ti)
    res = vg.align(src, c["T_guess"], src_covs=scov)
    fit = orc.fitness(src, dst, res["T"])
    np.savez_compressed(
        os.path.join(OUT, "vgicp_small.npz"), src=src[:, :3], dst=dst[:, :3], T_guess=c["T_guess"], T_true=c["T_true"], src_covs=scov,
        src_knn=sidx.astype(np.int32), vox_coords=vx["coords"], vox_npts=vx["npts"], vox_mean=vx["mean"], vox_cov=vx["cov"], lin_cost=lin["cost"],
        lin_H=lin["H"], lin_b=lin["b"], lin_n=lin["n"], Ti=Ti, err_cost=err, T_final=res["T"], converged=res["converged"],
        nr_iterations=res["nr_iterations"], fitness=fit)
    print("vgicp_small: voxels", len(vx["npts"]), "corr", lin["n"], "iters", res["nr_iterations"], "conv", res["converged"], "fitness", fit,
          "err", data.pose_err(res["T"], c["T_true"]))


if __name__ == "__main__":
    main()
