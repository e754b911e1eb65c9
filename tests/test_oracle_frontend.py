"""CPU checks of the frontend / loop-closure restatement in oracle/pyfrontend.py."""
import numpy as np
from oracle import pyfrontend as opf
from simpleslam_b200 import synth, workloads


def test_mobile_pose_properties():
    T = synth.se3_exp([1.0, -2.0, 0.7, 0.01, -0.02, 0.8])
    M = opf.mobile_pose(T)
    assert M[2, 3] == 0 and np.allclose(M[:2, 3], T[:2, 3]) and np.allclose(M[2, :3], [0, 0, 1]) and np.allclose(M[:3, 2], [0, 0, 1])
    yaw = np.arctan2(M[1, 0], M[0, 0])
    assert abs(yaw - 0.8) < 1e-3
    assert np.allclose(opf.mobile_pose(synth.se3_exp([0, 0, 0, 0.9, 0.1, 0.1]))[:3, :3], np.eye(3))  # axis far from z
    assert np.allclose(opf.mobile_pose(np.eye(4)), np.eye(4))


def test_transform_cloud_f32_matches_double_to_an_ulp():
    rng = np.random.RandomState(0)
    cl = np.zeros((1000, 8), np.float32)
    cl[:, :3] = rng.uniform(-50, 50, (1000, 3))
    cl[:, 4] = rng.rand(1000)
    T = synth.se3_exp([3, -4, 1, 0.1, 0.2, 0.3])
    out = opf.transform_cloud_f32(cl, T)
    ref = cl[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    assert np.max(np.abs(out[:, :3] - ref)) < 2e-5 and np.array_equal(out[:, 4], cl[:, 4])


def test_oracle_odometry_tracks_truth_and_builds_keyframes():
    seq = workloads.c5_sequence(24)
    o = opf.OracleOdometry("loam", threads=4)
    for f in seq["frames"]:
        P = o.step(f["scan"], f["stamp"], f["local_odom"])
    assert all(o.converged)
    assert np.linalg.norm(P[:2, 3] - seq["frames"][-1]["truth"][:2, 3]) < 0.25
    assert 8 <= len(o.kfs) <= 13            # ~12 m travelled, one keyframe per metre
    assert len(o.submaps) >= 8 and len(o.submap) > 5000
    # a frame without local odometry falls back to the last global pose
    f = seq["frames"][-1]
    P2 = o.step(f["scan"], f["stamp"] + 0.1, None)
    assert np.linalg.norm(P2[:2, 3] - P[:2, 3]) < 0.05


def test_frontend_and_scancontext_golden():
    """the oracle's frontend loop and ScanContext reproduce the frozen vectors of tests/golden/frontend_small.npz"""
    import os
    from oracle import pyoracle as orc
    from oracle import pyscancontext as osc
    import data
    g = np.load(os.path.join(data.GOLDEN, "frontend_small.npz"))
    offs = g["scan_offsets"]
    scans = [data.xyzi(g["scans"][offs[k]:offs[k + 1]]) for k in range(len(offs) - 1)]
    oo = opf.OracleOdometry("loam", threads=4)
    for k, s in enumerate(scans):
        P = oo.step(s, float(g["stamps"][k]), g["local_odom"][k])
        assert np.allclose(P, g["poses"][k], rtol=0, atol=1e-9), k
    assert list(g["converged"]) == oo.converged and int(g["n_keyframes"]) == len(oo.kfs)
    assert np.array_equal(g["submap_sizes"], [len(s) for s in oo.submaps])
    for k in range(4):
        assert np.array_equal(osc.make_scancontext(orc.voxel_downsample(scans[k], 0.5)["points"], 2.0), g["sc_desc"][k])
    d01 = osc.distance(g["sc_desc"][0], g["sc_desc"][1])
    d03 = osc.distance(g["sc_desc"][0], g["sc_desc"][3], sector_key_align=True)
    assert abs(d01[0] - g["sc_dist"][0]) < 1e-12 and d01[1] == g["sc_shift"][0]
    assert abs(d03[0] - g["sc_dist"][1]) < 1e-12 and d03[1] == g["sc_shift"][1]
