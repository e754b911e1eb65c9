"""the headless frontend loop (LidarOdometry::generateOdom + MapManager) and the loop-closure
verification, GPU path (simpleslam_b200/frontend.py over the C ABI) against the CPU restatement (oracle/pyfrontend.py)."""
import numpy as np
import pytest
import data
from oracle import pyfrontend as opf
from oracle import pyoracle as orc
from simpleslam_b200 import capi, frontend, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seq():
    return workloads.c5_sequence(45)


def test_submap_assembly_bit_exact(seq):
    """transform (float) + concat + 0.5 m downsample of keyframes on the device == the oracle's, record for record"""
    fr = seq["frames"]
    clouds = [np.ascontiguousarray(fr[i]["scan"]) for i in (0, 4, 9, 13)]
    poses = [fr[i]["truth"] for i in (0, 4, 9, 13)]
    c = capi.Context(capi.PCR_LOAM)
    pts, m = c.submap_build(clouds, poses, 0.5)
    ref = orc.voxel_downsample(np.concatenate([opf.transform_cloud_f32(cl, T) for cl, T in zip(clouds, poses)]), 0.5)["points"]
    assert m == len(ref) and np.array_equal(pts.view(np.uint32), ref.view(np.uint32))
    # the assembled submap is the register's target: aligning against it == aligning against the host copy
    src = orc.voxel_downsample(fr[5]["scan"], 0.5)["points"]
    T1, c1 = c.align(src, fr[5]["truth"])
    c2 = capi.Context(capi.PCR_LOAM)
    T2, cc2 = c2.scan2map(src, ref, fr[5]["truth"])
    assert c1 == cc2 and np.array_equal(T1, T2)
    # cached keyframes (same host arrays) and an empty keyframe list
    pts2, m2 = c.submap_build(clouds[:2] + [np.zeros((0, 8), np.float32)], poses[:3], 0.5)
    ref2 = orc.voxel_downsample(np.concatenate([opf.transform_cloud_f32(cl, T) for cl, T in zip(clouds[:2], poses[:2])]), 0.5)["points"]
    assert np.array_equal(pts2.view(np.uint32), ref2.view(np.uint32))
    _, m0 = c.submap_build([], [], 0.5)
    assert m0 == 0
    Te, ce = c.align(src, fr[5]["truth"])  # empty submap: LOAM finds fewer than 6 residuals -> not converged, pose kept
    assert not ce and np.allclose(Te, fr[5]["truth"], atol=1e-9)
    c.close(); c2.close()


@pytest.mark.parametrize("pcr_type", ["loam", "ndt"])
def test_odometry_loop_parity(seq, pcr_type):
    n = 45 if pcr_type == "loam" else 14
    lo = frontend.LidarOdometry(pcr_type)
    oo = opf.OracleOdometry(pcr_type)
    worst = (0.0, 0.0)
    for f in seq["frames"][:n]:
        P = lo.generateOdom(f["scan"], f["stamp"], f["local_odom"])
        Q = oo.step(f["scan"], f["stamp"], f["local_odom"])
        dt, dr = data.pose_err(P, Q)
        worst = (max(worst[0], dt), max(worst[1], dr))
        assert dt < 1e-4 and dr < 1e-4, (pcr_type, f["stamp"], dt, dr)
    assert lo.converged == oo.converged
    assert len(lo.map.keyframes) == len(oo.kfs) and lo.map.submap_idx == oo.submap_idx
    assert lo.map.submap_size == len(oo.submap) and lo.map.n_updates == len(oo.submaps)
    if pcr_type == "loam":  # the loop tracks the synthetic truth (NDT's 0.05-0.1 step clamp on a sparse 0.5 m submap does not)
        assert np.linalg.norm(P[:2, 3] - seq["frames"][n - 1]["truth"][:2, 3]) < 0.3
    lo.close()


def test_six_dof_to_mobile():
    rng = np.random.RandomState(0)
    for _ in range(50):
        w = rng.normal(size=3) * [0.02, 0.02, 1.5]
        from simpleslam_b200 import synth
        T = synth.se3_exp(np.concatenate([rng.normal(size=3) * 5, w]))
        A, B = frontend.six_dof_to_mobile(T), opf.mobile_pose(T)
        assert np.allclose(A, B, atol=1e-12) and A[2, 3] == 0 and abs(np.linalg.det(A[:3, :3]) - 1) < 1e-12
    tilted = synth.se3_exp([1, 2, 3, 1.0, 0.2, 0.1])  # axis far from z: rotation dropped
    assert np.allclose(frontend.six_dof_to_mobile(tilted)[:3, :3], np.eye(3))


def test_loop_closure_verification(seq):
    """candidate verification: history submap of the old keyframe, VGICP in LC mode from the current pose, fitness gate"""
    fr = seq["frames"]
    kfs = [(np.ascontiguousarray(fr[i]["scan"]), fr[i]["truth"]) for i in range(0, 40, 4)]
    # perturb the pose of the "current" keyframe a little: a revisit with odometry drift
    from simpleslam_b200 import synth
    cur = len(kfs) - 1
    kfs[cur] = (kfs[cur][0], kfs[cur][1] @ synth.se3_exp([0.15, -0.1, 0.0, 0, 0, np.deg2rad(0.8)]))
    v = frontend.LoopClosureVerifier(kfs)
    g = v.verify(cur - 1, cur)
    o = opf.verify_loop(kfs, cur - 1, cur)
    dt, dr = data.pose_err(g["T"], o["T"])
    assert g["map_points"] == o["map_points"] and g["converged"] == o["converged"] and dt < 1e-4 and dr < 1e-4
    assert abs(g["fitness"] - o["fitness"]) <= 1e-6 * max(o["fitness"], 1e-12) and g["accepted"] == o["accepted"]
    # a far-apart pair must be rejected by the fitness gate
    far = v.verify(0, cur)
    assert not far["accepted"]
    v.close()


def test_frontend_and_scancontext_against_golden():
    """GPU loop and descriptors against the committed golden vectors (tests/golden/frontend_small.npz)"""
    import os
    g = np.load(os.path.join(data.GOLDEN, "frontend_small.npz"))
    offs = g["scan_offsets"]
    scans = [data.xyzi(g["scans"][offs[k]:offs[k + 1]]) for k in range(len(offs) - 1)]
    lo = frontend.LidarOdometry("loam")
    for k, s in enumerate(scans):
        P = lo.generateOdom(s, float(g["stamps"][k]), g["local_odom"][k])
        dt, dr = data.pose_err(P, g["poses"][k])
        assert dt < 1e-4 and dr < 1e-4, (k, dt, dr)
    assert lo.converged == [bool(x) for x in g["converged"]] and len(lo.map.keyframes) == int(g["n_keyframes"])
    assert lo.map.submap_size == int(g["submap_sizes"][-1])
    ds = [lo.ctx.voxel_downsample(s, 0.5) for s in scans[:4]]
    desc, _, _ = lo.ctx.scancontext_make(ds, 2.0)
    assert np.array_equal(desc, g["sc_desc"])
    dist, shift = lo.ctx.scancontext_distance(desc, [(0, 1)], 0.1, False)
    assert abs(dist[0] - g["sc_dist"][0]) < 1e-12 and shift[0] == g["sc_shift"][0]
    dist, shift = lo.ctx.scancontext_distance(desc, [(0, 3)], 0.1, True)
    assert abs(dist[0] - g["sc_dist"][1]) < 1e-12 and shift[0] == g["sc_shift"][1]
    lo.close()


def test_loop_closure_manager_pipeline():
    """ScanContext candidate -> history submap -> VGICP (LC mode) -> fitness gate, GPU pipeline against the CPU restatement"""
    seq = workloads.c5_sequence(60, speed=30.0)
    kfs = [(np.ascontiguousarray(f["scan"]), f["truth"]) for f in seq["frames"]]
    kfs.append((kfs[7][0].copy(), kfs[7][1] @ synth_exp([0.1, -0.05, 0.0, 0, 0, np.deg2rad(0.5)])))   # a revisit of keyframe 7 with drift
    lcm = frontend.LoopClosureManager(kfs)
    lcm.addContext()
    loops = lcm.lcHandler()
    oloops, ochecked = opf.loop_closure_pass(kfs)
    assert [(a, b) for a, b, _ in loops] == [(a, b) for a, b, _ in oloops]
    assert len(lcm.checked) == len(ochecked) >= 1
    for g, o in zip(lcm.checked, ochecked):
        assert (g["old"], g["cur"]) == (o["old"], o["cur"]) and g["converged"] == o["converged"] and g["accepted"] == o["accepted"]
        dt, dr = data.pose_err(g["T"], o["T"])
        assert dt < 1e-4 and dr < 1e-4 and abs(g["fitness"] - o["fitness"]) <= 1e-6 * max(o["fitness"], 1e-12)
    assert (7, len(kfs) - 1) in [(a, b) for a, b, _ in loops]
    assert lcm.addContext() is None and lcm.lcHandler() == loops      # nothing new: idempotent
    lcm.close()


def synth_exp(x):
    from simpleslam_b200 import synth
    return synth.se3_exp(x)


def test_loop_closure_batch_equals_serial(seq):
    """independent candidates verified concurrently on a pool of VGICP contexts (one stream each; spread over
    the GPUs when there are several) give exactly what one context gives candidate after candidate"""
    import torch
    from simpleslam_b200 import synth
    fr = seq["frames"]
    kfs = [(np.ascontiguousarray(fr[i]["scan"]), fr[i]["truth"] @ synth.se3_exp([0.05 * (i % 3), -0.04, 0.0, 0, 0, np.deg2rad(0.3 * (i % 4))]))
           for i in range(0, 44, 4)]
    pairs = [(k - 1, k) for k in range(2, len(kfs))] + [(0, len(kfs) - 1), (3, 5)]
    serial = frontend.LoopClosureVerifier(kfs)
    ref = [serial.verify(o, c) for o, c in pairs]
    serial.close()
    devs = list(range(min(torch.cuda.device_count(), 4)))
    pool = frontend.LoopClosureVerifier(kfs, workers=4, devices=devs)
    got = pool.verify_batch(pairs)
    pool.close()
    assert [(g["old"], g["cur"]) for g in got] == pairs
    for g, r in zip(got, ref):
        assert g["converged"] == r["converged"] and g["accepted"] == r["accepted"] and g["map_points"] == r["map_points"]
        assert np.array_equal(g["T"], r["T"]) and g["fitness"] == r["fitness"]
    assert any(g["accepted"] for g in got) and not all(g["accepted"] for g in got)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["loam", "ndt", "vgicp"])
def test_downsample_align_equals_the_two_calls(method):
    """pcr_downsample_align (one frame of generateOdom: voxel filter + scan2Map, the downsampled scan stays on the device)
    gives bit for bit the pose of pcr_voxel_downsample followed by pcr_align, and the same number of points."""
    case = {"loam": data.loam_case, "ndt": data.ndt_case, "vgicp": data.vgicp_case}[method]()
    mid = {"loam": capi.PCR_LOAM, "ndt": capi.PCR_NDT, "vgicp": capi.PCR_VGICP}[method]
    ctx = capi.Context(mid)
    ctx.set_target(case["dst"])
    for leaf in (0.2, 0.5):
        ds = ctx.voxel_downsample(case["src"], leaf)
        T2, c2 = ctx.align(ds, case["T_guess"])
        T1, c1, m = ctx.downsample_align(case["src"], leaf, case["T_guess"])
        assert m == len(ds) and c1 == c2
        assert np.array_equal(T1, T2)
    # empty scan: nothing to register, the guess comes back unconverged like pcr_align of an empty cloud
    T0, c0 = ctx.align(np.zeros((0, 8), np.float32), case["T_guess"])
    T1, c1, m = ctx.downsample_align(np.zeros((0, 8), np.float32), 0.5, case["T_guess"])
    assert m == 0 and c1 == c0 and np.array_equal(T0, T1)
