"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the callers of the PCR hot path:
frontend::LidarOdometry::generateOdom (frontend/src/LidarOdometry.cpp:89-246), MapManager::{setCurPose, putKeyFrame,
updateMap} (frontend/src/MapManager.cpp:109-201), SixDof2Mobile (common/geometry/trans.hpp:68-86) and the loop-closure
verification of backend/src/LoopClosureManager.cpp:40-119, on top of the CPU oracle registers. Written independently of
simpleslam_b200/frontend.py (which it checks); same synchronous threading model and ascending-index keyframe order.
PARITY UNPINNED (the reference ships no fixtures for this path either)."""
import math
import numpy as np
from . import pyoracle as orc

KF_GAP = 1.0
RADIUS = 8.0


def transform_cloud_f32(cloud, pose):
    """pcp::transformPointCloud (pcp.hpp:38-62): tr = pose.cast<float>(), pto = tr * pfrom, evaluated
    ((r0 x + r1 y) + r2 z) + t in float32; intensity and padding copied."""
    tr = np.asarray(pose, dtype=np.float64).astype(np.float32)
    x, y, z = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    out = cloud.copy()
    for r in range(3):
        out[:, r] = ((tr[r, 0] * x + tr[r, 1] * y) + tr[r, 2] * z) + tr[r, 3]
    return out


def mobile_pose(T):
    """SixDof2Mobile: AngleAxis of the rotation; keep x, y and the rotation about +-z when |axis.z| > 0.95"""
    R = T[:3, :3]
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:  # Eigen::Quaternion from a rotation matrix
        s = math.sqrt(tr + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        v = [(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s]
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        v = [0.0, 0.0, 0.0]
        v[i] = 0.5 * s
        s = 0.5 / s
        w = (R[k, j] - R[j, k]) * s
        v[j] = (R[j, i] + R[i, j]) * s
        v[k] = (R[k, i] + R[i, k]) * s
    n = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    out = np.eye(4)
    out[0, 3], out[1, 3] = T[0, 3], T[1, 3]
    if n > 0:
        angle = 2.0 * math.atan2(n, abs(w))
        az = v[2] / (-n if w < 0 else n)
        if abs(az) > 0.95:
            sg = math.copysign(1.0, az)
            c, s = math.cos(angle), math.sin(angle) * sg
            out[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    return out


def register(pcr_type, src, dst, T, threads=8):
    if pcr_type == "loam":
        r = orc.loam_align(src, dst, T, threads=threads)
    elif pcr_type == "ndt":
        r = orc.Ndt(dst, 1.0).align(src, T, threads=threads)
    elif pcr_type == "vgicp":
        r = orc.Vgicp(dst, 1.0, 20, threads=threads).align(src, T, threads=threads)
    else:
        raise RuntimeError("unknown pcr type " + pcr_type)
    return r["T"], bool(r["converged"])


class OracleOdometry:
    def __init__(self, pcr_type="loam", grid_size=0.5, threads=8):
        self.pcr_type, self.grid, self.threads = pcr_type, float(grid_size), threads
        self.kfs = []            # (cloud, pose)
        self.submap = None       # (m, 8) float32 or None
        self.submap_idx = []
        self.cur = np.eye(4)
        self.last = np.eye(4)
        self.want_update = False
        self.last_pos = np.zeros(3)
        self.godom = []
        self.o2m = np.eye(4)
        self.o2m_init = False
        self.converged = []
        self.submaps = []        # history of rebuilt submaps (for parity checks)

    def _put_kf(self, scan, pose):
        if not self.kfs:
            self.kfs.append((scan, pose.copy()))
            return
        d2 = [float(np.sum((kf[1][:3, 3] - pose[:3, 3]) ** 2)) for kf in self.kfs]
        if min(d2) > KF_GAP:
            self.kfs.append((scan, pose.copy()))

    def _update_map(self):
        self.want_update = False
        idx = [i for i, kf in enumerate(self.kfs) if float(np.sum((kf[1][:3, 3] - self.cur[:3, 3]) ** 2)) < RADIUS * RADIUS]
        self.submap_idx = idx
        if idx:
            cat = np.concatenate([transform_cloud_f32(self.kfs[i][0], self.kfs[i][1]) for i in idx])
            self.submap = orc.voxel_downsample(cat, self.grid)["points"]
        else:
            self.submap = np.zeros((0, 8), np.float32)
        self.submaps.append(self.submap)

    def step(self, scan, stamp, local_odom=None):
        init = np.eye(4)
        if local_odom is not None and self.o2m_init:
            init = self.o2m @ local_odom
        elif len(self.godom) >= 2:
            init = self.godom[-1][1].copy()   # frames arrive in stamp order: the closest global odom is the newest
        conv = True
        if self.submap_idx:
            ds = orc.voxel_downsample(scan, self.grid)["points"]
            init, conv = register(self.pcr_type, ds, self.submap, init, self.threads)
        self.converged.append(conv)
        init = mobile_pose(init)
        self.cur = init.copy()
        if np.linalg.norm(self.last[:3, 3] - init[:3, 3]) > KF_GAP:
            self.last = init.copy()
            self.want_update = True
        if not self.submap_idx:
            self._put_kf(scan, init)
            self.want_update = True
        elif np.linalg.norm(init[:3, 3] - self.last_pos) > KF_GAP:
            self._put_kf(scan, init)
            self.last_pos = init[:3, 3].copy()
        self.godom.append((stamp, init.copy()))
        if local_odom is not None:
            self.o2m_init = True
            self.o2m = init @ np.linalg.inv(local_odom)
        if self.want_update:
            self._update_map()
        return init


def verify_loop(kfs, old_key, cur_key, ds=0.5, rng=1, thresh=0.3, threads=8):
    near = [k for k in range(old_key - rng, old_key + rng + 1) if 0 <= k < len(kfs)]
    cat = np.concatenate([transform_cloud_f32(kfs[k][0], kfs[k][1]) for k in near])
    target = orc.voxel_downsample(cat, ds)["points"]
    cloud, pose = kfs[cur_key]
    v = orc.Vgicp(target, 1.0, 20, threads=threads)
    r = v.align(cloud, pose, threads=threads, max_iterations=100, trans_eps=1e-6)
    Tf = r["T"].astype(np.float32)
    q = np.empty((len(cloud), 3), np.float32)
    for a in range(3):
        q[:, a] = ((Tf[a, 0] * cloud[:, 0] + Tf[a, 1] * cloud[:, 1]) + Tf[a, 2] * cloud[:, 2]) + Tf[a, 3]
    _, d2 = orc.knn(target, q.astype(np.float64), 1, metric_float=True, cell=1.0, threads=threads)
    fs = float(d2.mean())
    return dict(converged=bool(r["converged"]), fitness=fs, accepted=bool(r["converged"] and fs < thresh), T=r["T"], map_points=len(target))


def loop_closure_pass(kfs, ds=0.5, rng=1, thresh=0.3, lidar_height=2.0, threads=8, **sc_params):
    """LoopClosureManager::addContext + lcHandler over all keyframes: ScanContext candidates -> VGICP verification"""
    from . import pyscancontext as osc
    sc = osc.OracleScanContext(lidar_height=lidar_height, **sc_params)
    for cloud, _ in kfs:
        sc.add(orc.voxel_downsample(cloud, ds)["points"])
    loops, checked = [], []
    for i in range(len(kfs)):
        old, _ = sc.query(i)
        if old >= 0:
            r = verify_loop(kfs, old, i, ds, rng, thresh, threads)
            checked.append(dict(r, old=old, cur=i))
            if r["accepted"]:
                loops.append((old, i, np.linalg.inv(kfs[old][1]) @ kfs[i][1]))
    return loops, checked
