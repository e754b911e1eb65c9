"""Headless restatement of the callers either side of the PCR hot path, without ROS and
without threads, over the C ABI:

  * frontend::LidarOdometry::generateOdom      frontend/src/LidarOdometry.cpp:89-246
  * frontend::MapManager::{setCurPose, putKeyFrame, updateMap}   frontend/src/MapManager.cpp:109-201
  * geometry::trans::SixDof2Mobile             common/geometry/trans.hpp:68-86
  * backend::LoopClosureManager::{loopFindNearKeyframes, lcHandler verification}   backend/src/LoopClosureManager.cpp:40-119

The data-parallel work stays on the GPU: per-frame voxel downsample (pcr_voxel_downsample), submap assembly
(pcr_submap_build: transform + concat + downsample on the device, the result IS the register's target, so the submap
never crosses PCIe), registration (pcr_align against the resident submap), loop-closure VGICP + fitness.

Threading model. The reference rebuilds the submap on its own thread whenever the pose moved more than 1 m
(`notifyUpdateMap`) while the LO thread keeps registering against the old one; which frame first sees the new submap
depends on thread timing. Here the rebuild runs synchronously at the end of the frame that requested it — the
deterministic limit of the reference's behaviour (its offline `USE_BAG` mode with a fast map thread).
Keyframe radius results are taken in ascending keyframe index (nanoflann's unsorted radius search returns tree order).
"""
import numpy as np
from . import capi

MIN_KF_GAP = 1.0          # MapManager::minKFGap / LidarOdometry::minKFGap (MapManager.hpp:67, LidarOdometry.hpp:52)
SUBMAP_RADIUS = 8.0       # MapManager::mSurroundingKeyframeSearchRadius (MapManager.hpp:68)


def rot_to_quat(R):
    """Eigen::Quaternion(Matrix3) (w, x, y, z)"""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([w, (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s])
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[1 + i] = 0.5 * s
    s = 0.5 / s
    q[0] = (R[k, j] - R[j, k]) * s
    q[1 + j] = (R[j, i] + R[i, j]) * s
    q[1 + k] = (R[k, i] + R[i, k]) * s
    return q


def six_dof_to_mobile(T):
    """geometry::trans::SixDof2Mobile (trans.hpp:68-86): keep x, y and the rotation about +-z (if the rotation axis is
    within ~18 deg of z, else drop the rotation)."""
    q = rot_to_quat(T[:3, :3])
    n = np.linalg.norm(q[1:])
    if n > 0:  # Eigen::AngleAxis(Quaternion)
        angle = 2.0 * np.arctan2(n, abs(q[0]))
        if q[0] < 0:
            n = -n
        axis = q[1:] / n
    else:
        angle, axis = 0.0, np.array([1.0, 0.0, 0.0])
    res = np.eye(4)
    res[:2, 3] = T[:2, 3]
    tmp = axis[2]
    if abs(tmp) > 0.95:
        a = angle
        sz = np.copysign(1.0, tmp)
        c, s = np.cos(a), np.sin(a) * sz
        res[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    return res


class MapManager:
    """keyframes + submap (frontend/src/MapManager.cpp). The submap lives on the device as the register's target."""

    def __init__(self, ctx, grid_size=0.5, is_mapping=True):
        self.ctx = ctx
        self.grid_size = float(grid_size)
        self.is_mapping = is_mapping
        self.keyframes = []          # list of (cloud (n, 8) float32 C-contiguous, pose 4x4)
        self.closest_kf_idx = []
        self.submap_idx = []
        self.submap_size = 0
        self.cur_pose = np.eye(4)
        self.last_pose = np.eye(4)
        self.update_requested = False
        self.n_updates = 0

    def isSubmapEmpty(self):
        return len(self.submap_idx) == 0

    def notifyUpdateMap(self):
        self.update_requested = True

    def setCurPose(self, p):  # MapManager.cpp:109-119
        self.cur_pose = p.copy()
        if np.linalg.norm(self.last_pose[:3, 3] - p[:3, 3]) > MIN_KF_GAP:
            self.last_pose = p.copy()
            self.notifyUpdateMap()

    def putKeyFrame(self, cloud, pose):  # MapManager.cpp:122-149
        if not self.is_mapping:
            return False
        cloud = np.ascontiguousarray(cloud, dtype=np.float32)
        if not self.keyframes:
            self.keyframes.append((cloud, pose.copy()))
            return True
        pos = np.array([kf[1][:3, 3] for kf in self.keyframes])
        d2 = np.sum((pos - pose[:3, 3]) ** 2, axis=1)
        k = int(np.argmin(d2))
        if d2[k] > MIN_KF_GAP:  # the reference compares the SQUARED distance with minKFGap (:141)
            self.keyframes.append((cloud, pose.copy()))
            self.closest_kf_idx.append(k)
            return True
        return False

    def updateMap(self):  # MapManager.cpp:151-201
        self.update_requested = False
        if not self.keyframes:
            return
        pos = np.array([kf[1][:3, 3] for kf in self.keyframes])
        d2 = np.sum((pos - self.cur_pose[:3, 3]) ** 2, axis=1)
        idx = [int(i) for i in np.nonzero(d2 < SUBMAP_RADIUS * SUBMAP_RADIUS)[0]]  # nanoflann radius search: d2 < r^2
        self.submap_idx = idx
        _, self.submap_size = self.ctx.submap_build([self.keyframes[i][0] for i in idx], [self.keyframes[i][1] for i in idx], self.grid_size,
                                                    want_points=False, ids=idx)  # device copies cached by keyframe index (LRU-bounded)
        self.n_updates += 1


class LidarOdometry:
    """frontend::LidarOdometry (frontend/src/LidarOdometry.cpp) fed frame by frame."""

    def __init__(self, pcr_type="loam", grid_size=0.5, device=0, **params):
        method = {"loam": capi.PCR_LOAM, "ndt": capi.PCR_NDT, "vgicp": capi.PCR_VGICP}.get(pcr_type)
        if method is None:
            raise RuntimeError("such pcr type(%s) is not exist, please implemented your self!" % pcr_type)
        self.ctx = capi.Context(method, device=device, **params)
        self.grid_size = float(grid_size)
        self.map = MapManager(self.ctx, grid_size)
        self.last_pos = np.zeros(3)
        self.reloc_pose = np.eye(4)
        self.reloc = False
        self.global_odom = []        # list of (stamp, pose)
        self.odom2map = np.eye(4)
        self.odom2map_init = False
        self.converged = []

    def setRelocFlag(self, pose):
        self.reloc_pose = pose.copy()
        self.reloc = True

    def generateOdom(self, scan, stamp, local_odom=None):
        """one frame: raw scan (n, 8) float32, stamp (s), local odometry pose (4x4) or None. Returns the global pose."""
        init_pose = self.reloc_pose.copy()
        if self.reloc:
            self.reloc = False
            self.global_odom = []
        elif local_odom is not None and self.odom2map_init:
            init_pose = self.odom2map @ local_odom
        else:
            n = len(self.global_odom)
            if n:  # Frontend::getClosestItem: walk back from the newest while the stamp distance shrinks
                cidx, m = n - 1, abs(stamp - self.global_odom[-1][0])
                for i in range(n - 2, -1, -1):
                    t = abs(self.global_odom[i][0] - stamp)
                    if t < m:
                        m, cidx = t, i
                    else:
                        break
                if cidx > 0:
                    init_pose = self.global_odom[cidx][1].copy()
        conv = True
        if not self.map.isSubmapEmpty():
            # mVoxelGrid.filter (:170-171) + mPcr->scan2Map against the resident submap (:184) in one call: the downsampled
            # scan stays on the device (same result as voxel_downsample + align, tests/test_gpu_frontend.py)
            init_pose, conv, _ = self.ctx.downsample_align(scan, self.grid_size, init_pose)
        self.converged.append(bool(conv))
        init_pose = six_dof_to_mobile(init_pose)                       # :211
        self.map.setCurPose(init_pose)
        if self.map.isSubmapEmpty():
            self.map.putKeyFrame(scan, init_pose)
            self.map.notifyUpdateMap()
        else:  # selectKeyFrame (:81-87)
            if np.linalg.norm(init_pose[:3, 3] - self.last_pos) > MIN_KF_GAP:
                self.map.putKeyFrame(scan, init_pose)
                self.last_pos = init_pose[:3, 3].copy()
        self.global_odom.append((stamp, init_pose.copy()))
        if local_odom is not None:
            self.odom2map_init = True
            self.odom2map = init_pose @ np.linalg.inv(local_odom)
        if self.map.update_requested:   # the map thread, run synchronously (module docstring)
            self.map.updateMap()
        return init_pose

    def close(self):
        self.ctx.close()


class LoopClosureVerifier:
    """Verification half of backend::LoopClosureManager::lcHandler (LoopClosureManager.cpp:73-110): for a candidate pair
    (oldKey, curKey) build the history submap of oldKey (+-range keyframes, transformed, 0.5 m downsample), register the
    current keyframe's cloud with VGICP in loop-closure mode from its own pose, accept iff converged and fitness < thresh."""

    def __init__(self, keyframes, context_pc_ds=0.5, history_submap_range=1, fitness_score=0.3, device=0, workers=1, devices=None):
        """workers / devices: verify_batch spreads independent candidates over `workers` VGICP contexts (one CUDA stream
        each) placed round-robin on `devices` (default: [device]) — SURVEY §8f-2 'batch across candidates / GPUs'."""
        self.keyframes = keyframes
        self.ds = float(context_pc_ds)
        self.range = int(history_submap_range)
        self.thresh = float(fitness_score)
        devs = list(devices) if devices else [device]
        self.ctxs = []
        for w in range(max(1, int(workers))):
            c = capi.Context(capi.PCR_VGICP, device=devs[w % len(devs)])
            c.init_for_lc()
            self.ctxs.append(c)
        self.ctx = self.ctxs[0]

    def _verify_on(self, ctx, old_key, cur_key):
        n = len(self.keyframes)
        near = [k for k in range(old_key - self.range, old_key + self.range + 1) if 0 <= k < n]
        _, m = ctx.submap_build([self.keyframes[k][0] for k in near], [self.keyframes[k][1] for k in near], self.ds, want_points=False, ids=near)
        cloud, pose = self.keyframes[cur_key]
        T, conv = ctx.align(cloud, pose)
        fs = ctx.fitness()
        return dict(old=old_key, cur=cur_key, converged=bool(conv), fitness=fs, accepted=bool(conv and fs < self.thresh), T=T, map_points=m)

    def verify(self, old_key, cur_key):
        return self._verify_on(self.ctx, old_key, cur_key)

    def verify_batch(self, pairs):
        """independent candidates (old_key, cur_key) verified concurrently, one context per in-flight candidate; results in
        the order of `pairs`, identical to verifying them one after the other (no state is shared between candidates)"""
        pairs = list(pairs)
        if len(self.ctxs) == 1 or len(pairs) <= 1:
            return [self.verify(o, c) for o, c in pairs]
        import queue
        from concurrent.futures import ThreadPoolExecutor
        free = queue.Queue()
        for c in self.ctxs:
            free.put(c)

        def one(p):
            ctx = free.get()
            try:
                return self._verify_on(ctx, p[0], p[1])   # ctypes releases the GIL: the contexts' streams run concurrently
            finally:
                free.put(ctx)
        with ThreadPoolExecutor(max_workers=len(self.ctxs)) as ex:
            return list(ex.map(one, pairs))

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []


class ScanContext:
    """backend::context::ScanContext (backend/src/ScanContext.cpp): descriptors and descriptor distances on the GPU
    (pcr_scancontext_make / pcr_scancontext_distance), the candidate bookkeeping of query() (:232-290) on the host.
    Ring-key neighbours are taken in (d2, index) order; the ring-key set is refreshed with the reference's rule
    (rebuilt when more than numExcludeRecent + buildTreeGap keys are missing)."""

    def __init__(self, ctx, lidar_height=2.0, num_exclude_recent=40, build_tree_gap=10, num_candidates=10, search_ratio=0.1, dist_thres=0.4,
                 sector_key_align=False):
        self.ctx = ctx
        self.lidar_height = float(lidar_height)
        self.num_exclude_recent, self.build_tree_gap, self.num_candidates = int(num_exclude_recent), int(build_tree_gap), int(num_candidates)
        self.search_ratio, self.dist_thres, self.sector_key_align = float(search_ratio), float(dist_thres), bool(sector_key_align)
        self.polarcontexts, self.ringcontexts, self.sectorcontexts = [], [], []
        self.ring_sub = 0

    def size(self):
        return len(self.polarcontexts)

    def addContext(self, *clouds):
        """one or more (already downsampled) keyframe clouds -> descriptors, one kernel launch for all of them"""
        desc, rk, sk = self.ctx.scancontext_make(list(clouds), self.lidar_height)
        for k in range(len(clouds)):
            self.polarcontexts.append(desc[k]); self.ringcontexts.append(rk[k]); self.sectorcontexts.append(sk[k])

    def query(self, idx):
        """(matched keyframe or -1, yaw offset in rad) — ScanContext::query"""
        if idx <= self.num_exclude_recent + self.num_candidates:
            return -1, 0.0
        if self.ring_sub == 0 or idx - self.ring_sub > self.num_exclude_recent + self.build_tree_gap:
            self.ring_sub = idx - self.num_exclude_recent
        sub = np.asarray(self.ringcontexts[: self.ring_sub])
        d2 = np.sum((sub - self.ringcontexts[idx]) ** 2, axis=1)
        cand = np.lexsort((np.arange(len(sub)), d2))[: self.num_candidates]
        descs = np.stack([self.polarcontexts[idx]] + [self.polarcontexts[j] for j in cand])
        pairs = [(0, k + 1) for k in range(len(cand))]
        dist, shift = self.ctx.scancontext_distance(descs, pairs, self.search_ratio, self.sector_key_align)
        best, arg, nn = np.inf, 0, 0
        for k, j in enumerate(cand):
            if dist[k] < best:
                best, arg, nn = dist[k], int(shift[k]), int(j)
        if best > self.dist_thres:
            return -1, 0.0
        return nn, float(np.float32(np.float32(360.0 / 60.0) * arg) * np.pi / 180.0)


class LoopClosureManager:
    """backend::LoopClosureManager (backend/src/LoopClosureManager.cpp:28-119) without threads: addContext() turns every new
    keyframe into a ScanContext (0.5 m downsample first, :31-35), lcHandler() queries each new context and verifies the
    candidates with VGICP in loop-closure mode + the fitness gate. Returns the accepted loops as (oldKey, curKey,
    old_pose^-1 * cur_pose) — the reference stores the relative pose of the CURRENT keyframe poses, not the refined one
    (:106-108)."""

    def __init__(self, keyframes, context_pc_ds=0.5, history_submap_range=1, fitness_score=0.3, lidar_height=2.0, device=0, workers=1, devices=None,
                 **sc_params):
        self.keyframes = keyframes   # the MapManager's list of (cloud, pose): shared, grows as mapping proceeds
        self.ds = float(context_pc_ds)
        self.verifier = LoopClosureVerifier(keyframes, context_pc_ds, history_submap_range, fitness_score, device=device, workers=workers, devices=devices)
        self.ctb = ScanContext(self.verifier.ctx, lidar_height=lidar_height, **sc_params)
        self.n_contexts = 0
        self.lc_size = 0
        self.loops = []
        self.checked = []            # every verified candidate: dict of LoopClosureVerifier.verify

    def addContext(self):
        new = self.keyframes[self.n_contexts:]
        if new:
            self.ctb.addContext(*[self.verifier.ctx.voxel_downsample(kf[0], self.ds) for kf in new])
            self.n_contexts = len(self.keyframes)

    def lcHandler(self):
        # the reference verifies candidate after candidate (LoopClosureManager.cpp:73-110); the candidates of the new contexts
        # are independent of each other, so they are collected first and verified as one batch (verify_batch)
        cands = []
        for i in range(self.lc_size, self.ctb.size()):
            old, _yaw = self.ctb.query(i)
            if old >= 0:
                cands.append((old, i))
        for r in self.verifier.verify_batch(cands):
            self.checked.append(r)
            if r["accepted"]:
                self.loops.append((r["old"], r["cur"], np.linalg.inv(self.keyframes[r["old"]][1]) @ self.keyframes[r["cur"]][1]))
        self.lc_size = self.ctb.size()
        return self.loops

    def close(self):
        self.verifier.close()
