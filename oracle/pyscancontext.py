"""ORACLE — TEST INFRASTRUCTURE ONLY. numpy / pure-Python restatement of backend/src/ScanContext.cpp:
makeScanContext (:152-196), ring / sector keys (:198-230), fastAlignUsingVkey (:89-108), computeSimularity (:55-87),
distanceBtnScanContext (:116-150) and query (:232-290). PARITY UNPINNED (no fixtures in the reference).
Conventions: atan2f is taken as the correctly rounded float arctangent (double atan2 rounded to float); the ring-key
k-NN of query() breaks ties by (d2, index) (nanoflann's order is tree dependent)."""
import math
import numpy as np

RINGS, SECTORS, MAX_RADIUS = 20, 60, np.float32(80.0)


def make_scancontext(cloud, lidar_height=2.0):
    x = cloud[:, 0].astype(np.float32)
    y = cloud[:, 1].astype(np.float32)
    z = (cloud[:, 2].astype(np.float32) + np.float32(lidar_height)).astype(np.float32)
    rng = np.sqrt((x * x + y * y).astype(np.float32)).astype(np.float32)
    at = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)          # atan2f
    res = (at.astype(np.float64) + math.pi).astype(np.float32)                              # float res = atan2f + M_PI
    res = np.maximum(np.float32(0.0), np.minimum(np.float32(2 * math.pi), res))
    ang = (res.astype(np.float64) * 180.0 / math.pi).astype(np.float32)                     # rad2deg<float>
    desc = np.full((RINGS, SECTORS), -1000.0)
    keep = ~(rng > MAX_RADIUS)
    ring = np.maximum(np.minimum(RINGS, np.ceil((rng / MAX_RADIUS).astype(np.float32) * np.float32(RINGS)).astype(np.int64)), 1)
    sec = np.maximum(np.minimum(SECTORS, np.ceil((ang.astype(np.float64) / 360.0) * SECTORS).astype(np.int64)), 1)
    for r, s, zz in zip(ring[keep], sec[keep], z[keep]):
        if desc[r - 1, s - 1] < zz:
            desc[r - 1, s - 1] = float(zz)
    desc[desc == -1000.0] = 0.0
    return desc


def ring_key(desc):
    return desc.mean(axis=1)


def sector_key(desc):
    return desc.mean(axis=0)


def circshift(m, k):
    return np.roll(m, k, axis=-1)  # shifted.col((c + k) % n) = m.col(c)


def similarity(sc1, sc2):
    s, eff = 0.0, 0
    for c in range(sc1.shape[1]):
        a, b = sc1[:, c], sc2[:, c]
        na, nb = math.sqrt(float(np.dot(a, a))), math.sqrt(float(np.dot(b, b)))
        if na == 0 or nb == 0:
            continue
        s = s + float(np.dot(a, b)) / (na * nb)
        eff += 1
    return 1.0 - (s / eff if eff else float("nan"))


def distance(sc1, sc2, search_ratio=0.1, sector_key_align=False):
    """(min distance, argmin shift). sector_key_align=False is the reference: its sector key is a 60 x 1 matrix, so
    fastAlignUsingVkey's loop over cols() evaluates shift 0 only (:93, :122-124)."""
    align = 0
    if sector_key_align:
        k1, k2 = sector_key(sc1), sector_key(sc2)
        best = float("inf")
        for sft in range(SECTORS):
            k2s = circshift(k2, sft)
            q = 0.0
            for cc in range(SECTORS):  # sequential sum (Eigen's vectorised norm order is not pinned; ties between shifts are real:
                dd = float(k1[cc]) - float(k2s[cc])  # a rotation-symmetric scene gives 60 equal norms up to rounding)
                q += dd * dd
            d = math.sqrt(q)
            if d < best:
                best, align = d, sft
    radius = int(round(0.5 * search_ratio * SECTORS))
    space = [align]
    for ii in range(1, radius + 1):
        space.append((align + ii + SECTORS) % SECTORS)
        space.append((align - ii + SECTORS) % SECTORS)
    space.sort()
    best, arg = float("inf"), 0
    first = True
    for sft in space:
        d = similarity(sc1, circshift(sc2, sft))
        if d < best or (first and not (d == d) and False):
            best, arg = d, sft
        first = False
    if best == float("inf"):
        best = 1.7976931348623157e308  # std::numeric_limits<double>::max(): no candidate was smaller (all NaN)
    return best, arg


class OracleScanContext:
    def __init__(self, lidar_height=2.0, num_exclude_recent=40, build_tree_gap=10, num_candidates=10, search_ratio=0.1, dist_thres=0.4,
                 sector_key_align=False):
        self.h, self.excl, self.gap, self.k = lidar_height, num_exclude_recent, build_tree_gap, num_candidates
        self.ratio, self.thres, self.align = search_ratio, dist_thres, sector_key_align
        self.descs, self.rings = [], []
        self.ring_sub = 0

    def add(self, cloud):
        d = make_scancontext(cloud, self.h)
        self.descs.append(d)
        self.rings.append(ring_key(d))

    def query(self, i):
        if i <= self.excl + self.k:
            return -1, 0.0
        if self.ring_sub == 0 or i - self.ring_sub > self.excl + self.gap:
            self.ring_sub = i - self.excl
        sub = np.array(self.rings[: self.ring_sub])
        d2 = np.sum((sub - self.rings[i]) ** 2, axis=1)
        order = sorted(range(len(sub)), key=lambda j: (d2[j], j))[: self.k]
        best, arg, idx = float("inf"), 0, 0
        for j in order:
            dist, sft = distance(self.descs[i], self.descs[j], self.ratio, self.align)
            if dist < best:
                best, arg, idx = dist, sft, j
        if best > self.thres:
            return -1, 0.0
        return idx, float(np.float32(np.float32(360.0 / SECTORS) * arg) * math.pi / 180.0)
