"""per-block wall time of the C5 loop split by call (host-side tuning aid)"""
import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from simpleslam_b200 import frontend, workloads, capi
seq = workloads.c5_sequence(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
lo = frontend.LidarOdometry("loam")
acc = {"ds": 0.0, "align": 0.0, "submap": 0.0}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[key] += time.perf_counter() - t0; return r
    setattr(obj, name, g)
wrap(lo.ctx, "voxel_downsample", "ds"); wrap(lo.ctx, "align", "align"); wrap(lo.ctx, "submap_build", "submap")
t0 = time.perf_counter(); last = dict(acc); lastn = 0
for k, f in enumerate(seq["frames"]):
    lo.generateOdom(f["scan"], f["stamp"], f["local_odom"])
    if (k + 1) % 30 == 0:
        t1 = time.perf_counter()
        print("frames %3d-%3d  %.2f ms/frame  ds %.2f align %.2f submap %.2f (per frame)  kfs %d updates %d submap_pts %d iters %d" % (
            k - 29, k, 1e3 * (t1 - t0) / 30, *(1e3 * (acc[q] - last[q]) / 30 for q in ("ds", "align", "submap")), len(lo.map.keyframes), lo.map.n_updates - lastn,
            lo.map.submap_size, lo.ctx.stats()["iterations"]))
        t0 = t1; last = dict(acc); lastn = lo.map.n_updates
