import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from simpleslam_b200 import frontend, workloads, capi
import cProfile, pstats
seq = workloads.c5_sequence(150)
lo = frontend.LidarOdometry("loam")
for f in seq["frames"][:30]:
    lo.generateOdom(f["scan"], f["stamp"], f["local_odom"])
lo.close()
lo = frontend.LidarOdometry("loam")
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for f in seq["frames"]:
    lo.generateOdom(f["scan"], f["stamp"], f["local_odom"])
dt = time.perf_counter() - t0
pr.disable()
print("ms/frame", 1e3 * dt / len(seq["frames"]))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
