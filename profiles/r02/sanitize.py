"""Small workload for compute-sanitizer (round 2): every kernel family once, incl. a multi-scan NDT batch (two-level
reduction with several groups), the split LOAM kernels and the VGICP device-side LM.
  compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/r02/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import data
from simpleslam_b200 import capi
os.environ["PCR_LOAM_LPQ"] = "1"; os.environ["PCR_LOAM_TILE"] = "32"   # force the split search / fit kernels on a small batch
case = data.loam_case()
c = capi.Context(capi.PCR_LOAM); c.set_target(case["dst"])
scans = [case["src"][::2], case["src"][1::2], case["src"][::3]]
offs = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint64)
print("loam", c.batch_align(np.concatenate(scans), offs, [case["T_guess"]] * 3)[1], c.loam_last_shape())
print("ds", c.voxel_downsample(case["dst"], 0.5).shape)
c.close()
case = data.ndt_case()
c = capi.Context(capi.PCR_NDT); c.set_target(case["dst"])
print("ndt single", c.align(case["src"], case["T_guess"])[1])    # one scan owns the wave: groups of 32 blocks
scans = [case["src"][::2], case["src"][1::2], case["src"][::5]]
offs = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint64)
print("ndt batch", c.batch_align(np.concatenate(scans), offs, [case["T_guess"]] * 3)[1])
c.close()
case = data.vgicp_case()
c = capi.Context(capi.PCR_VGICP); c.set_target(case["dst"][::4])
print("vgicp", c.align(case["src"][::4], case["T_guess"])[1], c.fitness())
c.close()
