"""Single-scan latency probe (round 2): one resident scan registered against a static map, stats per call.
usage: python profiles/r02/lat_probe.py ndt|loam|vgicp [c2|c4]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from simpleslam_b200 import capi, workloads
import bench
method = sys.argv[1] if len(sys.argv) > 1 else "ndt"
which = sys.argv[2] if len(sys.argv) > 2 else "c2"
mid = dict(ndt=capi.PCR_NDT, loam=capi.PCR_LOAM, vgicp=capi.PCR_VGICP)[method]
ctx = capi.Context(mid, device=0)
wl = bench.build_workload({"ndt": "c2_ndt", "loam": "c1_loam", "vgicp": "c3_vgicp"}[method] if which == "c2" else "c4_" + method, lambda p, l: ctx.voxel_downsample(p, l), 2, 0)
src, dst, Tg, Tt = bench.step_inputs(wl, 0)
ctx.set_target(dst)
d = torch.from_numpy(np.ascontiguousarray(src)).cuda()
n, stride = src.shape[0], src.shape[1] * 4
for it in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    T, conv = ctx.align_device(d.data_ptr(), n, stride, Tg)
    w = 1e3 * (time.perf_counter() - t0)
    st = ctx.stats()
    print("n %d wall %.3f ms total %.3f kernel %.3f launches %d iters %s evals %s conv %s" % (n, w, st["ms_total"], st["ms_hot_kernel"], st["kernel_launches"], st["iterations"], (st["evaluations"], st["hessian_evals"]), conv))
