"""Job end to end from PINNED host memory: plain DMA of the 32-byte records vs packing them on the host with `cores`
threads first (round 2 diagnostics). python profiles/r02/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from simpleslam_b200 import capi, workloads
c0 = capi.Context(capi.PCR_NDT)
ds = lambda p, l: c0.voxel_downsample(p, l)
dst, _ = workloads.c4_map(ds, keep_raw=True)
scans, truths, guesses = workloads.c4_scans("ndt", ds, 0, 1024, 1024, workers=16)
offs = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint64)
host = torch.from_numpy(np.ascontiguousarray(np.concatenate(scans))).pin_memory().numpy()
print("cpus", os.cpu_count(), "bytes %.2f GB" % (host.nbytes / 1e9))
for cores, mode in ((4, "0"), (4, "1"), (8, "1"), (12, "1"), (16, "1")):
    os.environ["PCR_HOST_PACK"] = mode
    ctx = capi.Context(capi.PCR_NDT, cores=cores)
    ctx.set_target(dst)
    ctx.batch_align(host, offs, guesses)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.batch_align(host, offs, guesses)
        ts.append(time.perf_counter() - t0)
    print("cores %2d host_pack %s: %.1f ms -> %.0f registrations/s" % (cores, mode, 1e3 * min(ts), 1024 / min(ts)))
    ctx.close()
