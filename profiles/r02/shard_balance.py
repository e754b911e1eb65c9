"""How evenly do contiguous shards of the 1024-scan job load 8 ranks? (round 2 diagnostics, one GPU): time of every 128-scan
shard, contiguous vs interleaved (scan k -> rank k mod 8)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from simpleslam_b200 import capi, workloads
ctx = capi.Context(capi.PCR_NDT)
ds = lambda p, l: ctx.voxel_downsample(p, l)
dst, _ = workloads.c4_map(ds, keep_raw=True)
ctx.set_target(dst)
scans, truths, guesses = workloads.c4_scans("ndt", ds, 0, 1024, 1024, workers=16)
def run(idx):
    sc = [scans[i] for i in idx]
    offs = np.concatenate([[0], np.cumsum([len(s) for s in sc])]).astype(np.uint64)
    dev = torch.from_numpy(np.ascontiguousarray(np.concatenate(sc))).cuda()
    g = [guesses[i] for i in idx]
    ctx.batch_align(None, offs, g, device_ptr=dev.data_ptr(), stride=32)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.batch_align(None, offs, g, device_ptr=dev.data_ptr(), stride=32)
        ts.append(1e3 * (time.perf_counter() - t0))
    return min(ts)
for name, shards in (("contiguous", [list(range(r * 128, (r + 1) * 128)) for r in range(8)]), ("interleaved", [list(range(r, 1024, 8)) for r in range(8)])):
    t = [run(s) for s in shards]
    print(name, " ".join("%.2f" % x for x in t), "| max %.2f mean %.2f -> balance %.3f" % (max(t), np.mean(t), np.mean(t) / max(t)))
