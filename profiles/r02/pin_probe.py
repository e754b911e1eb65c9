"""What cudaPointerGetAttributes says about the host buffers bench.py hands the C ABI (round 2 diagnostics)."""
import ctypes, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
rt = ctypes.CDLL("libcudart.so")
class Attr(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("device", ctypes.c_int), ("devicePointer", ctypes.c_void_p), ("hostPointer", ctypes.c_void_p)]
def attr(a):
    at = Attr()
    t0 = time.perf_counter()
    rc = rt.cudaPointerGetAttributes(ctypes.byref(at), ctypes.c_void_p(a.ctypes.data))
    return rc, at.type, 1e6 * (time.perf_counter() - t0)
torch.cuda.init(); torch.zeros(1, device="cuda")
for mb in (4, 64, 512):
    a = np.zeros((mb << 20) // 32 * 8, np.float32).reshape(-1, 8)
    p = torch.from_numpy(a).pin_memory().numpy()
    print(mb, "MB pageable", attr(a), "pinned", attr(p), attr(p))
from simpleslam_b200 import capi
import bench
ctx = capi.Context(capi.PCR_NDT, cores=int(os.environ.get("PCR_BENCH_CORES", "4")))
wl = bench.build_workload("c2_ndt", lambda p, l: ctx.voxel_downsample(p, l), 2, 0)
s, d, Tg, _ = bench.step_inputs(wl, 0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hs, hd = pin(s), pin(d)
for mode in ("auto", "0", "1"):
    if mode == "auto": os.environ.pop("PCR_HOST_PACK", None)
    else: os.environ["PCR_HOST_PACK"] = mode
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.scan2map(hs, hd, Tg); tp = time.perf_counter() - t0
        sp, dp = np.array(s, copy=True), np.array(d, copy=True)
        t0 = time.perf_counter(); ctx.scan2map(sp, dp, Tg); tg = time.perf_counter() - t0
        t0 = time.perf_counter(); ctx.set_target(hd); t1 = time.perf_counter() - t0
        t0 = time.perf_counter(); ctx.set_target(dp); t2 = time.perf_counter() - t0
        print("mode %s rep %d: scan2map pinned %.2f ms pageable %.2f ms | set_target pinned %.2f pageable %.2f" % (mode, rep, 1e3 * tp, 1e3 * tg, 1e3 * t1, 1e3 * t2))
