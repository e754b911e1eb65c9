#!/usr/bin/env python
"""bench.py — scan registrations/sec of the PCR hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4_job_ndt|c4_job_loam|c2_ndt|...]

Default (the driver's line): BASELINE config 4 — a FIXED job of 1024 independent 64-beam scans localised with NDT against
a ~20 M-point static map (loc.cpp mode), sharded in contiguous blocks over the N ranks (strong scaling): rank 0 builds the
index once and broadcasts it over NCCL before the timed region; one "step" = the whole 1024-scan job = every rank
registers its shard (device-resident scans) and the poses are all-gathered. `value` = K * 1024 / device time (CUDA events
around the K steps, max over ranks).
  e2e        the same job through the reference-facing call pcr_batch_align with HOST buffers (pinned; pageable reported
             next to it): scan upload + align + pose read-back + gather inside the timed region.
  roofline   dominant kernel: algorithmic bytes (SURVEY.md §8(d) formulas) / its device time measured live.
  cpu_baseline / parity   the CPU oracle on a bounded sample of the same job; K poses of the timed job compared with it.
  workloads  (N = 1 only) the other BASELINE configs measured the same way: C1 LOAM, C2 NDT, C3 VGICP (LM and GN-20), C4 LOAM
             job, C5 offline LIO loop at 2000 frames — value / p50 / e2e / roofline / cpu_baseline / parity each.
--impl reference times the CPU oracle on the headline config (the reference's own libPCR needs PCL/Eigen/FLANN: not buildable here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# SURVEY.md §8(d) algorithmic bytes (SoA/float4 map, minimal traffic, no cache credit)
def algo_bytes_8d(method, point_evals, pairs, c_bar=None):
    """Algorithmic bytes by SURVEY.md §8(d)'s per-unit figures (SoA / float4 map, minimal traffic, no cache credit) — the
    numerator of roofline.achieved. point_evals = source points pushed through the kernel, summed over its launches.
      LOAM iteration   Ns * (16 + 27*8 + 16*C_bar), C_bar = mean population of the 27 one-metre cells around a query (oracle)
      NDT evaluation   Ns * (16 + 7*8 + 104*V_bar), V_bar = (point, leaf) pairs per point (counted by the kernel)
      VGICP evaluation Ns * (16 + 48 + 8) + 84 * correspondences"""
    if method == "ndt":
        return point_evals * (16.0 + 56.0) + 104.0 * pairs
    if method == "loam":
        return None if c_bar is None else point_evals * (16.0 + 27 * 8.0 + 16.0 * c_bar)
    if method == "vgicp":
        return point_evals * (16.0 + 48.0 + 8.0) + 84.0 * pairs
    raise ValueError(method)


def algo_bytes_examined(method, point_evals, index_reads, pairs):
    """Bytes THIS implementation's algorithm has to read once (its own record sizes, its pruned search), no cache credit:
    index_reads = spatial-index entries read (LOAM: x-rows looked up, 2 x 4 B each; NDT / VGICP: 4-byte table entries);
    pairs = candidate map points examined (LOAM, 16 B) / (point, leaf) pairs (NDT, 64-byte leaf record) / correspondences
    (VGICP, 80-byte voxel record). Always <= the §8(d) figure; reported next to it as roofline.achieved_examined."""
    if method == "ndt":
        return point_evals * 16.0 + index_reads * 4.0 + 64.0 * pairs
    if method == "loam":
        return point_evals * 16.0 + index_reads * 8.0 + 16.0 * pairs
    if method == "vgicp":
        return point_evals * (16.0 + 48.0) + index_reads * 4.0 + 80.0 * pairs
    raise ValueError(method)


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line), every 200 ms, through
    NVML in a background thread (same counters as `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`;
    an in-process NVML query disturbs the measured CUDA calls less than an nvidia-smi poller). Falls back to nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, enabled=True):
        self.rows = []      # (sm_mhz, max_mhz, reasons bitmask)
        self.proc = None
        self.mode = None
        self._stop = threading.Event()
        if not enabled or os.environ.get("PCR_BENCH_SAMPLER") == "off":   # only rank 0 reports clocks (extra pollers contend for the driver lock)
            return
        if os.environ.get("PCR_BENCH_SAMPLER") == "smi":
            self._start_smi(index)
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.mode = None
        self._start_smi(index)

    def _start_smi(self, index):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((sm, self.mx, rs))
            except Exception:
                pass
            self._stop.wait(0.2)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            rs = 0
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    rs |= self.BITS[n]
            self.rows.append((sm, mx, rs))

    def wait_samples(self, n=1, timeout=6.0):
        """the first sample takes a moment: block until n rows are in (or the timeout passes)"""
        if not self.mode:
            return
        t0 = time.perf_counter()
        while len(self.rows) < n and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if not self.mode:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable (not rank 0, or no NVML / nvidia-smi)"], "samples": 0}
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[2]
        reasons = sorted(n for n, b in self.BITS.items() if bits & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": self.mode}


def build_workload(name, downsample, n_scans, seed_offset):
    from simpleslam_b200 import workloads
    if name == "c2_ndt":
        return workloads.c2_ndt(downsample, n_scans, seed_offset=seed_offset)
    if name == "c1_loam":
        return workloads.c1_loam(downsample, n_scans, seed_offset=seed_offset)
    if name in ("c3_vgicp", "c3_vgicp_gn"):
        wl = workloads.c3_vgicp(n_scans, seed_offset=seed_offset)
        if name.endswith("_gn"):  # BASELINE config wording: "20 Gauss-Newton iterations" (the reference's shipped default is LM, <= 64)
            wl["name"] += ", Gauss-Newton capped at 20 iterations"
            wl["vgicp"] = dict(optimizer="GN", max_iterations=20)
        return wl
    if name in ("c4_loam", "c4_ndt"):
        return workloads.c4_batched(name[3:], downsample, n_scans, seed_offset=seed_offset)
    raise SystemExit("unknown workload " + name)


VGICP_ORACLE_OPTS = {}   # set by the c3_vgicp_gn workload


def oracle_register(method, src, dst, T, threads):
    """one full scan2Map with the CPU oracle, reference structure: index rebuilt on every call"""
    from oracle import pyoracle as orc
    if method == "ndt":
        return orc.Ndt(dst, 1.0).align(src, T, threads=threads)["T"]
    if method == "loam":
        return orc.loam_align(src, dst, T, threads=threads)["T"]
    return orc.Vgicp(dst, 1.0, 20, threads=threads).align(src, T, threads=threads, **VGICP_ORACLE_OPTS)["T"]


def step_inputs(wl, k):
    if wl["method"] == "vgicp":
        p = wl["pairs"][k % len(wl["pairs"])]
        return p["src"], p["dst"], p["T_guess"], p["T_true"]
    n = len(wl["scans"])
    return wl["scans"][k % n], wl["dst"], wl["guesses"][k % n], wl["truths"][k % n]


def pose_err(Ta, Tb):
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    f = np.linalg.norm(Ta[:3, :3] - Tb[:3, :3])
    return dt, float(2.0 * np.arcsin(min(1.0, f / (2.0 * np.sqrt(2.0)))))


def host_threads(world):
    """`cores` of the registers bench.py creates: the host threads that pack uploads (hostpack.hpp). The reference's shipped
    value is 4. One rank: like the reference arm, which runs on every host core, the GPU arm may use the cores of the box —
    three quarters of them, at most 12 (from 8 up pinned batches are packed too: job e2e 12.0 k -> 13.5 k). Several ranks on
    one host: the shipped 4 — the host's memory bandwidth is the bound there and packing pinned batches costs it twice the
    traffic of a plain DMA (measured at N = 2: 22.4 k plain, 17.5 k packed). PCR_BENCH_CORES overrides."""
    env = os.environ.get("PCR_BENCH_CORES")
    if env:
        return max(0, int(env))
    if world > 1:
        return 4
    return int(max(4, min(12, (3 * (os.cpu_count() or 4)) // 4)))


HOT_KERNEL = {"ndt": "ndt_round_kernel", "loam": "loam_search_kernel + loam_iter_kernel<fit> (one Gauss-Newton iteration)", "vgicp": "vgicp_eval_kernel"}


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s of B200_PROFILING.md (of fallback)"


def load_traffic(key, kernel_prefix):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same workload"""
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        for k in (key if isinstance(key, (list, tuple)) else [key]):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", fn))).get(k)
                if tr and tr["kernel"].startswith(kernel_prefix):
                    return tr["dram_bytes_per_launch"], "profiles/%s[%s]: %s (ncu --set full, one launch)" % (fn, k, tr["source"])
            except Exception:
                pass
    return None, None


def make_roofline(method, prof, c_bar, traffic_key):
    """prof: dict(hot_ms, hot_launches, pairs, pt_evals, idx_reads, ms_total) accumulated over the profiled pass"""
    peak, peak_src = load_peak()
    hot_ms = prof["hot_ms"]
    ab = algo_bytes_8d(method, prof["pt_evals"], prof["pairs"], c_bar)
    abx = algo_bytes_examined(method, prof["pt_evals"], prof["idx_reads"], prof["pairs"])
    formula = "SURVEY §8(d)"
    if ab is None:  # LOAM without the oracle's C_bar (--no-cpu-baseline): fall back to the bytes really examined
        ab, formula = abx, "examined bytes (C_bar not measured: --no-cpu-baseline)"
    achieved = ab / (hot_ms * 1e-3) / 1e9 if hot_ms > 0 else None
    achieved_x = abx / (hot_ms * 1e-3) / 1e9 if hot_ms > 0 else None
    traffic, traffic_src = load_traffic(traffic_key, HOT_KERNEL[method].split(" ")[0].split("<")[0])
    nl = max(prof["hot_launches"], 1)
    return {"bound": "hbm", "kernel": HOT_KERNEL[method], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src, "formula": formula, "c_bar": c_bar,
            "achieved_examined": achieved_x, "frac_examined": (achieved_x / peak) if achieved_x else None,
            "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None, "peak_source": peak_src,
            "launches": prof["hot_launches"], "avg_launch_us": 1e3 * hot_ms / nl, "algorithmic_bytes_per_launch": ab / nl,
            "examined_bytes_per_launch": abx / nl, "points_per_launch": prof["pt_evals"] / nl,
            "pairs_per_point": prof["pairs"] / max(prof["pt_evals"], 1), "index_reads_per_point": prof["idx_reads"] / max(prof["pt_evals"], 1),
            "kernel_share_of_step": hot_ms / prof["ms_total"] if prof.get("ms_total") else None,
            "measured_in": "a second pass over the same steps with CUDA events on the library stream around the kernel sequence",
            "note": "achieved = §8(d) algorithmic bytes (no cache credit) / measured kernel time; achieved_examined = the bytes this pruned / "
                    "compact-record implementation really needs. DRAM traffic is far below both: the index is L2-friendly and the kernels are "
                    "latency / issue bound, not HBM bound (DESIGN.md §4, profiles/)"}


def measure_index_build(ctx, dev, dst_host, raw_host=None, leaf=0.2, method="loam"):
    """the map-side kernels (the rows where an HBM fraction is meaningful): voxel downsample of the raw map
    (104 N + 16 M bytes) and the index build over the downsampled map (104 Nm bytes, + 104 B per NDT leaf), device-resident
    input, warm allocations, best of 3 host-blocking calls."""
    import torch
    peak, peak_src = load_peak()
    out = {"peak": peak, "unit": "GB/s", "peak_source": peak_src}
    d = torch.from_numpy(np.ascontiguousarray(dst_host)).to(dev)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        ctx.set_target_device(d.data_ptr(), d.shape[0], 32)
        ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    nbytes = 104.0 * d.shape[0]
    if method == "ndt":
        try:
            nbytes += 104.0 * len(ctx.ndt_leaves()["keys"])
        except Exception:
            pass
    out["index_build"] = {"points": int(d.shape[0]), "ms": 1e3 * t, "algorithmic_bytes": nbytes, "achieved": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / peak,
                          "what": "pcr_set_target_device: pack + bounding box + keys + radix sort + cell-sorted copy + dense table%s" % (" + leaf statistics" if method == "ndt" else "")}
    if raw_host is not None:
        r = torch.from_numpy(np.ascontiguousarray(raw_host)).to(dev)
        o = torch.empty((r.shape[0], 8), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        ts, m = [], 0
        for _ in range(4):
            t0 = time.perf_counter()
            m = ctx.voxel_downsample_device(r.data_ptr(), r.shape[0], 32, leaf, o.data_ptr(), r.shape[0])
            ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        nb = 104.0 * r.shape[0] + 16.0 * m
        out["voxel_downsample"] = {"points_in": int(r.shape[0]), "points_out": int(m), "leaf": leaf, "ms": 1e3 * t, "algorithmic_bytes": nb,
                                   "achieved": nb / t / 1e9, "frac": nb / t / 1e9 / peak,
                                   "what": "pcr_voxel_downsample_device: pack + bounding box + PCL keys + radix sort + gather + per-voxel float centroids"}
        del r, o
    del d
    torch.cuda.empty_cache()
    return out


def add_prof(prof, st, ms_total=None):
    prof["hot_ms"] += st["ms_hot_kernel"]
    prof["hot_launches"] += st["hot_kernel_launches"]
    prof["pairs"] += st["n_pairs"]
    prof["pt_evals"] += st["n_point_evals"]
    prof["idx_reads"] += st["n_index_reads"]
    prof["ms_total"] += st["ms_total"] if ms_total is None else ms_total
    for k in ("ms_aux_kernel", "aux_kernel_launches", "n_aux_items"):
        prof[k] = prof.get(k, 0) + st.get(k, 0)


def new_prof():
    return dict(hot_ms=0.0, hot_launches=0, pairs=0, pt_evals=0, idx_reads=0, ms_total=0.0)


def parity_vs_oracle(gpu_T, gpu_conv, oracle_results):
    """K poses of a timed batch against the oracle's poses of the same registrations (north_star: 1e-4 m / 1e-4 rad)"""
    dts, drs, flags = [], [], 0
    for T, cv, o in zip(gpu_T, gpu_conv, oracle_results):
        dt, dr = pose_err(np.asarray(T), o["T"])
        dts.append(dt); drs.append(dr)
        flags += int(bool(cv) == bool(o["converged"]))
    return {"checked": len(dts), "max_dt": float(max(dts)) if dts else None, "max_dr": float(max(drs)) if drs else None,
            "converged_flags_equal": flags, "tolerance": "1e-4 m / 1e-4 rad",
            "ok": bool(dts and max(dts) < 1e-4 and max(drs) < 1e-4 and flags == len(dts))}


def oracle_full(method, src, dst, T, threads):
    from oracle import pyoracle as orc
    if method == "ndt":
        return orc.Ndt(dst, 1.0).align(src, T, threads=threads)
    if method == "loam":
        return orc.loam_align(src, dst, T, threads=threads)
    return orc.Vgicp(dst, 1.0, 20, threads=threads).align(src, T, threads=threads, **VGICP_ORACLE_OPTS)


# ======================================================================================================================
# --impl reference
# ======================================================================================================================
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle restatement; libPCR itself needs PCL/Eigen/FLANN) on the host cores,
    on the headline config (or --workload): every step = a bounded sample of that workload."""
    if rank != 0:
        return
    from oracle import pyoracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    note = "reference libPCR needs PCL + Eigen + FLANN (+ ROS for the frontend), absent and no network: this arm times oracle/ (CPU restatement, OpenMP, all host cores)"
    if args.workload == "c5_lio":  # the oracle's frontend loop over a bounded prefix of the same sequence
        from oracle import pyfrontend as opf
        from simpleslam_b200 import workloads
        m = min(args.frames, 120)
        seq = workloads.c5_sequence(m)
        oo = opf.OracleOdometry(args.pcr, threads=cores)
        t0 = time.perf_counter()
        for f in seq["frames"]:
            oo.step(f["scan"], f["stamp"], f["local_odom"])
        tc = time.perf_counter() - t0
        val = m / tc
        print(json.dumps({
            "impl": "reference", "metric": "offline LIO mapping frames/sec (%s frontend)" % args.pcr.upper(), "value": val, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": m, "warmup": 0, "ms_per_step": 1e3 * tc / m, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic", "config": {"workload": seq["name"], "pcr": args.pcr},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": "first %d frames through oracle/pyfrontend.py" % m},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "note": note}))
        return
    ds = lambda pts, leaf: orc.voxel_downsample(pts, leaf)["points"]  # noqa: E731
    n_steps_all = args.steps + args.warmup
    if args.workload.startswith("c4_job"):
        from simpleslam_b200 import workloads
        method = args.workload.split("_")[-1]
        t0 = time.perf_counter()
        dst, n_raw = workloads.c4_map(ds)
        n_sample = min(n_steps_all, 8)
        scans, truths, guesses = workloads.c4_scans(method, ds, 0, n_sample, args.job_scans, workers=cores)
        wl = dict(name=c4_job_name(method, len(dst), args.job_scans), method=method, dst=dst, scans=scans, truths=truths, guesses=guesses)
        t_gen = time.perf_counter() - t0
        sample = ("every step = ONE scan2Map of the %d-scan job (scan k mod %d), index rebuilt per call as the reference's register does "
                  "(NdtRegister / LoamRegister::scan2Map, loc.cpp:47-63); value = registrations/s of that sample" % (args.job_scans, n_sample))
        scaling = "strong"
    else:
        wl = build_workload(args.workload, ds, min(64, n_steps_all), 0)
        method = wl["method"]
        t_gen = None
        sample = "every step = one full scan2Map (index rebuilt per call, as the reference does) on the full workload"
        scaling = "weak"
    VGICP_ORACLE_OPTS.update(wl.get("vgicp", {}))
    for k in range(args.warmup):
        s, d, Tg, _ = step_inputs(wl, k)
        oracle_register(method, s, d, Tg, cores)
    times = []
    for k in range(args.warmup, n_steps_all):
        s, d, Tg, _ = step_inputs(wl, k)
        t0 = time.perf_counter()
        oracle_register(method, s, d, Tg, cores)
        times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    val = args.steps / total
    s0, d0, Tg0, _ = step_inputs(wl, 0)
    t0 = time.perf_counter()
    oracle_register(method, s0, d0, Tg0, 4)   # the reference's shipped setting is cores = 4 (config/params.json:5)
    t4 = time.perf_counter() - t0
    extra = {}
    if method == "ndt":  # the same alignment with the voxel grid already built: what a static-map caller would pay on the CPU
        o = orc.Ndt(d0, 1.0)
        t0 = time.perf_counter()
        o.align(s0, Tg0, threads=cores)
        extra["align_only_ms_index_prebuilt"] = 1e3 * (time.perf_counter() - t0)
    out = {
        "impl": "reference", "metric": "scan registrations/sec (%s)" % method.upper(), "value": val, "unit": "registrations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "p50_align_ms": 1e3 * float(np.median(times)),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": wl["name"], "n_source": int(len(s0)), "n_target": int(len(d0))},
        "cpu_baseline": dict({"value": val, "unit": "registrations/s", "cores": cores, "kind": "port", "sample": sample,
                              "at_4_threads": {"ms_per_registration": 1e3 * t4, "cores": 4,
                                               "note": "the reference's shipped `cores`; the NDT arm is dominated by the serial voxel-grid build, so more threads do not help"}}, **extra),
        "e2e": {"value": val, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "note": note, "setup": {"data_generation_s": t_gen},
    }
    print(json.dumps(out))


def c4_job_name(method, n_map, n_job):
    return ("C4 batched localisation job (%s, loc.cpp mode): %d independent %s scans vs a %.1fM-pt static map, sharded over the ranks"
            % (method.upper(), n_job, "VLP-16 (0.5 m downsample)" if method == "loam" else "64-beam", n_map / 1e6))


# ======================================================================================================================
# C5: offline LIO loop
# ======================================================================================================================
def measure_c5(args, rank, world, local_rank, dist, dev, n_frames, cpu=True):
    """C5 offline LIO mapping: the headless frontend loop (voxel downsample -> scan2Map against the
    device-resident submap -> keyframes -> submap assembly on the device) over a synthetic figure-8 sequence.
    frames/s is wall clock over the whole loop with HOST scans (every copy inside), so value == e2e; N > 1 runs N
    independent replicas of the sequence (different noise seeds): the loop itself is sequential."""
    import torch
    from simpleslam_b200 import frontend, workloads, multigpu
    t_gen = time.perf_counter()
    seq = workloads.c5_sequence(n_frames, seed_offset=rank)
    t_gen = time.perf_counter() - t_gen
    frames = seq["frames"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    scans = [pin(f["scan"]) for f in frames]

    def run_once(profile):
        lo = frontend.LidarOdometry(args.pcr, device=local_rank)
        lo.ctx.set_profiling(profile)
        poses, prof, launches = [], new_prof(), 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f, sc in zip(frames, scans):
            had_map = not lo.map.isSubmapEmpty()
            poses.append(lo.generateOdom(sc, f["stamp"], f["local_odom"]))
            if profile and had_map:
                st = lo.ctx.stats()
                add_prof(prof, st)
                launches += st["kernel_launches"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        info = dict(kfs=len(lo.map.keyframes), updates=lo.map.n_updates, submap=lo.map.submap_size, converged=int(np.sum(lo.converged)))
        lo.close()
        return dt, poses, prof, launches, info

    run_once(False)   # warm-up: one untimed pass over the whole sequence (allocations, module load, pinned pages)
    if dist is not None:
        dist.barrier()
    dt, poses, _, _, info = run_once(False)            # timed: profiling off
    _, _, prof, launches, _ = run_once(True)           # second pass with per-kernel events for the roofline leg
    if dist is not None:
        dt = multigpu.max_over_ranks(dist, [dt], dev)[0]
    value = world * n_frames / dt
    if rank != 0:
        return None
    peak, peak_src = load_peak()
    truth = np.array([f["truth"][:2, 3] for f in frames])
    est = np.array([P[:2, 3] for P in poses])
    ape = float(np.sqrt(np.mean(np.sum((truth - est) ** 2, axis=1))))
    cpub, dev_vs_oracle, par = None, None, None
    if cpu:
        from oracle import pyfrontend as opf
        from oracle import pyoracle as orc
        orc.build()
        cores = os.cpu_count() or 1
        m = min(n_frames, 60)
        oo = opf.OracleOdometry(args.pcr, threads=cores)
        t0 = time.perf_counter()
        op = [oo.step(f["scan"], f["stamp"], f["local_odom"]) for f in frames[:m]]
        tc = time.perf_counter() - t0
        dev_vs_oracle = float(max(np.linalg.norm(a[:3, 3] - b[:3, 3]) for a, b in zip(poses[:m], op)))
        errs = [pose_err(a, b) for a, b in zip(poses[:m], op)]
        par = {"checked": m, "max_dt": float(max(e[0] for e in errs)), "max_dr": float(max(e[1] for e in errs)), "tolerance": "1e-4 m / 1e-4 rad",
               "ok": bool(max(e[0] for e in errs) < 1e-4 and max(e[1] for e in errs) < 1e-4), "what": "poses of the first %d frames against the oracle's frontend loop" % m}
        cpub = {"value": m / tc, "unit": "frames/s", "cores": cores, "kind": "port",
                "sample": "first %d frames of the same sequence through oracle/pyfrontend.py (CPU registers, OpenMP %d threads)" % (m, cores)}
    method = args.pcr
    abx = algo_bytes_examined(method, prof["pt_evals"], prof["idx_reads"], prof["pairs"])
    achieved = abx / (prof["hot_ms"] * 1e-3) / 1e9 if prof["hot_ms"] > 0 else None
    return {
        "metric": "offline LIO mapping frames/sec (%s frontend)" % method.upper(), "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": n_frames, "warmup": args.warmup, "ms_per_step": 1e3 * dt / n_frames, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if method != "ndt" else "f32 pair math, f64 accumulate", "data": "synthetic",
        "config": {"workload": seq["name"], "pcr": method, "step": "one frame of LidarOdometry::generateOdom (downsample, scan2Map against the "
                   "device-resident submap, keyframe gating, submap rebuild when the pose moved 1 m)", "l2": "not flushed: every frame brings a new scan from the host",
                   "timing": "wall clock over the whole loop (host-driven per frame); value == e2e", "parallelism": "%d independent replica(s)" % world},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": int(np.mean([s.nbytes for s in scans])), "d2h_bytes_per_step": 16 * 8 + 4,
                "what": "frontend.LidarOdometry.generateOdom(scan, stamp, local_odom) with pinned host scans"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": HOT_KERNEL[method] if method != "loam" else "loam_iter_kernel (fused, single scan)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": None,
                     "formula": "examined bytes (single-scan launches: latency bound, see DESIGN.md §4)", "launches": prof["hot_launches"],
                     "avg_launch_us": 1e3 * prof["hot_ms"] / max(prof["hot_launches"], 1), "kernel_share_of_step": prof["hot_ms"] / (1e3 * dt), "peak_source": peak_src},
        "cpu_baseline": cpub, "parity": par,
        "trajectory": {"ape_rmse_m_vs_truth": ape, "max_dev_m_vs_oracle_prefix": dev_vs_oracle, "path_length_m": float(np.sum(np.linalg.norm(np.diff(truth, axis=0), axis=1))),
                       **info},
        "setup": {"data_generation_s": t_gen},
    }


# ======================================================================================================================
# C1 / C2 / C3: one map, batches of independent scans (replicas at N > 1)
# ======================================================================================================================
def measure_replicas(args, name, rank, world, local_rank, dist, dev, cpu=True):
    import torch
    from simpleslam_b200 import capi, multigpu
    method_id = {"c2_ndt": capi.PCR_NDT, "c1_loam": capi.PCR_LOAM, "c3_vgicp": capi.PCR_VGICP, "c3_vgicp_gn": capi.PCR_VGICP, "c4_loam": capi.PCR_LOAM,
                 "c4_ndt": capi.PCR_NDT}[name]
    static_map = name.startswith("c4")  # loc.cpp mode: the map is registered once, e2e = batches of host scans
    gn = name == "c3_vgicp_gn"
    n_host_threads = host_threads(world)
    ctx = capi.Context(method_id, device=local_rank, cores=n_host_threads, **(dict(vgicp_optimizer=capi.PCR_LSQ_GN, vgicp_max_iters=20) if gn else {}))
    ds = lambda pts, leaf: ctx.voxel_downsample(pts, leaf)  # noqa: E731
    B = args.batch if args.batch > 0 else {"c2_ndt": 8, "c1_loam": 64, "c3_vgicp": 1, "c3_vgicp_gn": 1, "c4_loam": 128, "c4_ndt": 16}[name]
    n_steps_all = args.steps + args.warmup
    n_unique = min(64, n_steps_all * B)
    t_gen = time.perf_counter()
    wl = build_workload(name, ds, n_unique, rank)
    t_gen = time.perf_counter() - t_gen
    method = wl["method"]
    if method == "vgicp":
        B = 1
    VGICP_ORACLE_OPTS.clear()
    VGICP_ORACLE_OPTS.update(wl.get("vgicp", {}))

    # ---- static map: rank 0 builds the index, broadcasts it once over NCCL; the others import it
    setup = {}
    if method != "vgicp":
        t0 = time.perf_counter()
        if rank == 0 or dist is None:
            ctx.set_target(wl["dst"])
        setup["index_build_ms"] = 1e3 * (time.perf_counter() - t0)
        if dist is not None:
            nbytes, dt = multigpu.broadcast_target(ctx, dist, rank, dev)
            setup["index_broadcast_ms"] = 1e3 * dt
            setup["index_bytes"] = nbytes

    # device-resident copies of the unique scans; a step's batch = B of them concatenated on the device
    uniq_dev = [torch.from_numpy(np.ascontiguousarray(step_inputs(wl, u)[0])).to(dev) for u in range(n_unique)]
    dev_dst = [torch.from_numpy(np.ascontiguousarray(p["dst"])).to(dev) for p in wl["pairs"]] if method == "vgicp" else None
    batches = []
    for k in range(n_steps_all):
        ids = [(k * B + b) % n_unique for b in range(B)]
        cat = uniq_dev[ids[0]] if B == 1 else torch.cat([uniq_dev[i] for i in ids])
        offs = np.concatenate([[0], np.cumsum([uniq_dev[i].shape[0] for i in ids])]).astype(np.uint64)
        batches.append((ids, cat, offs))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    def resident_step(k):
        ids, cat, offs = batches[k]
        Tg = [step_inputs(wl, i)[2] for i in ids]
        Tt = [step_inputs(wl, i)[3] for i in ids]
        if method == "vgicp":
            dd = dev_dst[ids[0] % len(dev_dst)]
            ctx.set_target_device(dd.data_ptr(), dd.shape[0], 32)  # VGICP target = the other scan: part of every registration
            T, conv = ctx.align_device(cat.data_ptr(), cat.shape[0], 32, Tg[0])
            return [T], [conv], Tt
        Ts, convs = ctx.batch_align(None, offs, Tg, device_ptr=cat.data_ptr(), stride=32)
        return Ts, convs, Tt

    ctx.set_profiling(False)
    for k in range(args.warmup):
        flush.zero_()
        resident_step(k)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    ms, launches, errs, first = [], 0, [], None
    for k in range(args.warmup, n_steps_all):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Ts, convs, Tts = resident_step(k)
        wall_step = 1e3 * (time.perf_counter() - t0)
        st = ctx.stats()
        # device time of the step: CUDA events on the library stream around the whole (batched) align;
        # VGICP also rebuilds its target (the other scan) every step, so its step is timed by the host clock
        ms.append(wall_step if method == "vgicp" else st["ms_total"])
        launches += st["kernel_launches"]
        errs += [pose_err(T, Tt) for T, Tt in zip(Ts, Tts)]
        if first is None:
            first = (batches[k][0], Ts, convs)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall_total = time.perf_counter() - wall0
    t_resident = float(np.sum(ms)) / 1e3

    # ---- roofline pass (untimed for `value`): the same steps again with CUDA events around the hot-kernel sequence
    ctx.set_profiling(True)
    prof = new_prof()
    for k in range(args.warmup, n_steps_all):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        resident_step(k)
        wall_step = 1e3 * (time.perf_counter() - t0)
        add_prof(prof, ctx.stats(), ms_total=wall_step if method == "vgicp" else None)
    ctx.set_profiling(False)

    # ---- single-scan latency (resident): p50 / p95 of one registration at a time
    lat = []
    for k in range(min(args.steps, 12)):
        i = k % n_unique
        s, d, Tg, _ = step_inputs(wl, i)
        flush.zero_()
        torch.cuda.synchronize()
        if method == "vgicp":
            t0 = time.perf_counter()
            dd = dev_dst[i % len(dev_dst)]
            ctx.set_target_device(dd.data_ptr(), dd.shape[0], 32)
            ctx.align_device(uniq_dev[i].data_ptr(), uniq_dev[i].shape[0], 32, Tg)
            lat.append(1e3 * (time.perf_counter() - t0))
        else:
            ctx.align_device(uniq_dev[i].data_ptr(), uniq_dev[i].shape[0], 32, Tg)
            lat.append(ctx.stats()["ms_total"])

    # ---- e2e: the reference-facing call with HOST buffers (target upload + index build + align per call)
    e2e_steps = args.steps
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    e2e_t, e2e_cached_t, e2e_pageable_t, h2d, d2h = [], [], [], 0, 0
    host_dst = pin(wl["dst"]) if method != "vgicp" else None
    for k in range(2):  # warm-up of both host paths (staging buffers of the first call, packer threads): untimed
        s, d, Tg, _ = step_inputs(wl, k % n_unique)
        ctx.scan2map(pin(s), host_dst if host_dst is not None else pin(d), Tg)
        ctx.scan2map(np.array(s, copy=True), np.array(d, copy=True), Tg)
    for k in range(e2e_steps):
        s, d, Tg, _ = step_inputs(wl, k % n_unique)
        hs = pin(s)
        hd = host_dst if host_dst is not None else pin(d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.scan2map(hs, hd, Tg)
        e2e_t.append(time.perf_counter() - t0)
        h2d = hs.nbytes + hd.nbytes
        d2h = 16 * 8 + 4
        if k < 4:  # the same call on ordinary (pageable) host memory, what a pcl::PointCloud hands the adaptor
            sp, dp = np.array(s, copy=True), np.array(d, copy=True)
            t0 = time.perf_counter()
            ctx.scan2map(sp, dp, Tg)
            e2e_pageable_t.append(time.perf_counter() - t0)
        if method != "vgicp":  # static-map (loc.cpp) variant: target stays registered, only the scan crosses PCIe
            t0 = time.perf_counter()
            ctx.align(hs, Tg)
            e2e_cached_t.append(time.perf_counter() - t0)
    t_e2e = float(np.sum(e2e_t))

    if dist is not None:
        t_resident, t_e2e, wall_total = multigpu.max_over_ranks(dist, [t_resident, t_e2e, wall_total], dev)
        launches = int(multigpu.sum_over_ranks(dist, [float(launches)], dev)[0])
    value = world * args.steps * B / t_resident
    e2e_value = world * e2e_steps / t_e2e
    out = None
    if rank == 0:
        cpub, par, c_bar = None, None, None
        if cpu:
            from oracle import pyoracle as orc
            orc.build()
            cores = os.cpu_count() or 1
            n_cpu = 3 if method != "vgicp" else 1
            ids, gT, gconv = first
            tc, ores = [], []
            for k in range(n_cpu):   # the first registrations of the first timed batch: timing + parity in one go
                s, d, Tg, _ = step_inputs(wl, ids[k % len(ids)])
                t0 = time.perf_counter()
                ores.append(oracle_full(method, s, d, Tg, cores))
                tc.append(time.perf_counter() - t0)
            par = parity_vs_oracle([gT[k % len(ids)] for k in range(n_cpu)], [gconv[k % len(ids)] for k in range(n_cpu)], ores)
            cpub = {"value": n_cpu / float(np.sum(tc)), "unit": "registrations/s", "cores": cores, "kind": "port",
                    "sample": "%d full scan2Map calls (index rebuilt per call) of the same workload with the CPU oracle, OpenMP %d threads" % (n_cpu, cores),
                    "ms_per_registration": 1e3 * float(np.mean(tc))}
            s, d, Tg, _ = step_inputs(wl, 0)
            t0 = time.perf_counter()
            oracle_register(method, s, d, Tg, 4)
            cpub["at_4_threads"] = {"ms_per_registration": 1e3 * (time.perf_counter() - t0), "cores": 4}
            if method == "loam":
                # C_bar of SURVEY §8(d), reported by the oracle: population of the 27 one-metre cells around the scan's points
                qs = []
                for u in range(min(4, n_unique)):
                    su, _, Tu, _ = step_inputs(wl, u)
                    qs.append(su[:, :3].astype(np.float64) @ Tu[:3, :3].T + Tu[:3, 3])
                c_bar = float(orc.neighbourhood27(step_inputs(wl, 0)[1], np.concatenate(qs), 1.0, threads=cores))
        s0, d0, _, _ = step_inputs(wl, 0)
        errs = np.array(errs)
        roof = make_roofline(method, prof, c_bar, name)
        idx_roof = measure_index_build(ctx, dev, wl["dst"], None, 0.2, method) if (method != "vgicp" and world == 1) else None
        if method == "vgicp" and prof.get("ms_aux_kernel"):
            # the k-NN behind the covariances dominates a VGICP registration: its own line (16 B per query + 16 B per candidate examined)
            peak, _ = load_peak()
            roof["dominant_kernel"] = {"kernel": "gicp_knn_kernel", "ms_per_step": prof["ms_aux_kernel"] / args.steps,
                                       "share_of_step": prof["ms_aux_kernel"] / prof["ms_total"], "launches": prof["aux_kernel_launches"],
                                       "queries": prof["n_aux_items"], "us_per_launch": 1e3 * prof["ms_aux_kernel"] / max(prof["aux_kernel_launches"], 1)}
        out = {
            "metric": "scan registrations/sec (%s)" % method.upper(), "value": value, "unit": "registrations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_resident / args.steps, "p50_align_ms": float(np.median(lat)),
            "p95_align_ms": float(np.percentile(lat, 95)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"ndt": "f32 pair math, f64 accumulate", "loam": "f32 search, f64 fit", "vgicp": "f64"}[method], "data": "synthetic",
            "config": {"workload": wl["name"], "n_source": int(len(s0)), "n_target": int(len(d0)), "registrations_per_step": B,
                       "step": "one batch of %d independent scan registrations against the static map (pcr_batch_align, device-resident inputs)" % B
                               if method != "vgicp" else "one scan-to-scan registration incl. target covariance / voxel build",
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "timing": "sum over steps of CUDA-event time on the library stream around the whole align; max over ranks",
                       "parallelism": "replicas: %d rank(s), map index broadcast once (NCCL), no per-iteration collective" % world},
            "e2e": {"value": e2e_value, "unit": "registrations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / e2e_steps,
                    "what": "pcr_scan2map(src, dst, pose), ONE scan per call, pinned host buffers: target upload + index build + scan upload + align",
                    "host_threads": n_host_threads, "host_threads_note": "pcr_params.cores of the register: host threads that pack uploads to 16-byte records (from 8 up also pinned sources)",
                    "pageable_ms_per_step": (1e3 * float(np.mean(e2e_pageable_t))) if e2e_pageable_t else None,
                    "pageable_value": (world / float(np.mean(e2e_pageable_t))) if e2e_pageable_t else None,
                    "static_map_ms_per_step": (1e3 * float(np.mean(e2e_cached_t))) if e2e_cached_t else None},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpub, "parity": par,
            "setup": dict(setup, data_generation_s=t_gen), "wall_s_timed_region": wall_total, "map_kernels": idx_roof,
            "pose_error_vs_truth": {"median_m": float(np.median(errs[:, 0])), "median_rad": float(np.median(errs[:, 1]))},
        }
    ctx.close()
    del uniq_dev, batches, flush
    torch.cuda.empty_cache()
    return out


# ======================================================================================================================
# C4: the fixed 1024-scan localisation job, sharded over the ranks (strong scaling)
# ======================================================================================================================
def measure_c4_job(args, method, rank, world, local_rank, dist, dev, cpu=True):
    import torch
    from simpleslam_b200 import capi, multigpu, workloads
    method_id = capi.PCR_NDT if method == "ndt" else capi.PCR_LOAM
    n_host_threads = host_threads(world)
    ctx = capi.Context(method_id, device=local_rank, cores=n_host_threads)
    ds = lambda pts, leaf: ctx.voxel_downsample(pts, leaf)  # noqa: E731
    n_job = args.job_scans
    lo, hi = multigpu.shard(n_job, rank, world)
    setup = {}
    # ---- the static map: rank 0 builds the index, NCCL broadcast, the others import (outside the timed region, reported)
    t0 = time.perf_counter()
    dst, raw_map, n_map = None, None, 0
    if rank == 0:
        dst, raw_map = workloads.c4_map(ds, keep_raw=True)
        n_map = len(dst)
    setup["map_generation_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    if rank == 0:
        ctx.set_target(dst)
    setup["index_build_ms"] = 1e3 * (time.perf_counter() - t0)
    if dist is not None:
        nbytes, dt = multigpu.broadcast_target(ctx, dist, rank, dev)
        setup["index_broadcast_ms"] = 1e3 * dt
        setup["index_bytes"] = nbytes
        setup["index_broadcast_GBps"] = nbytes / dt / 1e9 if dt > 0 else None
        n_map = int(multigpu.max_over_ranks(dist, [float(n_map)], dev)[0])
    # ---- my shard of the scans
    t0 = time.perf_counter()
    scans, truths, guesses = workloads.c4_scans(method, ds, lo, hi, n_job, workers=max(2, (os.cpu_count() or 8) // max(1, min(world, 8))))
    setup["scan_generation_s"] = time.perf_counter() - t0
    n_local = hi - lo
    offs = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint64)
    host_cat = np.ascontiguousarray(np.concatenate(scans)) if n_local else np.zeros((0, 8), np.float32)
    dev_cat = torch.from_numpy(host_cat).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    def job_step(resident=True, host=None):
        """one pass of the whole job: my shard + the all-gather of every rank's poses"""
        if n_local:
            if resident:
                Ts, convs = ctx.batch_align(None, offs, guesses, device_ptr=dev_cat.data_ptr(), stride=32)
            else:
                Ts, convs = ctx.batch_align(host, offs, guesses)
        else:
            Ts, convs = [], np.zeros(0, bool)
        st = ctx.stats() if n_local else None
        if dist is not None:
            allT, allc = multigpu.gather_poses(dist, world, Ts, convs, n_job, dev)
        else:
            allT, allc = np.asarray(Ts), np.asarray(convs)
        return allT, allc, Ts, convs, st

    def timed(n_steps, **kw):
        """barrier + synchronize on both sides of exactly n_steps job steps; device time from CUDA events, wall clock beside it"""
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, last = 0, None
        w0 = time.perf_counter()
        e0.record()
        for _ in range(n_steps):
            flush.zero_()
            ws = time.perf_counter()
            last = job_step(**kw)
            step_log.append((1e3 * (time.perf_counter() - ws), last[4]["ms_total"] if last[4] else 0.0))
            launches += last[4]["kernel_launches"] if last[4] else 0
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        wall = time.perf_counter() - w0
        t_dev = e0.elapsed_time(e1) * 1e-3
        if dist is not None:
            t_dev, wall = multigpu.max_over_ranks(dist, [t_dev, wall], dev)
            launches = int(multigpu.sum_over_ranks(dist, [float(launches)], dev)[0])
        return t_dev, wall, launches, last

    ctx.set_profiling(False)
    step_log = []
    for _ in range(args.warmup):  # same shape as a timed step: the flush kernel's module is loaded lazily on its first launch
        flush.zero_()             # (≈ 0.5 s on a fresh box) and would otherwise land inside the timed region
        job_step()
    t_dev, wall, launches, last = timed(args.steps)
    resident_steps = list(step_log)
    value = args.steps * n_job / t_dev
    allT, allc, myT, myconv, st_last = last

    # ---- roofline pass: one more job step with CUDA events around the hot-kernel sequence
    ctx.set_profiling(True)
    prof = new_prof()
    flush.zero_()
    r = job_step()
    if r[4]:
        add_prof(prof, r[4])
    ctx.set_profiling(False)

    # ---- single-scan latency (resident)
    lat = []
    for k in range(min(n_local, 12)):
        a, b = int(offs[k]), int(offs[k + 1])
        flush.zero_()
        torch.cuda.synchronize()
        ctx.align_device(dev_cat.data_ptr() + a * 32, b - a, 32, guesses[k])
        lat.append(ctx.stats()["ms_total"])

    # ---- e2e: host scans through the reference-facing call; pinned, then pageable
    e2e_steps = max(2, min(args.steps, 5))
    pinned = torch.from_numpy(host_cat).pin_memory().numpy() if n_local else host_cat
    job_step(resident=False, host=pinned)
    t_e2e, wall_e2e, _, _ = timed(e2e_steps, resident=False, host=pinned)
    job_step(resident=False, host=host_cat)  # warm-up of the pageable path (pinned staging, packer threads)
    t_pg, _, _, _ = timed(2, resident=False, host=host_cat)
    e2e_value = e2e_steps * n_job / t_e2e
    h2d_all = host_cat.nbytes
    if dist is not None:
        h2d_all = int(multigpu.sum_over_ranks(dist, [float(host_cat.nbytes)], dev)[0])

    out = None
    if rank == 0:
        cpub, par, c_bar = None, None, None
        if cpu and dst is not None:
            from oracle import pyoracle as orc
            orc.build()
            cores = os.cpu_count() or 1
            K = 3
            t0 = time.perf_counter()
            o_full = oracle_full(method, scans[0], dst, guesses[0], cores)   # reference semantics: index rebuilt inside the call
            t_full = time.perf_counter() - t0
            ores = [o_full]
            t_al = []
            ondt = orc.Ndt(dst, 1.0) if method == "ndt" else None
            for k in range(1, K):
                t0 = time.perf_counter()
                ores.append(ondt.align(scans[k], guesses[k], threads=cores) if ondt is not None else orc.loam_align(scans[k], dst, guesses[k], threads=cores))
                t_al.append(time.perf_counter() - t0)
            par = parity_vs_oracle(myT[:K], myconv[:K], ores)
            par["what"] = "the first %d registrations of the timed job against the CPU oracle" % K
            cpub = {"value": 1.0 / t_full, "unit": "registrations/s", "cores": cores, "kind": "port", "ms_per_registration": 1e3 * t_full,
                    "sample": "1 full scan2Map of the job's first scan with the CPU oracle (index over the %.1fM-pt map rebuilt inside the call, as the "
                              "reference's register does), OpenMP %d threads; %d more alignments for the parity check" % (n_map / 1e6, cores, K - 1),
                    "align_only_ms": 1e3 * float(np.mean(t_al)) if (t_al and ondt is not None) else None}
            if method == "loam":
                qs = [scans[u][:, :3].astype(np.float64) @ guesses[u][:3, :3].T + guesses[u][:3, 3] for u in range(min(4, n_local))]
                c_bar = float(orc.neighbourhood27(dst, np.concatenate(qs), 1.0, threads=cores))
        roof = make_roofline(method, prof, c_bar, ["c4_job_" + method, "c4_" + method])
        idx_roof = measure_index_build(ctx, dev, dst, raw_map, 0.2, method) if world == 1 else None
        out = {
            "metric": "scan registrations/sec (%s)" % method.upper(), "value": value, "unit": "registrations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "p50_align_ms": float(np.median(lat)) if lat else None,
            "p95_align_ms": float(np.percentile(lat, 95)) if lat else None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"ndt": "f32 pair math, f64 accumulate", "loam": "f32 search, f64 fit"}[method], "data": "synthetic",
            "config": {"workload": c4_job_name(method, n_map, n_job), "job_scans": n_job, "scans_per_rank": [multigpu.shard(n_job, r, world)[1] - multigpu.shard(n_job, r, world)[0] for r in range(world)],
                       "n_source": int(len(scans[0])) if n_local else 0, "n_target": int(n_map),
                       "step": "the whole %d-scan job: every rank registers its contiguous shard (pcr_batch_align_device, device-resident scans), then the "
                               "poses of all ranks are all-gathered (NCCL)" % n_job,
                       "l2": "flushed before every step (256 MiB memset inside the timed region); a shard of scans is larger than L2 anyway for NDT",
                       "timing": "CUDA events around exactly K steps, barrier + synchronize on both sides, max over ranks",
                       "parallelism": "%d rank(s): scans sharded, index built on rank 0 and broadcast once over NCCL before the timed region, no collective while a registration runs" % world},
            "e2e": {"value": e2e_value, "unit": "registrations/s", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(n_job * (16 * 8 + 4)),
                    "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps,
                    "what": "pcr_batch_align of every rank's shard from PINNED host memory (scan upload + align + pose read-back) + the gather, per step",
                    "host_threads": n_host_threads, "host_threads_note": "pcr_params.cores of the register: host threads that pack uploads to 16-byte records (from 8 up also pinned sources); the shipped default 4 gives plain DMA of the 32-byte records",
                    "pageable_ms_per_step": 1e3 * t_pg / 2, "pageable_value": 2 * n_job / t_pg},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpub, "parity": par,
            "setup": setup, "wall_s_timed_region": wall, "map_kernels": idx_roof,
            "job": {"converged": int(np.sum(allc)), "of": int(n_job)},
            "rank0_steps_ms": {"wall": [round(a, 3) for a, _ in resident_steps], "library_events": [round(b, 3) for _, b in resident_steps]},
            "pose_error_vs_truth": {"median_m": float(np.median([pose_err(T, Tt)[0] for T, Tt in zip(myT, truths)])) if n_local else None},
        }
    ctx.close()
    del dev_cat, flush
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL writes its banner / debug lines to stdout: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_bench_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank, enabled=(rank == 0))  # started before warm-up: the first sample takes a moment
    cpu = not args.no_cpu_baseline
    w = args.workload
    if w.startswith("c4_job"):
        out = measure_c4_job(args, w.split("_")[-1], rank, world, local_rank, dist, dev, cpu)
    elif w == "c5_lio":
        out = measure_c5(args, rank, world, local_rank, dist, dev, args.frames, cpu)
    else:
        out = measure_replicas(args, w, rank, world, local_rank, dist, dev, cpu)
    sampler.wait_samples(3, timeout=3.0)
    clocks = sampler.stop()  # covers warm-up, the timed steps and the e2e steps of the headline workload
    if out is not None:
        out["clocks"] = clocks
    if w == "c4_job_ndt" and world == 1 and not args.no_workloads:
        # the other BASELINE configs, measured the same way in the same process (N = 1 only)
        subs = {}
        sub_args = argparse.Namespace(**vars(args))
        sub_args.steps, sub_args.warmup = min(args.steps, 10), 3
        for name in ("c4_job_loam", "c1_loam", "c2_ndt", "c3_vgicp", "c3_vgicp_gn", "c5_lio"):
            t0 = time.perf_counter()
            try:
                if name.startswith("c4_job"):
                    r = measure_c4_job(sub_args, "loam", rank, world, local_rank, dist, dev, cpu)
                elif name == "c5_lio":
                    r = measure_c5(sub_args, rank, world, local_rank, dist, dev, args.frames, cpu)
                else:
                    r = measure_replicas(sub_args, name, rank, world, local_rank, dist, dev, cpu)
                r["wall_s"] = time.perf_counter() - t0
                subs[name] = r
            except Exception as e:  # a sub-workload must not take the headline line down
                subs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        out["workloads"] = subs
    if rank == 0 and out is not None:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4_job_ndt",
                    choices=["c4_job_ndt", "c4_job_loam", "c2_ndt", "c1_loam", "c3_vgicp", "c3_vgicp_gn", "c4_loam", "c4_ndt", "c5_lio"])
    ap.add_argument("--job-scans", type=int, default=1024, help="c4_job_*: scans of the fixed job (BASELINE config 4: 1024)")
    ap.add_argument("--frames", type=int, default=2000, help="c5_lio: frames of the sequence (BASELINE config 5: 2000)")
    ap.add_argument("--pcr", default="loam", choices=["loam", "ndt", "vgicp"], help="c5_lio: frontend.pcr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="headline line only (skip the per-config sub-results)")
    ap.add_argument("--batch", type=int, default=0, help="registrations per step (0 = workload default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # timing rules: W >= 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
