"""Markdown table of a bench.py JSON line with `workloads` (the driver's line): python profiles/bench_table.py profiles/r02_bench_default.json"""
import json
import sys


def row(name, d):
    if "error" in d:
        return "| %s | error: %s |" % (name, d["error"])
    r = d.get("roofline") or {}
    e = d["e2e"]
    cpu = d.get("cpu_baseline") or {}
    par = d.get("parity") or {}
    unit = "reg/s" if "registr" in d["unit"] else d["unit"]
    p50 = ("%.3f ms" % d["p50_align_ms"]) if d.get("p50_align_ms") else "—"
    pg = e.get("pageable_value")
    return "| %s | %s %s | %s | %s (pageable %s) | %s %s, %s cores | %s: %.1f µs, frac %.3f (examined %s) | %s |" % (
        name, ("%.1f" % d["value"]) if d["value"] < 1000 else ("%.0f" % d["value"]), unit, p50,
        ("%.0f" % e["value"]), ("%.0f" % pg) if pg else "—",
        ("%.2f" % cpu["value"]) if cpu.get("value") is not None else "—", cpu.get("unit", ""), cpu.get("cores", "—"),
        (r.get("kernel") or "").split(" ")[0], r.get("avg_launch_us") or 0.0, r.get("frac") or 0.0,
        ("%.3f" % r["frac_examined"]) if r.get("frac_examined") else "—",
        ("ok: %d poses, max %.1e m / %.1e rad" % (par.get("checked", 0), par.get("max_dt") or 0, par.get("max_dr") or 0)) if par else "—")


def main():
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("| Workload | resident `value` | p50 align | `e2e` (host buffers) | CPU oracle | dominant kernel: µs / launch, frac §8(d) | parity vs oracle |")
    print("|---|---|---|---|---|---|---|")
    print(row("**C4 job NDT** (headline, %d GPU)" % d["n_gpus"], d))
    for k, v in (d.get("workloads") or {}).items():
        print(row(k, v))
    mk = d.get("map_kernels")
    if mk:
        print()
        for k in ("voxel_downsample", "index_build"):
            if k in mk:
                print("* %s: %.2f ms, %.0f GB/s = %.3f of the measured HBM peak (%s)" % (k, mk[k]["ms"], mk[k]["achieved"], mk[k]["frac"], mk[k]["what"]))


if __name__ == "__main__":
    main()
