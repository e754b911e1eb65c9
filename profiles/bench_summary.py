"""One line per bench JSON: python profiles/bench_summary.py gpurun_out/bench_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "FAILED", e)
        continue
    r = d["roofline"]
    cpu = d.get("cpu_baseline") or {}
    print("%s\n  value %.1f/s  ms/step %.3f  p50 %.3f ms | e2e %.1f/s (%.3f ms/step, static-map %s) | kernel %s: %.1f us/launch, frac %.3f (%.0f GB/s), share %.2f, launches %d, pts/launch %.0f, pairs/pt %.1f | cpu %.2f ms/reg (%s cores) | n %s / %s | setup %s | err %s | clocks %s"
          % (f, d["value"], d["ms_per_step"], d.get("p50_align_ms", float("nan")), d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("static_map_ms_per_step"),
             r["kernel"], r["avg_launch_us"], r["frac"] or 0, r["achieved"] or 0, r["kernel_share_of_step"] or 0, r["launches"], r.get("points_per_launch", 0),
             r.get("pairs_per_point", 0), cpu.get("ms_per_registration", float("nan")), cpu.get("cores"), d["config"].get("n_source"), d["config"].get("n_target"),
             d.get("setup"), d.get("pose_error_vs_truth"), d.get("clocks")))
