import json,sys
for w in sys.argv[1:]:
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%w))
    except Exception as e:
        print(w,'FAILED',e); continue
    r=d["roofline"]
    print(w, "value %.1f"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.1f (%.3f ms, static %s)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["static_map_ms_per_step"]), "kern us %.1f frac %.3f share %.2f launches %d"%(r["avg_launch_us"], r["frac"] or 0, r["kernel_share_of_step"], r["launches"]), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["ms_per_registration"], "p50 %.3f B %s"%(d["p50_align_ms"], d["config"].get("registrations_per_step")), "n", d["config"]["n_source"], d["config"]["n_target"], "err", d["pose_error_vs_truth"])
