"""Fold the per-N bench lines of profiles/scale.sh into profiles/<round>_scaling.json: python profiles/scale_summary.py r02"""
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {}
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench_%s_n*_*.json" % rnd))):
    m = re.match(r"bench_%s_n(\d+)_(.*)\.json" % rnd, os.path.basename(f))
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    s = d.get("setup", {})
    out.setdefault(m.group(2), {})[m.group(1)] = {
        "value": d["value"], "unit": d["unit"], "scaling": d["scaling"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"],
        "e2e_pageable": d["e2e"].get("pageable_value"), "roofline_frac": d["roofline"]["frac"], "index_broadcast_ms": s.get("index_broadcast_ms"),
        "index_bytes": s.get("index_bytes"), "index_broadcast_GBps": s.get("index_broadcast_GBps"), "clocks": d.get("clocks")}
for wl, rows in out.items():
    base = rows.get("1", {}).get("value")
    for n, r in sorted(rows.items(), key=lambda kv: int(kv[0])):
        r["speedup_vs_1"] = (r["value"] / base) if base else None
        print("%-12s N=%s value %.0f e2e %.0f speedup %s broadcast %s ms" % (wl, n, r["value"], r["e2e"], ("%.2f" % r["speedup_vs_1"]) if base else "-", r["index_broadcast_ms"]))
json.dump(out, open(os.path.join(ROOT, "profiles", "%s_scaling.json" % rnd), "w"), indent=1, sort_keys=True)
