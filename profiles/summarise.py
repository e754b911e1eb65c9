"""Turn the raw ncu outputs of profiles/capture.sh (gpurun_out/) into the committed, judged artefacts under profiles/:
  <round>_launches_<workload>.txt   per-kernel launch counts, total / mean time and share of the profiled command
  <round>_ncu_<workload>.txt        raw-page metrics + top stall instructions of the `--set full` capture
  <round>_traffic.json              DRAM bytes per launch of the dominant kernel (bench.py reads it for roofline.traffic)
usage: python profiles/summarise.py r01 [workload ...]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
# workload -> captures (name, kernel); the FIRST one is the kernel bench.py reports roofline.traffic for
CAPS = {"c4_job_ndt": [("ndt", "ndt_round_kernel")], "c2_ndt": [("ndt", "ndt_round_kernel")], "c4_ndt": [("ndt", "ndt_round_kernel")],
        "c1_loam": [("search", "loam_search_kernel"), ("fit", "loam_fit_kernel")], "c4_loam": [("search", "loam_search_kernel"), ("fit", "loam_fit_kernel")],
        "c3_vgicp": [("knn", "gicp_knn_kernel"), ("eval", "vgicp_eval_kernel")]}


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("pcr::", "").replace("void ", "")


def launches(rnd, wl):
    path = os.path.join(SRC, "launches_%s_%s.csv" % (rnd, wl))
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Block Size"], r["Grid Size"], 0.0, 0])
        if r["Metric Name"] == "smsp__thread_inst_executed_per_inst_executed.ratio":
            a[4] += float(r["Metric Value"].replace(",", "")); a[5] += 1
            continue
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        t = float(r["Metric Value"].replace(",", "")) * (1e-3 if r["Metric Unit"] == "ns" else 1.0)  # -> us
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --workload %s --steps 2 --warmup 3 --no-cpu-baseline --no-workloads" % wl,
           "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_share_of_step, not absolutes",
           "%-62s %8s %12s %10s %7s %11s  %s" % ("kernel", "launches", "total_us", "mean_us", "share", "lanes/inst", "last block x grid")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-62s %8d %12.1f %10.2f %6.1f%% %11s  %s x %s" % (k[:62], a[0], a[1], a[1] / max(a[0], 1), 100 * a[1] / tot,
                                                                      ("%.1f" % (a[4] / a[5])) if a[5] else "-", a[2], a[3]))
    out.append("%-62s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
    open(os.path.join(OUT, "%s_launches_%s.txt" % (rnd, wl)), "w").write("\n".join(out) + "\n")
    return agg


def full(rnd, wl):
    out, first = [], None
    for name, k in CAPS[wl]:
        rep = os.path.join(SRC, "prof_%s_%s_%s.ncu-rep" % (rnd, wl, name))
        if not os.path.exists(rep):
            continue
        txt = subprocess.run([sys.executable, os.path.join(OUT, "ncu_summary.py"), rep, "25"], capture_output=True, text=True).stdout
        bk = subprocess.run([sys.executable, os.path.join(OUT, "ncu_buckets.py"), rep, "250"], capture_output=True, text=True).stdout
        head = "# ncu --set full --clock-control none --import-source on -k regex:%s (one launch of a steady-state step): bench.py --workload %s\n" % (k, wl)
        out.append(head + txt + "\n# stall samples / executed instructions by SASS range\n" + bk)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(raw.splitlines()))
        hdr, units, vals = r[0], r[1], r[2]

        def get(nm):
            i = hdr.index(nm)
            v = float(vals[i].replace(",", ""))
            u = units[i].lower()
            mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            return v * mult
        rec = {"kernel": k, "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
               "duration_us_under_ncu": get("gpu__time_duration.sum"), "source": os.path.basename(rep)}
        if first is None:
            first = rec
            first["others"] = {}
        else:
            first["others"][name] = rec
    if out:
        open(os.path.join(OUT, "%s_ncu_%s.txt" % (rnd, wl)), "w").write("\n\n".join(out))
    return first


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    wls = sys.argv[2:] or list(CAPS)
    tpath = os.path.join(OUT, "%s_traffic.json" % rnd)
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for wl in wls:
        a = launches(rnd, wl)
        t = full(rnd, wl)
        if t:
            traffic[wl] = t
        print(wl, "launch list:", "ok" if a else "missing", "| full capture:", "ok" if t else "missing")
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
