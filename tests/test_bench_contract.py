"""bench.py's one-line JSON contract, checked on the committed line of the last GPU run (profiles/r02_bench_c4_job_ndt.json)
and on the reference-arm line: every key the task statement names is present and consistent. No GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_default_line_has_every_contract_key():
    d = _line("r02_bench_c4_job_ndt.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "parity", "workloads"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "1024" in d["config"]["workload"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0 and d["parity"]["ok"] and d["clocks"]["sm_mhz"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name in ("c4_job_loam", "c1_loam", "c2_ndt", "c3_vgicp", "c3_vgicp_gn", "c5_lio"):
        w = d["workloads"][name]
        assert w["value"] > 0 and w["parity"]["ok"] and w["e2e"]["value"] > 0 and w["cpu_baseline"]["value"] > 0, name


def test_reference_arm_line():
    d = _line("r02_bench_reference.json")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    ours = _line("r02_bench_c4_job_ndt.json")
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and d["higher_is_better"] == ours["higher_is_better"]


def test_cli_parses():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--impl" in r.stdout and "--gpus" in r.stdout and "--workload" in r.stdout
