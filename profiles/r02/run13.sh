# GPU session script (round 2, #13): NDT request tail after the row-parallel final reduce: parity, single-scan latency with the tail trace
timeout 900 python -m pytest tests -m gpu -q -x -k "ndt or batch or multi" 2>&1 | tail -4
for m in "ndt c2" "ndt c4"; do echo "== $m"; PCR_NDT_TAIL_TRACE=1 timeout 300 python profiles/r02/lat_probe.py $m 2>&1 | tail -6; done
timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b13_job.json 2> gpurun_out/b13_job.err
timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b13_c2_ndt.json 2> gpurun_out/b13_c2_ndt.err
