# GPU session script (round 2, #14): NDT two-level request reduction: parity, single-scan latency with the tail trace, job line
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for m in "ndt c2" "ndt c4"; do echo "== $m"; PCR_NDT_TAIL_TRACE=1 timeout 300 python profiles/r02/lat_probe.py $m 2>&1 | tail -4; done
timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b14_job.json 2> gpurun_out/b14_job.err
timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b14_c2_ndt.json 2> gpurun_out/b14_c2_ndt.err
timeout 300 python bench.py --workload c4_ndt --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b14_c4_ndt.json 2> gpurun_out/b14_c4_ndt.err
