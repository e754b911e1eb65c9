# GPU session script (round 2, #36): NDT accumulators paired in shared memory (LDS.128 / STS.128): parity, job / C2 / single scan
timeout 900 python -m pytest tests -m gpu -q -x -k "ndt or batch or multi or c4 or frontend or robust" 2>&1 | tail -3
timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b36_job.json 2> gpurun_out/b36_job.err
timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b36_c2_ndt.json 2> gpurun_out/b36_c2_ndt.err
for m in "ndt c2"; do PCR_NDT_TAIL_TRACE=1 timeout 300 python profiles/r02/lat_probe.py $m 2>&1 | tail -2; done
