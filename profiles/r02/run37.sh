# GPU session script (round 2, #37): NDT point prefetch one trip ahead: parity, job / C2 / C4 batch
timeout 900 python -m pytest tests -m gpu -q -x -k "ndt or batch or multi or c4" 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b37_job_$i.json 2> gpurun_out/b37_job_$i.err; done
timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b37_c2_ndt.json 2> gpurun_out/b37_c2_ndt.err
timeout 300 python bench.py --workload c4_ndt --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b37_c4_ndt.json 2> gpurun_out/b37_c4_ndt.err
