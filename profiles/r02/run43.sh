# GPU session script (round 2, #43): e2e with 12 host threads packing uploads (pinned sources too): default bench line, parity
timeout 900 python -m pytest tests -m gpu -q -x -k "batch or robust or capi or multi" 2>&1 | tail -3
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02_c4_job_ndt.json 2> gpurun_out/bench_r02_c4_job_ndt.err; tail -c 200 gpurun_out/bench_r02_c4_job_ndt.err
PCR_BENCH_CORES=4 timeout 600 python bench.py --steps 4 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b43_job_cores4.json 2>/dev/null
