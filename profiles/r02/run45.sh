# GPU session script (round 2, #45): default bench line with 12 host threads (pinned batches >= 256 MB packed)
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02_c4_job_ndt.json 2> gpurun_out/bench_r02_c4_job_ndt.err; tail -c 200 gpurun_out/bench_r02_c4_job_ndt.err
