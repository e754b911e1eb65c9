# GPU session script (round 2, #35): recapture C3 (evaluation kernel changed), final default line
bash profiles/capture.sh r02y "c3_vgicp" > gpurun_out/capture_r02y.log 2>&1
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02_c4_job_ndt.json 2> gpurun_out/bench_r02_c4_job_ndt.err; tail -c 200 gpurun_out/bench_r02_c4_job_ndt.err
ls gpurun_out | grep r02y
