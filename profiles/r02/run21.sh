# GPU session script (round 2, #21): default bench line (all sub-workloads) after the host-pack / warm-up changes, per-workload lines, LOAM C4 recapture
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02_c4_job_ndt.json 2> gpurun_out/bench_r02_c4_job_ndt.err; tail -c 200 gpurun_out/bench_r02_c4_job_ndt.err
for wl in c2_ndt c1_loam c3_vgicp c4_loam c4_ndt; do timeout 400 python bench.py --workload $wl --steps 6 --warmup 3 --no-workloads > gpurun_out/bench_r02_$wl.json 2> gpurun_out/bench_r02_$wl.err; done
bash profiles/capture.sh r02x "c4_loam" > gpurun_out/capture_r02x.log 2>&1
ls gpurun_out | grep r02x
