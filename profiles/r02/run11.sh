# GPU session script (round 2, #11): parity, LOAM lookahead A/B (launch metrics), stability of the job line (per-step times)
timeout 900 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -8
for w in c1_loam c4_loam; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b11_$w.json 2> gpurun_out/b11_$w.err; tail -c 300 gpurun_out/b11_$w.err; done
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:loam --launch-skip 20 -c 12 --csv --log-file gpurun_out/l11_c4_loam.csv python bench.py --workload c4_loam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l11_c4_loam.log 2>&1
for i in 1 2 3 4; do timeout 600 python bench.py --steps 8 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b11_job_$i.json 2> gpurun_out/b11_job_$i.err; done
for i in 1 2; do PCR_NDT_LOOKAHEAD=64 timeout 600 python bench.py --steps 8 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b11_jobla_$i.json 2> gpurun_out/b11_jobla_$i.err; done
ls gpurun_out | grep b11_
