# GPU session script (round 2, #6): parity suite, LOAM bench lines + launch metrics after the batch-path rewrite
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -25
for w in c1_loam c4_loam; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b6_$w.json 2> gpurun_out/b6_$w.err; tail -c 300 gpurun_out/b6_$w.err; done
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for w in c4_loam c1_loam; do
timeout 600 ncu --metrics $M --clock-control none -k regex:loam --launch-skip 20 -c 22 --csv --log-file gpurun_out/l6_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l6_$w.log 2>&1
done
timeout 600 ncu --metrics $M --clock-control none -k regex:ndt_round --launch-skip 30 -c 30 --csv --log-file gpurun_out/l6_c4_ndt.csv python bench.py --workload c4_ndt --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l6_c4_ndt.log 2>&1
ls gpurun_out | grep 6_
