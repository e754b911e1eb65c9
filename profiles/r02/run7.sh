# GPU session script (round 2, #7): parity suite, bench lines, launch metrics, NDT full capture
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -25
for w in c1_loam c4_loam c2_ndt c4_ndt c3_vgicp; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b7_$w.json 2> gpurun_out/b7_$w.err; tail -c 300 gpurun_out/b7_$w.err; done
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:loam --launch-skip 20 -c 20 --csv --log-file gpurun_out/l7_c4_loam.csv python bench.py --workload c4_loam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l7_c4_loam.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:ndt_round --launch-skip 30 -c 30 --csv --log-file gpurun_out/l7_c4_ndt.csv python bench.py --workload c4_ndt --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l7_c4_ndt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ndt_round --launch-skip 30 --launch-count 1 -f -o gpurun_out/prof_r02_c4_ndt python bench.py --workload c4_ndt --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncufull_r02_c4_ndt.log 2>&1
ls gpurun_out | grep 7_
