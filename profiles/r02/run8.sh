# GPU session script (round 2, #8): parity suite, bench lines, k-NN min-pop sweep, launch metrics
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -25
for w in c1_loam c4_loam c3_vgicp; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b8_$w.json 2> gpurun_out/b8_$w.err; tail -c 300 gpurun_out/b8_$w.err; done
for mp in 4 6 8 10 12; do PCR_KNN_MINPOP=$mp timeout 300 python bench.py --workload c3_vgicp --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b8_c3mp$mp.json 2>/dev/null; done
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:loam --launch-skip 20 -c 20 --csv --log-file gpurun_out/l8_c4_loam.csv python bench.py --workload c4_loam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l8_c4_loam.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:gicp_knn --launch-skip 4 -c 4 --csv --log-file gpurun_out/l8_c3.csv python bench.py --workload c3_vgicp --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l8_c3.log 2>&1
ls gpurun_out | grep 8_
