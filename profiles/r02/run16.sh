# GPU session script (round 2, #16): k-NN bounded by the Morton neighbour's result + Morton sort over the used bits only
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --workload c3_vgicp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b16_c3_vgicp_$i.json 2> gpurun_out/b16_c3_vgicp_$i.err; done
timeout 300 python bench.py --workload c5_lio --frames 500 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b16_c5.json 2> gpurun_out/b16_c5.err
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:gicp_knn -c 6 --csv --log-file gpurun_out/l16_c3.csv python bench.py --workload c3_vgicp --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l16_c3.log 2>&1
