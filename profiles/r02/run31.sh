# GPU session script (round 2, #31): LOAM search with cp.async (LDGSTS) staging of a lane's candidates in shared memory: A/B, parity
for st in 0 1; do
  PCR_LOAM_STAGE=$st timeout 300 python bench.py --workload c4_loam --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b31_c4_loam_st$st.json 2> gpurun_out/b31_c4_loam_st$st.err
  PCR_LOAM_STAGE=$st timeout 300 python bench.py --workload c1_loam --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b31_c1_loam_st$st.json 2> gpurun_out/b31_c1_loam_st$st.err
  PCR_LOAM_STAGE=$st timeout 600 python bench.py --workload c4_job_loam --steps 6 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b31_jobloam_st$st.json 2> gpurun_out/b31_jobloam_st$st.err
done
PCR_LOAM_STAGE=1 timeout 600 python -m pytest tests -m gpu -q -x -k "loam or c4 or batch" 2>&1 | tail -3
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
PCR_LOAM_STAGE=1 timeout 600 ncu --metrics $M --clock-control none -k regex:loam_search --launch-skip 10 -c 6 --csv --log-file gpurun_out/l31_c4_loam.csv python bench.py --workload c4_loam --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/l31_c4_loam.log 2>&1
