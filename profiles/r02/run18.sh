# GPU session script (round 2, #18): LOAM search with an L1 prefetch of the next row's first points (A/B), parity with it on
for pf in 0 1; do
  PCR_LOAM_PREFETCH=$pf timeout 300 python bench.py --workload c4_loam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b18_c4_loam_pf$pf.json 2> gpurun_out/b18_c4_loam_pf$pf.err
  PCR_LOAM_PREFETCH=$pf timeout 300 python bench.py --workload c1_loam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b18_c1_loam_pf$pf.json 2> gpurun_out/b18_c1_loam_pf$pf.err
  PCR_LOAM_PREFETCH=$pf timeout 600 python bench.py --workload c4_job_loam --steps 6 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b18_jobloam_pf$pf.json 2> gpurun_out/b18_jobloam_pf$pf.err
done
PCR_LOAM_PREFETCH=1 timeout 600 python -m pytest tests -m gpu -q -x -k "loam or c4 or batch" 2>&1 | tail -3
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
PCR_LOAM_PREFETCH=1 timeout 600 ncu --metrics $M --clock-control none -k regex:loam_search --launch-skip 10 -c 6 --csv --log-file gpurun_out/l18_c4_loam.csv python bench.py --workload c4_loam --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/l18_c4_loam.log 2>&1
