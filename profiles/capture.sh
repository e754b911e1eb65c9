#!/bin/bash
# Capture recipe (B200_PROFILING.md): plain bench first (its numbers are the bench values), then the ncu launch list and
# `--set full` captures of the dominant kernels of each workload. Run on the GPU box through gpurun:
#   gpurun --timeout 2400 -- 'bash profiles/capture.sh r02 "c4_job_ndt c4_ndt c4_loam c2_ndt c1_loam c3_vgicp"'
# c4_job_ndt is bench.py's default (the headline line); its plain run also carries the per-config `workloads`.
# Outputs land in gpurun_out/; profiles/summarise.py turns them into the committed profiles/<round>_*.{txt,json}.
set -u
ROUND=${1:-r02}
WLS=${2:-"c4_job_ndt c4_ndt c4_loam c2_ndt c1_loam c3_vgicp"}
STEPS=${STEPS:-4}
mkdir -p gpurun_out
# workload -> "name:kernel-regex:launch-skip" (several captures per workload)
declare -A CAPS=(
  [c4_job_ndt]="ndt:ndt_round_kernel:8"
  [c2_ndt]="ndt:ndt_round_kernel:4"
  [c4_ndt]="ndt:ndt_round_kernel:4"
  [c1_loam]="search:loam_search_kernel:9 fit:loam_fit_kernel:9"
  [c4_loam]="search:loam_search_kernel:9 fit:loam_fit_kernel:9"
  [c3_vgicp]="knn:gicp_knn_kernel:4 eval:vgicp_eval_kernel:12"
)
for wl in $WLS; do
  EXTRA="--no-workloads"; [ $wl = c4_job_ndt ] && EXTRA=""
  python bench.py --workload $wl --steps $STEPS --warmup 3 $EXTRA > gpurun_out/bench_${ROUND}_$wl.json 2> gpurun_out/bench_${ROUND}_$wl.err || { echo "bench $wl failed"; tail -5 gpurun_out/bench_${ROUND}_$wl.err; continue; }
  ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${ROUND}_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/ncu_${ROUND}_$wl.log 2>&1
  for cap in ${CAPS[$wl]}; do
    IFS=: read name k skip <<< "$cap"
    ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 -f \
        -o gpurun_out/prof_${ROUND}_${wl}_$name python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/ncufull_${ROUND}_${wl}_$name.log 2>&1
  done
done
ls -la gpurun_out | grep ${ROUND} | tail -40
