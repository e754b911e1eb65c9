#!/bin/bash
# Capture recipe (B200_PROFILING.md): plain bench first (its numbers are the bench values), then the ncu launch list and one
# `--set full` capture of the dominant kernel of each workload. Run on the GPU box through gpurun:
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r01 "c2_ndt c4_loam c4_ndt"'
# Outputs land in gpurun_out/; profiles/summarise.py turns them into the committed profiles/<round>_*.{csv,txt}.
set -u
ROUND=${1:-r01}
WLS=${2:-"c2_ndt c1_loam c3_vgicp c4_loam c4_ndt"}
STEPS=${STEPS:-4}
mkdir -p gpurun_out
declare -A KERN=( [c2_ndt]=ndt_eval_kernel [c4_ndt]=ndt_eval_kernel [c1_loam]=loam_iter_kernel [c4_loam]=loam_iter_kernel [c3_vgicp]=gicp_knn_kernel )
declare -A SKIP=( [c2_ndt]=40 [c4_ndt]=40 [c1_loam]=10 [c4_loam]=10 [c3_vgicp]=8 )
for wl in $WLS; do
  python bench.py --workload $wl --steps $STEPS --warmup 3 > gpurun_out/bench_${ROUND}_$wl.json 2> gpurun_out/bench_${ROUND}_$wl.err || { echo "bench $wl failed"; tail -5 gpurun_out/bench_${ROUND}_$wl.err; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${ROUND}_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${ROUND}_$wl.log 2>&1
  k=${KERN[$wl]}
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip ${SKIP[$wl]} --launch-count 1 -f \
      -o gpurun_out/prof_${ROUND}_${wl}_$k python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncufull_${ROUND}_$wl.log 2>&1
done
ls -la gpurun_out | tail -30
