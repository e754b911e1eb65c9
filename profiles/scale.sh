#!/bin/bash
# Strong scaling of the headline job (1024 scans, 18.6 M-pt map; torchrun, one process per GPU, NCCL broadcast of the index):
#   gpurun --gpus 8 -- 'bash profiles/scale.sh r02 "8 4 2 1"'
# Writes gpurun_out/bench_<round>_n<N>_<workload>.json; profiles/scale_summary.py folds them into profiles/<round>_scaling.json.
ROUND=${1:-r02}
WLS=${WLS:-"c4_job_ndt c4_job_loam"}
for n in ${2:-"2"}; do
  for wl in $WLS; do
    if [ $n = 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 --workload $wl --no-cpu-baseline --no-workloads 2> gpurun_out/bench_${ROUND}_n${n}_$wl.err | grep "^{" > gpurun_out/bench_${ROUND}_n${n}_$wl.json || echo "failed $n $wl"
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 6 --warmup 3 \
        --workload $wl --no-cpu-baseline --no-workloads 2> gpurun_out/bench_${ROUND}_n${n}_$wl.err | grep "^{" > gpurun_out/bench_${ROUND}_n${n}_$wl.json || { echo "failed $n $wl"; tail -5 gpurun_out/bench_${ROUND}_n${n}_$wl.err; }
    fi
  done
done
ls -la gpurun_out | grep "bench_${ROUND}_n"
