#!/bin/bash
# 1 -> N GPU scaling of the batched workloads (torchrun, NCCL broadcast of the index): bash profiles/scale.sh r01 "2 4 8"
ROUND=${1:-r01}
for n in ${2:-"2"}; do
  for wl in c2_ndt c4_loam c4_ndt; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 \
      --workload $wl --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/bench_${ROUND}_n${n}_$wl.json || echo "failed $n $wl"
  done
done
