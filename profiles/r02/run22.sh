# GPU session script (round 2, #22): rows / candidates per query and warp lock-step efficiency of the LOAM search (diagnostics)
PCR_LOAM_HIST=1 timeout 300 python bench.py --workload c4_loam --steps 1 --warmup 3 --no-cpu-baseline --no-workloads 2>&1 | grep "loam hist" | tail -2
PCR_LOAM_HIST=1 timeout 300 python bench.py --workload c1_loam --steps 1 --warmup 3 --no-cpu-baseline --no-workloads 2>&1 | grep "loam hist" | tail -2
