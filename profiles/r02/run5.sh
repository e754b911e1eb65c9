# GPU session script (round 2, #5): the default bench line (headline C4 job + sub-workloads), then ncu --set full captures
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/b5_default.json 2> gpurun_out/b5_default.err; tail -c 600 gpurun_out/b5_default.err
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 --launch-skip $4 --launch-count 1 -f -o gpurun_out/prof_r02_$1 python bench.py --workload $2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncufull_r02_$1.log 2>&1
}
cap c4_loam_search c4_loam loam_search 8
cap c4_loam_fit c4_loam "loam_iter_kernel" 8
cap c2_ndt c2_ndt "ndt_round_kernelILi2ELb0" 4
cap c3_knn c3_vgicp gicp_knn 4
ls -la gpurun_out | grep -E "prof_r02|b5_"
