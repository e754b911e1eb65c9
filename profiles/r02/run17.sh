# GPU session script (round 2, #17): NDT float kernel at 6 blocks / SM (80 registers, variant library) vs 5 (96); the reference arm line
for v in "" fb6; do
  L=""; [ -n "$v" ] && L=$PWD/simpleslam_b200/csrc/libpcr_cuda_$v.so
  PCR_LIB=$L timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b17_job_$v.json 2> gpurun_out/b17_job_$v.err
  PCR_LIB=$L timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b17_c2_ndt_$v.json 2> gpurun_out/b17_c2_ndt_$v.err
  echo "== variant '$v'"; PCR_LIB=$L PCR_NDT_TAIL_TRACE=1 timeout 300 python profiles/r02/lat_probe.py ndt c2 2>&1 | tail -2
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -c 600 gpurun_out/bench_r02_reference.json
