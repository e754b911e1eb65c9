# GPU session script (round 2, #19): pageable uploads packed by host threads (hostpack.hpp): parity, e2e pageable A/B
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for hp in auto 0; do
  E=""; [ $hp = 0 ] && E="PCR_HOST_PACK=0"
  env $E timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b19_job_hp$hp.json 2> gpurun_out/b19_job_hp$hp.err
  env $E timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b19_c2_ndt_hp$hp.json 2> gpurun_out/b19_c2_ndt_hp$hp.err
  env $E timeout 300 python bench.py --workload c1_loam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b19_c1_loam_hp$hp.json 2> gpurun_out/b19_c1_loam_hp$hp.err
done
nproc; grep -c processor /proc/cpuinfo
