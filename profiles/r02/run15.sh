# GPU session script (round 2, #15): NDT warp-per-scan init + programmatic dependent launch of the round kernels: parity, A/B
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for pdl in 1 0; do for m in "ndt c2" "ndt c4"; do echo "== $m PDL=$pdl"; PCR_NDT_PDL=$pdl PCR_NDT_TAIL_TRACE=1 timeout 300 python profiles/r02/lat_probe.py $m 2>&1 | tail -2; done; done
for pdl in 1 0; do
PCR_NDT_PDL=$pdl timeout 600 python bench.py --steps 6 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b15_job_pdl$pdl.json 2> gpurun_out/b15_job_pdl$pdl.err
PCR_NDT_PDL=$pdl timeout 300 python bench.py --workload c2_ndt --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b15_c2_ndt_pdl$pdl.json 2> gpurun_out/b15_c2_ndt_pdl$pdl.err
done
