# GPU session script (round 2, #30): final state: full GPU suite, smoke, default bench line (with all sub-workloads), reference arm
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02_c4_job_ndt.json 2> gpurun_out/bench_r02_c4_job_ndt.err; tail -c 200 gpurun_out/bench_r02_c4_job_ndt.err
timeout 400 python bench.py --workload c5_lio --steps 1 --warmup 3 --no-workloads > gpurun_out/bench_r02_c5_lio.json 2> gpurun_out/bench_r02_c5_lio.err
