# GPU session script (round 2, #10): parity suite, the headline job (chunk-size sweep for the e2e leg), NDT / LOAM lines
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -25
for mb in 128 256 512 1024; do PCR_BATCH_CHUNK_MB=$mb timeout 600 python bench.py --steps 5 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/b10_job_mb$mb.json 2> gpurun_out/b10_job_mb$mb.err; tail -c 300 gpurun_out/b10_job_mb$mb.err; done
for w in c2_ndt c4_ndt; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b10_$w.json 2> gpurun_out/b10_$w.err; tail -c 300 gpurun_out/b10_$w.err; done
timeout 600 python bench.py --steps 5 --warmup 3 --workload c4_job_loam --no-cpu-baseline > gpurun_out/b10_jobloam.json 2> gpurun_out/b10_jobloam.err
ls gpurun_out | grep b10_
