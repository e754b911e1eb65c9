# GPU session script (round 2, #28): fused frame call + slab allocation of small device buffers: parity, C5 loop, frame stages
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python profiles/prof_c5_host.py 2>&1 | grep -E "ms/frame|submap_build|downsample_align|generateOdom" | head
PCR_TRACE=1 timeout 300 python profiles/prof_c5_host.py 2>&1 | grep -i "pcr trace" | awk '{print $12}' | sort -n | tail -5
timeout 600 python bench.py --workload c5_lio --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b28_c5.json 2> gpurun_out/b28_c5.err; tail -c 300 gpurun_out/b28_c5.err
timeout 300 python bench.py --workload c1_loam --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/b28_c1_loam.json 2> gpurun_out/b28_c1_loam.err
