# GPU session script (round 2, #12): default bench line after the warm-up fix (with workloads), VGICP after the device-side LM, single-scan latency probes
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/b12_default.json 2> gpurun_out/b12_default.err; tail -c 300 gpurun_out/b12_default.err
timeout 300 python bench.py --workload c3_vgicp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b12_c3_vgicp.json 2> gpurun_out/b12_c3_vgicp.err; tail -c 300 gpurun_out/b12_c3_vgicp.err
for m in "ndt c2" "ndt c4" "loam c2" "vgicp c2"; do echo "== $m"; timeout 300 python profiles/r02/lat_probe.py $m 2>&1 | tail -4; done
M=gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/l12_lat_ndt_c4.csv python profiles/r02/lat_probe.py ndt c4 > gpurun_out/l12_lat_ndt_c4.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/l12_lat_ndt_c2.csv python profiles/r02/lat_probe.py ndt c2 > gpurun_out/l12_lat_ndt_c2.log 2>&1
ls gpurun_out | grep 12_
