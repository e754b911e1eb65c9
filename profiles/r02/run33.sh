# GPU session script (round 2, #33): VGICP evaluation kernel with a sliced final reduction: parity, C3
timeout 600 python -m pytest tests -m gpu -q -x -k "vgicp or frontend or robust" 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --workload c3_vgicp --steps 10 --warmup 3 --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('c3 value %.1f ms/step %.3f p50 %.3f eval us %.1f'%(d['value'],d['ms_per_step'],d['p50_align_ms'],r['avg_launch_us']))"; done
