# GPU session script (round 2, #40): smallest cloud worth packing on the host (C5 frames upload 0.85 MB raw scans)
for kb in 1024 512 256 64; do PCR_HOST_PACK_MIN_KB=$kb timeout 300 python bench.py --workload c5_lio --frames 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('min_kb $kb frames/s %.0f ms/frame %.3f'%(d['value'],d['ms_per_step']))"; done
