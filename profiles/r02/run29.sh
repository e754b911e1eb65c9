# GPU session script (round 2, #29): programmatic dependent launch at 128 / 256 / 512 scans per batch (a rank of an 8 / 4 / 2-GPU job)
for n in 128 256 512; do for pdl in 0 1; do PCR_NDT_PDL=$pdl timeout 300 python bench.py --job-scans $n --steps 6 --warmup 3 --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('scans $n pdl $pdl value %.0f ms/step %.3f'%(d['value'],d['ms_per_step']))"; done; done
