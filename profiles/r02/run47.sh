# GPU session script (round 2, #47): one pack decision per batch: chunked-batch tests, job e2e
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_robustness.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-workloads --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('job value %.0f e2e %.0f (%.1f ms) pageable %.0f threads %s'%(d['value'],e['value'],e['ms_per_step'],e['pageable_value'],e.get('host_threads')))"
