# shape sweep of the LOAM iteration kernel (tuning aid): lanes per query x cell-width policy
python -m pytest tests/test_gpu_loam.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
for wl in c1_loam c4_loam; do
 for fine in ${FINES:-6}; do
  for lpq in ${LPQS:-1 2 4 8}; do
    PCR_LOAM_LPQ=$lpq PCR_LOAM_FINE_ABOVE=$fine python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sw_${wl}_f${fine}_l${lpq}.json 2>/dev/null || echo fail $wl $fine $lpq
  done
 done
done
