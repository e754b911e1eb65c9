# GPU session script (round 2, #34): finer trace of a single-scan NDT round (block 0's phases)
for m in "ndt c2" "ndt c4"; do echo "== $m"; PCR_NDT_TAIL_TRACE=1 timeout 300 python profiles/r02/lat_probe.py $m 2>&1 | tail -3; done
