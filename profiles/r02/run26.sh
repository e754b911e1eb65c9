# GPU session script (round 2, #26): every kernel family once on small inputs (profiles/r02/sanitize.py). compute-sanitizer is
# closed on this pool, so the script only checks that the forced variants (split LOAM kernels on a small batch, single-scan
# and 3-scan NDT, VGICP device-side LM) run and converge.
timeout 300 python profiles/r02/sanitize.py 2>&1 | tail -6
