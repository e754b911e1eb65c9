# GPU session script (round 2, #32): LOAM variant parity incl. the staged search kernel
timeout 900 python -m pytest tests/test_gpu_loam_variants.py tests/test_gpu_loam.py -m gpu -q 2>&1 | tail -4
