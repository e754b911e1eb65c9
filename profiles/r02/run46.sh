# GPU session script (round 2, #46): host-register API + adaptor target cache with pinned target
timeout 600 python -m pytest tests/test_cpp_adaptor.py tests/test_gpu_robustness.py tests/test_capi.py -m gpu -q 2>&1 | tail -4
./tests/cpp/test_adaptor 2>&1 | tail -7
