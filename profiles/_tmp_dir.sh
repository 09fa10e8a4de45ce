timeout 900 python -m pytest tests/test_gpu_random.py tests/test_gpu_parity.py tests/test_gpu_softmax.py tests/test_gpu_abi.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python profiles/bench_kernels.py detect 2>&1 | grep "_us"
timeout 300 python profiles/bench_kernels.py detect 2>&1 | grep "_us"
