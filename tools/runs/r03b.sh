set -x
timeout 600 python tools/time_variants.py build/variants/mono0.so > gpurun_out/r03b_variants.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r03b_pytest.log
