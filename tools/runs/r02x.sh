set -x
timeout 1200 python -m pytest tests/test_gpu_inter.py tests/test_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02x_pytest_gpu.log
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r02x_pka_profile.log 2>&1
