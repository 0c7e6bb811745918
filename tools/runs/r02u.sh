set -x
timeout 600 python tools/time_variants.py > gpurun_out/r02u_variants.log 2>&1
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 1000 > gpurun_out/r02u_pka_profile.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_inter.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02u_pytest.log
