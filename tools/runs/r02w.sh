set -x
timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r02w_variants.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02w_pytest_gpu.log
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r02w_pka_profile.log 2>&1
