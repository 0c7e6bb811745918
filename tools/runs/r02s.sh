set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02s_pytest_gpu.log
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 1000 > gpurun_out/r02s_pka_profile.log 2>&1
