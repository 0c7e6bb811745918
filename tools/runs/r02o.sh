set -x
timeout 900 python -m pytest tests/test_gpu_inter.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02o_pytest_inter.log
timeout 600 python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r02o_pka_5keV_2M.log 2>&1
