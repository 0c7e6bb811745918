set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02p_pytest_gpu.log
timeout 600 python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r02p_pka_5keV_2M.log 2>&1
