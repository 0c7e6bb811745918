set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02y_pytest_gpu_2gpus.log
