# round 2, session 2: two-GPU parity tests on the final kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r04o_pytest_gpu_2gpus.log
