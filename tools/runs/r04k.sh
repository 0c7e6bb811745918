# round 2, session 2: one LDS.128 for the near offsets, far loop four offsets per load -- A/B against the previous build, parity
set -x
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py build/variants/prev_head.so build/variants/nc.so > gpurun_out/r04k_variants_fe.log 2>&1
timeout 600 python tools/time_variants.py build/variants/prev_head.so >> gpurun_out/r04k_variants_fe.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r04k_pytest.log
