# round 2, session 2: four staged offsets per 16-byte load in the near loops (EAM_OFF_V4) -- A/B on pure Fe and the alloy, parity tests
set -x
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py build/variants/off_v4_0.so > gpurun_out/r04i_variants_fe.log 2>&1
timeout 600 python tools/time_variants.py build/variants/off_v4_0.so >> gpurun_out/r04i_variants_fe.log 2>&1
RATIO=97,2,1 timeout 600 python tools/time_variants.py build/variants/off_v4_0.so > gpurun_out/r04i_variants_alloy.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r04i_pytest.log
