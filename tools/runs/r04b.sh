# round 2, session 2: 256-bit monomial rows for every pair outside the staged tables + minority entries beyond the warp's prefix
# skipped -- A/B against the previous head on the 97:2:1 alloy and on pure Fe, the whole GPU test suite, the cascade profile
set -x
mkdir -p gpurun_out
RATIO=97,2,1 timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r04b_variants_alloy.log 2>&1
timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r04b_variants_fe.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r04b_pytest_gpu.log
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r04b_pka_profile.log 2>&1
