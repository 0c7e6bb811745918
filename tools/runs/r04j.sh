# round 2, session 2: slope of elec[maj] from three numbers (EAM_ELEC3: one 128-bit + one 64-bit gather instead of two 128-bit) -- A/B, parity
set -x
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py build/variants/elec3_0.so > gpurun_out/r04j_variants_fe.log 2>&1
timeout 600 python tools/time_variants.py build/variants/elec3_0.so >> gpurun_out/r04j_variants_fe.log 2>&1
RATIO=97,2,1 timeout 600 python tools/time_variants.py build/variants/elec3_0.so > gpurun_out/r04j_variants_alloy.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r04j_pytest.log
