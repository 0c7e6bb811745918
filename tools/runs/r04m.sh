# round 2, session 2: near groups of 16 (rho) / 8 (force) pairs per trip -- A/B against the previous build, force 16, alloy, parity
set -x
mkdir -p gpurun_out
timeout 900 python tools/time_variants.py build/variants/prev_head.so build/variants/force_g16.so > gpurun_out/r04m_variants_fe.log 2>&1
RATIO=97,2,1 timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r04m_variants_alloy.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r04m_pytest.log
