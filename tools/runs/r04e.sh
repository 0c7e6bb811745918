# round 2, session 2: k_force_minor on address-ordered offset lists -- A/B on the 97:2:1 alloy, per-kernel times, alloy parity tests
set -x
mkdir -p gpurun_out
RATIO=97,2,1 timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r04e_variants_alloy.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_force|k_rho' --launch-skip 404 --launch-count 6 --csv --log-file gpurun_out/r04e_alloy_kernel_times.csv python tools/ncu_target.py 100 200 3 97 2 1 > gpurun_out/r04e_alloy_ncu.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r04e_pytest.log
