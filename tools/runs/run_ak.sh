set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r01ak_pytest_gpu.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r01ak_bench_n1.json 2> gpurun_out/r01ak_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01ak_bench_reference.json 2> gpurun_out/r01ak_bench_reference.err
python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01ak_bench_n1_alloy.json 2> gpurun_out/r01ak_bench_n1_alloy.err
python bench.py --steps 50 --warmup 5 --cells 200 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r01ak_bench_200cells_16M.json 2> gpurun_out/r01ak_bench_200.err
python tools/energy_drift.py 100 1000 > gpurun_out/r01ak_energy_drift_fe_1000steps.log 2>&1
python tools/energy_drift.py 100 1000 97 2 1 > gpurun_out/r01ak_energy_drift_alloy_1000steps.log 2>&1
python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r01ak_pka_5keV_2M.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01ak_launches.csv python bench.py --steps 3 --warmup 3 --equil 200 --no-cpu-baseline > gpurun_out/r01ak_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(force|rho)_f' --launch-skip 404 --launch-count 2 -o gpurun_out/r01ak_full python tools/ncu_target.py 100 200 3 > gpurun_out/r01ak_ncu.log 2>&1
