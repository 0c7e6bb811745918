# round 2, session 2: final single-GPU evidence of the committed tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r04z_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_verlet|k_(force|rho)_f' --launch-skip 300 --launch-count 6 -o gpurun_out/r04z_full python tools/ncu_target.py 100 100 3 > gpurun_out/r04z_ncu.log 2>&1
ncu -i gpurun_out/r04z_full.ncu-rep --page raw --csv > gpurun_out/r04z_full_raw.csv 2>/dev/null
python tools/make_traffic.py gpurun_out/r04z_full_raw.csv "profiles/r04z_ncu_full_summary.txt (ncu --set full --clock-control none, tools/ncu_target.py 100 100 3: bcc Fe 100^3 cells thermalised 100 steps; per launch)" > gpurun_out/r04z_make_traffic.log 2>&1
cp profiles/traffic.json gpurun_out/r04z_traffic.json
python tools/ncu_summary.py gpurun_out/r04z_full_raw.csv > gpurun_out/r04z_ncu_full_summary.txt 2>&1
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r04z_bench_n1.json 2> gpurun_out/r04z_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r04z_bench_reference.json 2> gpurun_out/r04z_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r04z_launches.csv python bench.py --steps 3 --warmup 3 --equil 60 --no-cpu-baseline --no-hooks --no-parity --configs "" > gpurun_out/r04z_launches_bench.log 2>&1
timeout 600 python tools/energy_drift.py 100 1000 > gpurun_out/r04z_energy_drift_fe_1000steps.log 2>&1
timeout 600 python tools/energy_drift.py 100 1000 97 2 1 > gpurun_out/r04z_energy_drift_alloy_1000steps.log 2>&1
timeout 600 python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r04z_pka_5keV_2M.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r04z_smoke.log 2>&1
