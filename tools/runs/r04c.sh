# round 2, session 2: GPU tests of the tree with the 256-bit monomial rows, ncu --set full capture of the current kernel sources ->
# profiles/traffic.json (stamped with their hash), bench line with the PKA config, per-kernel times of the alloy step
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r04c_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_verlet|k_(force|rho)_f' --launch-skip 300 --launch-count 6 -o gpurun_out/r04c_full python tools/ncu_target.py 100 100 3 > gpurun_out/r04c_ncu.log 2>&1
ncu -i gpurun_out/r04c_full.ncu-rep --page raw --csv > gpurun_out/r04c_full_raw.csv 2>/dev/null
python tools/make_traffic.py gpurun_out/r04c_full_raw.csv "profiles/r04c_ncu_full_summary.txt (ncu --set full --clock-control none, tools/ncu_target.py 100 100 3: bcc Fe 100^3 cells thermalised 100 steps; per launch)" > gpurun_out/r04c_make_traffic.log 2>&1
cp profiles/traffic.json gpurun_out/r04c_traffic.json
python tools/ncu_summary.py gpurun_out/r04c_full_raw.csv > gpurun_out/r04c_ncu_full_summary.txt 2>&1
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r04c_bench_n1.json 2> gpurun_out/r04c_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_force|k_rho' --launch-skip 404 --launch-count 6 --csv --log-file gpurun_out/r04c_alloy_kernel_times.csv python tools/ncu_target.py 100 200 3 97 2 1 > gpurun_out/r04c_alloy_ncu.log 2>&1
rm -f gpurun_out/r04c_full.ncu-rep.tmp
