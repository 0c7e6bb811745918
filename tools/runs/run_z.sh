set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r01z_pytest_gpu.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r01z_bench_n1.json 2> gpurun_out/r01z_bench_n1.err
python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01z_bench_n1_alloy.json 2> gpurun_out/r01z_bench_n1_alloy.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01z_launches.csv python bench.py --steps 3 --warmup 3 --equil 20 --no-cpu-baseline > gpurun_out/r01z_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(force|rho)_f' --launch-skip 404 --launch-count 2 -o gpurun_out/r01z_full python tools/ncu_target.py 100 200 3 > gpurun_out/r01z_ncu.log 2>&1
