set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r01x_pytest_gpu.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r01x_bench_n1.json 2> gpurun_out/r01x_bench_n1.err
python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01x_bench_n1_alloy.json 2> gpurun_out/r01x_bench_n1_alloy.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01x_launches.csv python bench.py --steps 3 --warmup 3 --equil 20 --no-cpu-baseline > gpurun_out/r01x_b.log 2>&1
