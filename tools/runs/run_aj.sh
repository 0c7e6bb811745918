set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r01aj_pytest_gpu.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r01aj_bench_n1.json 2> gpurun_out/r01aj_bench_n1.err
python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01aj_bench_n1_alloy.json 2> gpurun_out/r01aj_bench_n1_alloy.err
