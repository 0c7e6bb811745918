set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r03a_pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-parity --configs "alloy" --no-cpu-baseline --no-hooks > gpurun_out/r03a_bench_n1.json 2> gpurun_out/r03a_bench_n1.err
