set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r01ao_pytest_gpu.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r01ao_bench_n1.json 2> gpurun_out/r01ao_bench_n1.err
