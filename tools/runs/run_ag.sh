set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r01ag_pytest_gpu.log
python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01ag_bench_n1_alloy.json 2> gpurun_out/r01ag_bench_n1_alloy.err
MISA_B200_OPTS=minor_staged=0 python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01ag_bench_n1_alloy_global.json 2> gpurun_out/r01ag_bench_n1_alloy_global.err
