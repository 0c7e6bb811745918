set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r01af_pytest_gpu.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r01af_bench_n1.json 2> gpurun_out/r01af_bench_n1.err
MISA_B200_OPTS=mark=0 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r01af_bench_n1_nomark.json 2> gpurun_out/r01af_bench_n1_nomark.err
python bench.py --steps 100 --warmup 10 --ratio 97 2 1 --no-cpu-baseline > gpurun_out/r01af_bench_n1_alloy.json 2> gpurun_out/r01af_bench_n1_alloy.err
