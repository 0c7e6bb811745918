set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02m_pytest_parity.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-parity --configs "" --no-cpu-baseline --no-hooks > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
for T in 3 5 8; do
MISA_B200_OPTS=slab_planes=$T timeout 600 python bench.py --steps 20 --warmup 10 --no-parity --configs "" --no-cpu-baseline --no-hooks > gpurun_out/r02m_bench_n1_T$T.json 2> gpurun_out/r02m_bench_n1_T$T.err
done
