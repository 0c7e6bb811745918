set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r01an_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r01an_smoke.log 2>&1
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r01an_bench_n1.json 2> gpurun_out/r01an_bench_n1.err
