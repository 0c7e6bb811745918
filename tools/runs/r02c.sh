set -x
python tools/hooks_step.py 20 2 > gpurun_out/r02c_hooks20.log 2>&1
python tools/hooks_step.py 100 3 > gpurun_out/r02c_hooks100.log 2>&1
python bench.py --steps 100 --warmup 10 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
