# round 2, first GPU call: tests (incl. whole-box parity at 50^3 / 100^3), bench line, A/B of the no-conflict ceiling,
# launch list, ncu full of the streaming + stencil kernels
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02a_pytest_gpu.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err
python tools/time_variants.py build/variants/*.so > gpurun_out/r02a_variants.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python tools/ncu_target.py 100 50 5 > gpurun_out/r02a_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_verlet|k_(force|rho)_f' --launch-skip 300 --launch-count 6 -o gpurun_out/r02a_full python tools/ncu_target.py 100 100 3 > gpurun_out/r02a_ncu.log 2>&1
