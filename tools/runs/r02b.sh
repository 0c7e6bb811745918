# round 2, second GPU call: parity tolerance fixed, streaming kernels at 32 registers, activity words through mapped memory
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02b_pytest_gpu.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
python tools/time_variants.py build/variants/*.so > gpurun_out/r02b_variants.log 2>&1
./tools/pcie_probe > gpurun_out/r02b_pcie_probe.log 2>&1
python tools/pka_cascade.py 100 5000 2000 > gpurun_out/r02b_pka_5keV_2M.log 2>&1
nvidia-smi topo -m > gpurun_out/r02b_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02b_topo.txt; numactl -H >> gpurun_out/r02b_topo.txt 2>&1
