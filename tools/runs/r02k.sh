set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r02k_bench_n8.json 2> gpurun_out/r02k_bench_n8.err
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02k_pytest_gpu_8gpus.log
python bench.py --steps 100 --warmup 10 --no-parity --configs "" --no-cpu-baseline --no-hooks > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err
