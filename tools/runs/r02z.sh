set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r02z_bench_n8.json 2> gpurun_out/r02z_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r02z_bench_n4.json 2> gpurun_out/r02z_bench_n4.err
python bench.py --steps 100 --warmup 10 --no-parity --configs "" --no-cpu-baseline --no-hooks > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "eight" 2>&1 | tail -5 > gpurun_out/r02z_pytest_gpu_8gpus.log
