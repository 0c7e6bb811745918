set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r01al_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r01al_bench_n2.json 2> gpurun_out/r01al_bench_n2.err
MISA_B200_OPTS=late=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r01al_bench_n2_front.json 2> gpurun_out/r01al_bench_n2_front.err
