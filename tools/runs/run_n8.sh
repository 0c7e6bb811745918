set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k eight 2>&1 | tail -10 > gpurun_out/r01ad_pytest_n8.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r01ad_bench_n8.json 2> gpurun_out/r01ad_bench_n8.err
MISA_B200_OPTS=p2p=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r01ad_bench_n8_nccl.json 2> gpurun_out/r01ad_bench_n8_nccl.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 100 --warmup 10 --ratio 97 2 1 > gpurun_out/r01ad_bench_n8_alloy.json 2> gpurun_out/r01ad_bench_n8_alloy.err
