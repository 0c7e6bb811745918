set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r01am_bench_n8.json 2> gpurun_out/r01am_bench_n8.err
MISA_B200_OPTS=late=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r01am_bench_n8_front.json 2> gpurun_out/r01am_bench_n8_front.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 100 --warmup 10 --ratio 97 2 1 > gpurun_out/r01am_bench_n8_alloy.json 2> gpurun_out/r01am_bench_n8_alloy.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r01am_bench_n4.json 2> gpurun_out/r01am_bench_n4.err
