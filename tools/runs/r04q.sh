# round 2, session 2: four-GPU bench line on the final kernels (headline configuration only)
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 100 --warmup 10 --configs "alloy" > gpurun_out/r04q_bench_n4.json 2> gpurun_out/r04q_bench_n4.err
