# round 2, session 2: two-GPU evidence of the current tree: multi-GPU parity tests, the N = 2 bench line (alloy + cells200 configs in it)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r04d_pytest_gpu_2gpus.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r04d_bench_n2.json 2> gpurun_out/r04d_bench_n2.err
