# 2 GPUs: multi-GPU parity tests at HEAD, N=2 bench (parity self-check on the 2x1x1 grid, configs), A/B of the all-reduce in front of rho
set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_frontier.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02d_pytest_gpu_2gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
MISA_B200_OPTS=dmax_flags=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --no-parity --configs "" > gpurun_out/r02d_bench_n2_allreduce.json 2> gpurun_out/r02d_bench_n2_allreduce.err
python tools/time_variants.py > gpurun_out/r02d_variants.log 2>&1
