set -x
for o in "p2p_fence=0" "p2p_fence=1" "p2p_fence=2"; do
MISA_B200_OPTS=$o BENCH_P2P_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 10 --no-parity --configs "" > "gpurun_out/r02i_bench_n2_$o.json" 2> "gpurun_out/r02i_bench_n2_$o.err"
done
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02i_pytest_gpu_2gpus.log
