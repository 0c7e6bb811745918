set -x
for g in "2,1,1" "1,1,2"; do
BENCH_P2P_DEBUG=1 BENCH_GRID=$g python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 10 --no-parity --configs "" > "gpurun_out/r02h_bench_n2_$g.json" 2> "gpurun_out/r02h_bench_n2_$g.err"
done
