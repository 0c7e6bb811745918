set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02l_pytest_gpu_2gpus.log
for o in "" "push_fused=0"; do
MISA_B200_OPTS=$o timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 --no-parity --configs "" > "gpurun_out/r02l_bench_n2_$o.json" 2> "gpurun_out/r02l_bench_n2_$o.err"
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-parity --configs "" --no-cpu-baseline --no-hooks > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
