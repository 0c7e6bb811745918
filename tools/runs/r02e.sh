set -x
for o in "" "late=0" "late=0,dmax_flags=0" "p2p=0"; do
MISA_B200_OPTS=$o python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 --no-parity --configs "" > "gpurun_out/r02e_bench_n2_$o.json" 2> "gpurun_out/r02e_bench_n2_$o.err"
done
python bench.py --steps 100 --warmup 10 --no-parity --configs "" --no-cpu-baseline --no-hooks > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
