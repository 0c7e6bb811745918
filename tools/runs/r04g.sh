# round 2, session 2: eight-GPU evidence of the tree: the N = 8 bench line (parity self-check, alloy + cells200 configs) and the 8-GPU oracle test
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r04g_bench_n8.json 2> gpurun_out/r04g_bench_n8.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "eight" 2>&1 | tail -5 > gpurun_out/r04g_pytest_gpu_8gpus.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-parity --configs "alloy" --no-cpu-baseline --no-hooks > gpurun_out/r04g_bench_n1_same_box.json 2> gpurun_out/r04g_bench_n1.err
