set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01w_launches.csv python bench.py --steps 3 --warmup 3 --equil 20 --no-cpu-baseline > gpurun_out/r01w_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(force|rho)_[ab]' --launch-skip 808 --launch-count 4 -o gpurun_out/r01w_full python tools/ncu_target.py 100 200 3 > gpurun_out/r01w_ncu.log 2>&1
