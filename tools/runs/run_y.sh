set -x
python tools/time_sym.py 100 200 > gpurun_out/r01y_time_sym.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(force|rho)_[ab]' --launch-skip 808 --launch-count 4 -o gpurun_out/r01y_full python tools/ncu_target.py 100 200 3 > gpurun_out/r01y_ncu.log 2>&1
