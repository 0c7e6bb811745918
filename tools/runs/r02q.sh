set -x
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 1000 > gpurun_out/r02q_pka_profile.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 21000 -c 150 --csv --log-file gpurun_out/r02q_pka_launches.csv python tools/pka_cascade.py 100 5000 600 > gpurun_out/r02q_b.log 2>&1
