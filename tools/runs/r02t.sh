set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 21000 -c 120 --csv --log-file gpurun_out/r02t_pka_launches.csv python tools/pka_cascade.py 100 5000 600 > gpurun_out/r02t_b.log 2>&1
