set -x
MISA_B200_BACKTRACE=1 timeout 300 python tools/pka_cascade.py 100 5000 800 > gpurun_out/r02r_pka.log 2>&1
