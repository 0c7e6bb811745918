set -x
PKA_PROFILE=1 timeout 600 python tools/pka_cascade.py 100 5000 1000 > gpurun_out/r02q_pka_profile.log 2>&1
