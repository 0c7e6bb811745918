set -x
mkdir -p gpurun_out
timeout 900 python tools/time_variants.py build/variants/*.so > gpurun_out/r04n_variants.log 2>&1
