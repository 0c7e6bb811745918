set -x
timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r02v_variants.log 2>&1
timeout 600 python tools/time_variants.py build/variants/prev_head.so >> gpurun_out/r02v_variants.log 2>&1
