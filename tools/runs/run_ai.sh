set -x
python tools/time_variants.py build/variants/*.so > gpurun_out/r01ai_variants.log 2>&1
