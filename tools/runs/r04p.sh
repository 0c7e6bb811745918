# round 2, session 2: dilute force epilogue fetches df_j with the position -- A/B on the alloy
set -x
mkdir -p gpurun_out
RATIO=97,2,1 timeout 600 python tools/time_variants.py build/variants/prev_head.so > gpurun_out/r04p_variants_alloy.log 2>&1
