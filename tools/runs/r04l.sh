# round 2, session 2: rho kernel -- eight near pairs per trip, threads per CTA, after the LSU trims
set -x
mkdir -p gpurun_out
timeout 900 python tools/time_variants.py build/variants/*.so > gpurun_out/r04l_variants.log 2>&1
