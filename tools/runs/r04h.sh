# round 2, session 2: GPU tests after the removal of the PHI_TEX / staged-minority experiments
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r04h_pytest_gpu.log
