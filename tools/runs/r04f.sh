# round 2, session 2: compute-sanitizer over every kernel family (tools/sanitize_target.py)
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r04f_memcheck.log 2>&1
tail -5 gpurun_out/r04f_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py thermal alloy > gpurun_out/r04f_racecheck.log 2>&1
tail -5 gpurun_out/r04f_racecheck.log
