# round 2 (late): compute-sanitizer over every kernel family on small boxes (tools/sanitize_target.py)
set -x
python tools/sanitize_target.py > gpurun_out/r03c_plain.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/r03c_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r03c_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 7 python tools/sanitize_target.py thermal alloy cascade > gpurun_out/r03c_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r03c_racecheck.log
timeout 400 compute-sanitizer --tool synccheck --print-limit 30 --error-exitcode 7 python tools/sanitize_target.py thermal alloy cascade world > gpurun_out/r03c_synccheck.log 2>&1; echo "synccheck exit $?" >> gpurun_out/r03c_synccheck.log
tail -5 gpurun_out/r03c_plain.log gpurun_out/r03c_memcheck.log gpurun_out/r03c_racecheck.log gpurun_out/r03c_synccheck.log
