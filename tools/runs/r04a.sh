# round 2, session 2: neighbour fields through ld.global.nc instead of TLD (EAM_NB_LDG bit mask) -- A/B on one box, and the TEX
# front-end counters of the default kernels (is the texture unit's quad rate the wall of the rho kernel?)
set -x
mkdir -p gpurun_out
timeout 1200 python tools/time_variants.py build/variants/*.so > gpurun_out/r04a_variants.log 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_tex.sum,l1tex__texin_requests_mem_texture.sum,l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__texin_sm2tex_req_cycles_active.sum,l1tex__texin_sm2tex_req_cycles_stalled.sum,l1tex__f_wavefronts.sum,l1tex__f_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__f_cycles_active.sum,l1tex__f_tex2sm_cycles_active.sum,l1tex__f_tex2sm_cycles_stalled.sum,l1tex__data_pipe_tex_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
timeout 600 ncu --clock-control none -k regex:'k_(force|rho)_f' --launch-skip 402 --launch-count 2 --metrics $M --csv --log-file gpurun_out/r04a_tex_metrics.csv python tools/ncu_target.py 100 200 3 > gpurun_out/r04a_ncu.log 2>&1
MISA_B200_LIB=$PWD/build/variants/ldg_all.so timeout 600 ncu --clock-control none -k regex:'k_(force|rho)_f' --launch-skip 402 --launch-count 2 --metrics $M --csv --log-file gpurun_out/r04a_tex_metrics_ldg_all.csv python tools/ncu_target.py 100 200 3 > gpurun_out/r04a_ncu2.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r04a_smi.log
