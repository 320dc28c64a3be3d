set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:global64h -c 1 -o gpurun_out/r2g_global64h python tools/prof_attn.py 1 new > gpurun_out/r2g_ncu1.log 2>&1; tail -2 gpurun_out/r2g_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_h -c 1 -o gpurun_out/r2g_window_h python tools/prof_attn.py 1 new > gpurun_out/r2g_ncu2.log 2>&1; tail -2 gpurun_out/r2g_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sam_attn_window_tcgen05 -c 1 -o gpurun_out/r2g_window_old python tools/prof_attn.py 1 new > gpurun_out/r2g_ncu3.log 2>&1; tail -2 gpurun_out/r2g_ncu3.log
