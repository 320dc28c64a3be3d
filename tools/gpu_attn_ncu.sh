python tools/prof_attn.py 5 new 2>&1 | tee gpurun_out/attn_shapes_s3.log
ncu --set full --clock-control none --import-source on -k regex:sam_attn_global64 -s 2 -c 1 -f -o gpurun_out/prof_attn_global64 python tools/prof_attn.py 1 new > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sam_attn_window -s 2 -c 1 -f -o gpurun_out/prof_attn_window_s3 python tools/prof_attn.py 1 new >> gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
