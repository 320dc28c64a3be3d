ncu --set full --clock-control none --import-source on -k regex:decode_attn_paged -s 3 -c 1 -f -o gpurun_out/prof_decode_attn python tools/prof_decode.py > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rope_kv_store -s 3 -c 1 -f -o gpurun_out/prof_rope python tools/prof_decode.py >> gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
