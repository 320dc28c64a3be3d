set -x
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/prof_gemm_qkv_v2 python tools/prof_gemm.py 1 > gpurun_out/ncu_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/prof_gemm_mlp1_v2 python tools/prof_gemm.py 1 >> gpurun_out/ncu_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sam_attn_tcgen05 -s 2 -c 1 -f -o gpurun_out/prof_attn_global python tools/prof_attn.py 1 new >> gpurun_out/ncu_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sam_attn_window -s 2 -c 1 -f -o gpurun_out/prof_attn_window python tools/prof_attn.py 1 new >> gpurun_out/ncu_k.log 2>&1
tail -3 gpurun_out/ncu_k.log
