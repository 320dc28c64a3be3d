ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/prof_gemm_mlp1_v3 python tools/prof_gemm.py 1 > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f -o gpurun_out/prof_gemm_proj_v3 python tools/prof_gemm.py 1 >> gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log
