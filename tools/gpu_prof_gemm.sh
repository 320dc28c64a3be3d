set -x
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/prof_gemm_qkv python tools/prof_gemm.py 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/prof_gemm_mlp1 python tools/prof_gemm.py 1 >> gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
