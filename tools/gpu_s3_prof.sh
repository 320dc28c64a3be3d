python tools/prof_raster.py 2>&1 | tee gpurun_out/prof_raster.txt
python tools/prof_decode.py 2>&1 | tee gpurun_out/prof_decode_s3.txt
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 1 -c 1 -f -o gpurun_out/prof_dec_gemm_qkv python tools/prof_decode.py > gpurun_out/ncu_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemv_small_m -s 1 -c 1 -f -o gpurun_out/prof_dec_gemv_o python tools/prof_decode.py >> gpurun_out/ncu_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemv_small_m -s 130 -c 1 -f -o gpurun_out/prof_dec_gemv_down python tools/prof_decode.py >> gpurun_out/ncu_s3.log 2>&1
tail -3 gpurun_out/ncu_s3.log
ls -la gpurun_out/*.ncu-rep
