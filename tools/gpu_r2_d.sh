set -x
timeout 1200 python -m pytest tests/test_decode_stream_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python tools/prof_decode.py fused > gpurun_out/r2d_decode_fused.log 2>&1; cat gpurun_out/r2d_decode_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 8 -c 4 -o gpurun_out/r2d_decode_stream python tools/prof_decode.py fused > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log
