set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_decode_stream_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python tools/prof_decode.py fused > $O/r2r_decode_fused.log 2>&1; cat $O/r2r_decode_fused.log
