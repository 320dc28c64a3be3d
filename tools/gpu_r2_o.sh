set -x
timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_harness_gpu.py tests/test_optim_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/prof_decode.py fused > gpurun_out/r2o_decode_fused.log 2>&1; cat gpurun_out/r2o_decode_fused.log
