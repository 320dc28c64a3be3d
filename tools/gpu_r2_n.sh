set -x
timeout 900 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -x -s -k "evaluate" > gpurun_out/r2n_full_depth.log 2>&1; grep -v "residual stream" gpurun_out/r2n_full_depth.log | tail -30
timeout 600 python -m pytest tests/test_harness_gpu.py tests/test_optim_gpu.py -m gpu -q -x -s 2>&1 | tail -8
