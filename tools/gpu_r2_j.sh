set -x
timeout 900 python -m pytest tests/test_lift_gpu.py tests/test_raster_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/prof_tail.py 8 > gpurun_out/r2j_tail.log 2>&1; cat gpurun_out/r2j_tail.log
timeout 300 python tools/prof_tail.py 1 > gpurun_out/r2j_tail1.log 2>&1; cat gpurun_out/r2j_tail1.log
timeout 600 python tools/prof_fit.py > gpurun_out/r2j_fit.log 2>&1; cat gpurun_out/r2j_fit.log
