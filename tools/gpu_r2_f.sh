set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "sam_attention" 2>&1 | tail -8
timeout 300 python tools/prof_attn.py 10 new > gpurun_out/r2f_attn.log 2>&1; cat gpurun_out/r2f_attn.log
