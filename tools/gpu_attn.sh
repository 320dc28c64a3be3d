set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "sam_attention_tcgen05" 2>&1 | tail -25 > gpurun_out/pytest_attn.log; tail -25 gpurun_out/pytest_attn.log
timeout 120 python tools/prof_attn.py 5 > gpurun_out/attn_shapes.log 2>&1; cat gpurun_out/attn_shapes.log
