set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_decode_stream_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 300 python tools/prof_decode.py chain2 > $O/r2t_chain2.log 2>&1; cat $O/r2t_chain2.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -8
