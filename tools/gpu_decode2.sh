python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -2
python tools/prof_decode.py 2>&1 | head -6
