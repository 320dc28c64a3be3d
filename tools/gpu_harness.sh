python -m pytest tests/test_harness_gpu.py -m gpu -q -x 2>&1 | tail -8
