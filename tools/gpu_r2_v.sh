set -x
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "stage_level" 2>&1 | tail -30
