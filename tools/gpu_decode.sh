set -x
python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -15 > gpurun_out/pytest.log; tail -12 gpurun_out/pytest.log
python tools/prof_decode.py > gpurun_out/decode_ops.log 2>&1; cat gpurun_out/decode_ops.log
