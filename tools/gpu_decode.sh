set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest.log; tail -8 gpurun_out/pytest.log
python tools/prof_decode.py > gpurun_out/decode_ops.log 2>&1; tail -5 gpurun_out/decode_ops.log
bash tools/gpu_bench.sh
