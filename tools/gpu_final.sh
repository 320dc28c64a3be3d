set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-3000
