set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q -rA 2>&1 | tail -70 > gpurun_out/pytest.log; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -30 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_first.log 2>&1; tail -5 gpurun_out/bench_first.log
