set -x
python -m pytest tests -m gpu -q -rA 2>&1 | tail -80 > gpurun_out/pytest.log; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest.log | tail -15
python tools/prof_gemm.py 10 > gpurun_out/gemm_shapes.log 2>&1; cat gpurun_out/gemm_shapes.log
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 4 -f -o gpurun_out/prof_gemm_r1 python tools/prof_gemm.py 1 > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 21000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/bench_ncu.log 2>&1; tail -2 gpurun_out/bench_ncu.log | cut -c1-300
ls -la gpurun_out
