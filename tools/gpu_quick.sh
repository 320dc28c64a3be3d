set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
python tools/prof_gemm.py 10 > gpurun_out/gemm_shapes.log 2>&1; cat gpurun_out/gemm_shapes.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('roofline'), d.get('kernel_time_shares'), d.get('attention_tflops'))"
