set -x
timeout 1200 python -m pytest tests/test_decode_stream_gpu.py tests/test_lift_gpu.py tests/test_raster_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2b_pytest.log; tail -25 gpurun_out/r2b_pytest.log
timeout 300 python tools/prof_tail.py 8 > gpurun_out/r2b_tail.log 2>&1; cat gpurun_out/r2b_tail.log
timeout 600 python tools/prof_decode.py fused > gpurun_out/r2b_decode_fused.log 2>&1; cat gpurun_out/r2b_decode_fused.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -5 gpurun_out/r2b_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac'): print(k, d.get(k))
print(d['e2e']); print(d.get('kernel_ms_eager_pass'))
P
timeout 600 python -m pytest tests/test_bench_gpu.py -m gpu -q -x 2>&1 | tail -5
