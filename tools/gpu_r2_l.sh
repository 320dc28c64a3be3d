set -x
timeout 1800 python -m pytest tests -m gpu -q -x --ignore=tests/test_full_depth_gpu.py --ignore=tests/test_bench_gpu.py 2>&1 | tail -8
timeout 600 python tools/prof_decode.py fused > gpurun_out/r2l_decode_fused.log 2>&1; cat gpurun_out/r2l_decode_fused.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -5 gpurun_out/r2l_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac','sam_attention_tflops'): print(k, d.get(k))
print(d['e2e']); print(d.get('roofline'))
P
