set -x
timeout 1500 python -m pytest tests -m gpu -q -x --ignore=tests/test_full_depth_gpu.py --ignore=tests/test_bench_gpu.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -x -s 2>&1 | grep -v "residual stream" | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -5 gpurun_out/r2e_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac','sam_attention_tflops'): print(k, d.get(k))
print(d['e2e']); print(d.get('kernel_ms_eager_pass')); print(d.get('roofline'))
P
