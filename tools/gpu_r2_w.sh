set -x
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_full_depth_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-view-cache-pass > gpurun_out/r2w_bench.json 2>gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac','sam_attention_tflops'): print(k, d.get(k))
print(d['e2e']); print(d.get('kernel_ms_eager_pass'))
P
