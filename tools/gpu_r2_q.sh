set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_ops_gpu.py tests/test_decode_stream_gpu.py tests/test_model_gpu.py tests/test_lift_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python tools/prof_decode.py prefetch > $O/r2q_decode_prefetch.log 2>&1; cat $O/r2q_decode_prefetch.log
timeout 600 python tools/prof_tail.py 8 30 > $O/r2q_tail.log 2>&1; cat $O/r2q_tail.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2q_bench.json 2> $O/r2q_bench.err; tail -5 $O/r2q_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac','sam_attention_tflops'): print(k, d.get(k))
print(d['e2e']); print(d.get('kernel_ms_eager_pass'))
P
