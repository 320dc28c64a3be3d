set -x
timeout 1200 python -m pytest tests/test_decode_stream_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python tools/prof_decode.py fused > gpurun_out/r2c_decode_fused.log 2>&1; cat gpurun_out/r2c_decode_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 8 -c 3 -o gpurun_out/r2c_decode_stream python tools/prof_decode.py fused > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -5 gpurun_out/r2c_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac'): print(k, d.get(k))
print(d['e2e']); print(d.get('kernel_ms_eager_pass'))
P
