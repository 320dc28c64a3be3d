set -x
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_decode_stream_gpu.py -m gpu -q -x 2>&1 | tail -10
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -5 gpurun_out/r2k_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','stage_ms','decode_hbm','path_tensor_frac','sam_attention_tflops'): print(k, d.get(k))
print(d['e2e']); print(d.get('kernel_ms_eager_pass')); print(d.get('roofline'))
P
timeout 600 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu-baseline --no-profile --no-view-cache-pass > gpurun_out/r2k_bench_b1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_b1.json').read().strip().splitlines()[-1]); print('batch1', d['value'], d['ms_per_step'], d['stage_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -o gpurun_out/r2_gemm_traffic python tools/prof_gemm.py 1 > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log
timeout 600 python tools/prof_decode.py fused > gpurun_out/r2k_decode_fused.log 2>&1; cat gpurun_out/r2k_decode_fused.log
