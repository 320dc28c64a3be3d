set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_default_s3.log 2>&1; tail -1 gpurun_out/bench_default_s3.log > gpurun_out/bench_default_s3.json; cut -c1-700 gpurun_out/bench_default_s3.json
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_s3.log 2>&1; tail -4 gpurun_out/bench_ref_s3.log | cut -c1-900
run() { echo "== $*"; timeout 600 python bench.py --no-cpu-baseline --no-profile --steps 3 --warmup 3 "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'args': '$*', 'images_per_s': round(d['value'],2), 'ms_per_step': round(d['ms_per_step'],1), 'e2e': round(d['e2e']['value'],2), 'sm_mhz': d['clocks']['sm_mhz'], 'stage_ms': d['stage_ms'], 'decode_hbm_frac': round(d['decode_hbm']['frac'],3)}))"; }
( run --sam-chunk 16; run --sam-chunk 32; run --batch 1; run --batch 4; run --batch 16; run --batch 64; run --batch 256 ) 2>&1 | tee gpurun_out/sweep_s3.txt
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_bf16|gemv_small|sam_attn|flash_attn|decode_|layernorm|rmsnorm|silu_mul|add_bcast|rope_kv|im2col|embed_|gather_rows|argmax|bilinear|cam_gate|upscale_hyper|lift_kernel|attn_few|csr_spmv|sigmoid_where|preprocess_u8|sam_relpos|finalize|cast_" -c 12500 --csv --log-file gpurun_out/launches_r1_s3.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/bench_ncu_s3.log 2>&1; tail -1 gpurun_out/bench_ncu_s3.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:raster_tile -s 4 -c 1 -f -o gpurun_out/prof_raster_tile python tools/prof_raster.py > gpurun_out/ncu_raster.log 2>&1; tail -2 gpurun_out/ncu_raster.log
ls -la gpurun_out/
