#!/bin/bash
# Reproduces the round-2 evidence under profiles/ on ONE B200 (gpurun -- 'bash tools/gpu_measure.sh'); raw files land in
# gpurun_out/ and are summarised in the build container with tools/ncu_top.py / summarize_launches.py / ncu_traffic.py.
#   1 GPU tests  2 bench lines (ours + reference arm)  3 launch list of the bench command  4 ncu --set full captures
set -x
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_pytest_gpu.log; cat $O/r2_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_default_1gpu.json 2> $O/r2_bench_default_1gpu.err; tail -c 400 $O/r2_bench_default_1gpu.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; tail -c 300 $O/r2_bench_reference.json
# launch list (cold-cache, serialised: shares only)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-view-cache-pass > $O/r2_launches_bench.log 2>&1
# one --set full capture per kernel family of the path
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:decode_stream -s 8 -c 4 -o $O/r2_decode_stream python tools/prof_decode.py fused > $O/r2_ncu_decode.log 2>&1
timeout 600 $NCU -k regex:decode_attn_paged -s 4 -c 1 -o $O/r2_decode_attn python tools/prof_decode.py fused > /dev/null 2>&1
timeout 600 $NCU -k regex:lift_warp -c 6 -o $O/r2_lift python tools/prof_tail.py 8 2 > /dev/null 2>&1
timeout 600 $NCU -k regex:bilinear -c 1 -o $O/r2_bilinear python tools/prof_tail.py 8 2 > /dev/null 2>&1
timeout 600 $NCU -k regex:global64h -c 1 -o $O/r2_attn_global64h python tools/prof_attn.py 1 new > /dev/null 2>&1
timeout 600 $NCU -k regex:sam_attn_window_tcgen05 -c 1 -o $O/r2_attn_window python tools/prof_attn.py 1 new > /dev/null 2>&1
timeout 600 $NCU -k regex:gemm_bf16_tcgen05 -o $O/r2_gemm_traffic python tools/prof_gemm.py 1 > /dev/null 2>&1
# small decoder-tail and norm kernels: first instance of each inside one tiny-config evaluate() at real widths is not available,
# so they are captured from a full-size step (13B weights: about two minutes under ncu)
timeout 1500 $NCU -k regex:"attn_few_queries|attn_few_keys|upscale_hyper_dot|layernorm_reg|rmsnorm_kernel|cam_gate|fill_rows" -c 14 \
    -o $O/r2_small_kernels python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-view-cache-pass > $O/r2_ncu_small.log 2>&1
ls -la $O/*.ncu-rep
