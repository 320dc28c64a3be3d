#!/bin/bash
# Reproduces the round-2 evidence under profiles/ on ONE B200 (gpurun -- 'bash tools/gpu_measure.sh [quick]'); everything lands in
# gpurun_out/ as TEXT: each .ncu-rep is summarised on the box (tools/ncu_top.py / ncu_traffic.py / summarize_launches.py) and then
# deleted, because gpurun only brings back 64 MiB.  Copy the r2_*.txt / .json / .md files into profiles/.
#   1 GPU tests  2 bench lines (ours + reference arm)  3 launch list of the bench command  4 ncu --set full captures
set -x
O=gpurun_out
if [ "$1" != "ncu" ]; then
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_pytest_gpu.log; cat $O/r2_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_default_1gpu.json 2> $O/r2_bench_default_1gpu.err; tail -c 400 $O/r2_bench_default_1gpu.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; tail -c 300 $O/r2_bench_reference.json
timeout 600 python tools/prof_decode.py fused > $O/r2_decode_layer_ops_in_graph.txt 2>&1
timeout 600 python tools/prof_decode.py prefetch >> $O/r2_decode_layer_ops_in_graph.txt 2>&1
timeout 600 python tools/prof_tail.py 8 30 > $O/r2_decoder_tail.txt 2>&1
timeout 600 python tools/prof_attn.py 5 new > $O/r2_attn_shapes.txt 2>&1
fi
[ "$1" == "quick" ] && exit 0
# launch list (cold-cache, serialised: shares only)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-view-cache-pass > $O/r2_launches_bench.log 2>&1
python tools/summarize_launches.py $O/r2_launches.csv > $O/r2_launches_bench.md; rm -f $O/r2_launches.csv; head -12 $O/r2_launches_bench.md
# one --set full capture per kernel family of the path; summarised here, report deleted
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, extra ncu flags, command...
    local name=$1 rx=$2 extra=$3; shift 3
    timeout 900 $NCU -k regex:"$rx" $extra -o $O/$name "$@" > $O/$name.log 2>&1
    if [ -f $O/$name.ncu-rep ]; then python tools/ncu_top.py $O/$name.ncu-rep 25 > $O/$name.txt 2>&1; fi
}
cap r2_decode_stream decode_stream "-s 8 -c 4" python tools/prof_decode.py fused
cap r2_decode_attn decode_attn_paged "-s 4 -c 1" python tools/prof_decode.py fused
cap r2_lift lift_warp "-c 6" python tools/prof_tail.py 8 2
cap r2_bilinear bilinear "-c 1" python tools/prof_tail.py 8 2
cap r2_attn_small "attn_few|upscale_hyper" "-s 6 -c 3" python tools/prof_tail.py 8 2
cap r2_attn_global64h global64h "-c 1" python tools/prof_attn.py 1 new
cap r2_attn_window sam_attn_window_tcgen05 "-c 1" python tools/prof_attn.py 1 new
timeout 600 $NCU -k regex:gemm_bf16_tcgen05 -o $O/r2_gemm_traffic python tools/prof_gemm.py 1 > $O/r2_gemm_traffic.log 2>&1
python tools/ncu_traffic.py $O/r2_gemm_traffic.ncu-rep $O/r2_traffic.json > /dev/null 2>&1
python tools/ncu_top.py $O/r2_gemm_traffic.ncu-rep 25 > $O/r2_gemm_shipping.txt 2>&1
# small decoder-tail and norm kernels from a full-size step (13B weights: about two minutes under ncu)
cap r2_small_kernels "layernorm_reg|rmsnorm_kernel|cam_gate|fill_rows|decode_prepare|decode_finish|argmax" "-c 10" \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-view-cache-pass
rm -f $O/*.ncu-rep
ls -la $O | tail -30
