set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, extra ncu flags, command...
    local name=$1 rx=$2 extra=$3; shift 3
    timeout 900 $NCU -k regex:"$rx" $extra -o $O/$name "$@" > $O/$name.log 2>&1
    if [ -f $O/$name.ncu-rep ]; then python tools/ncu_top.py $O/$name.ncu-rep 40 > $O/$name.txt 2>&1; fi
}
cap r2s_decode_attn decode_attn_paged "-s 4 -c 1" python tools/prof_decode.py fused
cap r2s_decode_stream decode_stream "-s 8 -c 4" python tools/prof_decode.py fused
rm -f $O/*.ncu-rep
cat $O/r2s_decode_attn.txt
