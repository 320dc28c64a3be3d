python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -3
python tools/prof_decode.py chain 2>&1 | tail -5
python tools/prof_decode.py 2>&1 | tail -12
run() { echo "== $*"; timeout 600 python bench.py --no-cpu-baseline --no-profile --steps 4 --warmup 3 "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(round(d['value'],2), 'img/s', round(d['ms_per_step'],1), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['stage_ms'], d['decode_hbm'])"; }
run --pdl 0
run --pdl 1
