# overlap sweep: SAM encoder next to the decode steps (one B200)
python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "overlapped or batched" 2>&1 | tail -3
run() { echo "== $*"; timeout 600 python bench.py --no-cpu-baseline --no-profile --steps 4 --warmup 3 "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(round(d['value'],2), 'img/s', round(d['ms_per_step'],1), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['stage_ms'])"; }
run --overlap 0
run --overlap 1 --sm-limit 104
run --overlap 1 --sm-limit 0 --limited-chunks 0
run --overlap 1 --sm-limit 120
run --overlap 1 --sm-limit 88
run --overlap 1 --sm-limit 104 --limited-chunks 4
run --overlap 1 --sm-limit 104 --limited-chunks 6
run --overlap 1 --sm-limit 104 --limited-chunks 8
run --overlap 1 --sm-limit 72 --limited-chunks 8
