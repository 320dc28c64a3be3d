python examples/demo_synthetic.py --out /tmp/ivlm_demo --config tiny --batch 2 2>&1 | tail -4
python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print(d.get('stage_ms'), d.get('decode_hbm')['frac'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
