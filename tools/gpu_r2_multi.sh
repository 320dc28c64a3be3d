set -x
N=${1:-2}
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','gpu_launches','stage_ms','decode_hbm'): print(k, d.get(k))
print(d['e2e']); print(d.get('roofline'))
P
timeout 900 python -m pytest tests/test_bench_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --workload oafford_pc --no-profile > gpurun_out/r2_bench_oafford_n$N.json 2> gpurun_out/r2_bench_oafford_n$N.err; echo rc=$?; tail -c 600 gpurun_out/r2_bench_oafford_n$N.json
