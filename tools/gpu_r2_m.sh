set -x
timeout 1800 python -m pytest tests -m gpu -q -x --ignore=tests/test_full_depth_gpu.py --ignore=tests/test_bench_gpu.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -x -s 2>&1 | grep -v "residual stream" | tail -6
IVLM_FIT_ITERS=250 timeout 900 python bench.py --workload joint_fit --steps 2 --warmup 1 --batch 8 > gpurun_out/r2m_bench_joint.json 2> gpurun_out/r2m_bench_joint.err; tail -3 gpurun_out/r2m_bench_joint.err; python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench_joint.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','fit_ms_per_sample','fit_ms_per_iteration')}, d['e2e'])"
