set -x
timeout 1500 python -m pytest tests/test_raster_gpu.py tests/test_optim_gpu.py tests/test_model_gpu.py tests/test_harness_gpu.py -m gpu -q -x 2>&1 | tail -12
timeout 900 python -m pytest tests/test_bench_gpu.py -m gpu -q -x -k "joint_fit or one_gpu" 2>&1 | tail -8
IVLM_FIT_ITERS=250 timeout 900 python bench.py --workload joint_fit --steps 2 --warmup 1 --batch 8 > gpurun_out/r2i_bench_joint.json 2> gpurun_out/r2i_bench_joint.err; tail -3 gpurun_out/r2i_bench_joint.err; tail -c 900 gpurun_out/r2i_bench_joint.json
