set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))"; free -g | head -2; grep -m1 "model name" /proc/cpuinfo
python -m pytest tests -m gpu -q -x --ignore=tests/test_full_depth_gpu.py 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; tail -6 gpurun_out/r2a_pytest.log
python -m pytest tests/test_full_depth_gpu.py -m gpu -q -x -s > gpurun_out/r2a_full_depth.log 2>&1; tail -30 gpurun_out/r2a_full_depth.log
python tools/prof_tail.py 8 > gpurun_out/r2a_tail.log 2>&1; cat gpurun_out/r2a_tail.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lift_warp -c 6 -o gpurun_out/r2a_lift python tools/prof_tail.py 8 2 > gpurun_out/r2a_lift_ncu.log 2>&1; tail -3 gpurun_out/r2a_lift_ncu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; cat gpurun_out/r2a_bench_ref.json; tail -5 gpurun_out/r2a_bench_ref.err
