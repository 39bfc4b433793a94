set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1f.log; cat gpurun_out/pytest_gpu_r1f.log
timeout 300 python tools/profile_plan.py > gpurun_out/profile_plan_r1j.txt 2>&1; head -16 gpurun_out/profile_plan_r1j.txt
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; cut -c1-200 gpurun_out/bench_r1j.json; tail -2 gpurun_out/bench_r1j.err
