set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/profile_plan.py 64 > gpurun_out/profile_plan_r1c.txt 2>&1; head -30 gpurun_out/profile_plan_r1c.txt
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; cut -c1-300 gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_r1c.err
