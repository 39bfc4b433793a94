set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_r1d.log; cat gpurun_out/pytest_gpu_r1d.log
timeout 300 python tools/tap_nodata_probe.py > gpurun_out/tap_nodata_probe.txt 2>&1; cat gpurun_out/tap_nodata_probe.txt
timeout 300 python tools/profile_plan.py > gpurun_out/profile_plan_r1h.txt 2>&1; head -16 gpurun_out/profile_plan_r1h.txt
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; cut -c1-200 gpurun_out/bench_r1h.json; tail -2 gpurun_out/bench_r1h.err
