set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1e.log; cat gpurun_out/pytest_gpu_r1e.log
timeout 300 python tools/tap_nodata_probe.py > gpurun_out/tap_chunk_probe.txt 2>&1; cat gpurun_out/tap_chunk_probe.txt
timeout 300 python tools/profile_plan.py > gpurun_out/profile_plan_r1i.txt 2>&1; head -16 gpurun_out/profile_plan_r1i.txt
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; cut -c1-200 gpurun_out/bench_r1i.json; tail -2 gpurun_out/bench_r1i.err
