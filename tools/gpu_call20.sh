set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1h.log; cat gpurun_out/pytest_gpu_r1h.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; cut -c1-200 gpurun_out/bench_r1l.json; tail -2 gpurun_out/bench_r1l.err
CSD_NO_DEFER_FINALIZE=1 timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1l_nodefer.json 2> gpurun_out/bench_r1l_nodefer.err; cut -c1-200 gpurun_out/bench_r1l_nodefer.json
