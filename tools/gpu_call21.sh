set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1i.log; cat gpurun_out/pytest_gpu_r1i.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; cut -c1-200 gpurun_out/bench_r1m.json; tail -2 gpurun_out/bench_r1m.err
CSD_NO_DEFER_FINALIZE=1 timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1m_nodefer.json 2> gpurun_out/bench_r1m_nodefer.err; cut -c1-200 gpurun_out/bench_r1m_nodefer.json
