set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_r1_final4.log; cat gpurun_out/pytest_gpu_r1_final4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1_final5.json 2> gpurun_out/bench_r1_final5.err; cut -c1-200 gpurun_out/bench_r1_final5.json; tail -2 gpurun_out/bench_r1_final5.err
