set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; cut -c1-300 gpurun_out/bench_r1b.json; tail -3 gpurun_out/bench_r1b.err
