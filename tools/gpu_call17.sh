set -x
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_sampling.py -m gpu -x -q 2>&1 | tail -5
CSD_HEAD_MODE=2 timeout 300 python tools/profile_plan.py > gpurun_out/profile_plan_head2.txt 2>&1; head -8 gpurun_out/profile_plan_head2.txt; grep "'conv', [0-9]*, [0-9]*, 6, " gpurun_out/profile_plan_head2.txt
CSD_HEAD_MODE=2 timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_head2.json 2> gpurun_out/bench_head2.err; cut -c1-200 gpurun_out/bench_head2.json; tail -2 gpurun_out/bench_head2.err
CSD_HEAD_MODE=0 timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_head0.json 2> gpurun_out/bench_head0.err; cut -c1-200 gpurun_out/bench_head0.json; tail -2 gpurun_out/bench_head0.err
