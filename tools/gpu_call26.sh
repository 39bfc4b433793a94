set -x
timeout 100 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_network.py tests/test_gpu_sampling.py tests/test_gpu_ddpm.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_tail_skip.log; cat gpurun_out/pytest_tail_skip.log
timeout 100 python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r1_tail.json 2> gpurun_out/bench_r1_tail.err; cut -c1-200 gpurun_out/bench_r1_tail.json; tail -2 gpurun_out/bench_r1_tail.err
