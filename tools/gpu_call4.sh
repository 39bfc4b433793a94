set -x
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -x -q -k "transposed or fused_groupnorm" -s 2>&1 | grep -v "^$" | tail -40
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/profile_plan.py 64 > gpurun_out/profile_plan_persist.txt 2>&1; head -36 gpurun_out/profile_plan_persist.txt
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_persist.json 2> gpurun_out/bench_persist.err; cut -c1-400 gpurun_out/bench_persist.json; tail -3 gpurun_out/bench_persist.err
