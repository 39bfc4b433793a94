set -x
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -x -q -k "fused_groupnorm or gn_coeffs or transposed_halo or transposed_segments" -s 2>&1 | tail -30
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/profile_plan.py 64 > gpurun_out/profile_plan_fused.txt 2>&1; head -40 gpurun_out/profile_plan_fused.txt
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; cut -c1-400 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
