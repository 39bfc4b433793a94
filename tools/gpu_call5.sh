set -x
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -x -q -k "transposed or fused_groupnorm" -s 2>&1 | grep -v "^$" | tail -40
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/profile_plan.py 64 > gpurun_out/profile_plan_persist2.txt 2>&1; head -36 gpurun_out/profile_plan_persist2.txt
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_persist2.json 2> gpurun_out/bench_persist2.err; cut -c1-400 gpurun_out/bench_persist2.json; tail -3 gpurun_out/bench_persist2.err
CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS python -m conditional_score_diffusion_b200.build --force > /dev/null && timeout 300 python tools/conv_phase_timing.py 2>&1 | tee gpurun_out/conv_phase_timing_persist2.txt
