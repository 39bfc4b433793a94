set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r1.log; tail -5 gpurun_out/pytest_gpu_r1.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
timeout 300 python tools/profile_plan.py 64 > gpurun_out/profile_plan_r1.txt 2>&1; head -60 gpurun_out/profile_plan_r1.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/forward_launches_r1.csv python tools/forward_ncu_target.py 64 > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 6 -c 3 -o gpurun_out/conv_full_r1 python tools/conv_ncu_target.py 64 > gpurun_out/ncu2.log 2>&1; tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
