set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r1c.log; cat gpurun_out/pytest_gpu_r1c.log
timeout 300 python tools/train_step_profile.py 50 64 fused > gpurun_out/train_profile_r1_final.txt 2>&1; head -7 gpurun_out/train_profile_r1_final.txt
timeout 300 python bench.py --workload train --steps 20 > gpurun_out/bench_train_r1.json 2> gpurun_out/bench_train_r1.err; cut -c1-300 gpurun_out/bench_train_r1.json
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; cut -c1-300 gpurun_out/bench_r1_final2.json; tail -2 gpurun_out/bench_r1_final2.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/train_launches_r1.csv python tools/train_ncu_target.py 50 > gpurun_out/ncu_train.log 2>&1; tail -1 gpurun_out/ncu_train.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wgrad_direct|gn_bwd_apply" -s 20 -c 6 -o gpurun_out/train_kernels_full_r1 -f python tools/train_ncu_target.py 50 > gpurun_out/ncu_train_full.log 2>&1; tail -1 gpurun_out/ncu_train_full.log
