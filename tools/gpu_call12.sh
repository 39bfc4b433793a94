set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r1_final.log; cat gpurun_out/pytest_gpu_r1_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --workload train --steps 20 > gpurun_out/bench_train_r1.json 2> gpurun_out/bench_train_r1.err; cut -c1-260 gpurun_out/bench_train_r1.json
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/bench_r1_final3.json 2> gpurun_out/bench_r1_final3.err; cut -c1-260 gpurun_out/bench_r1_final3.json; tail -2 gpurun_out/bench_r1_final3.err
timeout 300 python tools/train_step_profile.py 50 64 fused > gpurun_out/train_profile_r1_final.txt 2>&1; head -3 gpurun_out/train_profile_r1_final.txt
