set -x
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_ddpm.py -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/train_launches_r1b.csv python tools/train_ncu_target.py 50 > gpurun_out/ncu_train_b.log 2>&1; tail -2 gpurun_out/ncu_train_b.log
