set -x
timeout 600 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload train --steps 20 > gpurun_out/bench_train_r1n.json 2> gpurun_out/bench_train_r1n.err; cut -c1-260 gpurun_out/bench_train_r1n.json; tail -2 gpurun_out/bench_train_r1n.err
