set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1g.log; cat gpurun_out/pytest_gpu_r1g.log
timeout 300 python tools/tap_nodata_probe.py --quick > gpurun_out/tap_probe_epi.txt 2>&1; cat gpurun_out/tap_probe_epi.txt
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; cut -c1-200 gpurun_out/bench_r1k.json; tail -2 gpurun_out/bench_r1k.err
timeout 300 python bench.py --workload train --steps 20 > gpurun_out/bench_train_r1k.json 2> gpurun_out/bench_train_r1k.err; cut -c1-260 gpurun_out/bench_train_r1k.json
