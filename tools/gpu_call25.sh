set -x
timeout 200 python bench.py --impl torch_eager_gpu --steps 10 > gpurun_out/bench_torch_eager_gpu_r1b.json 2> gpurun_out/bench_torch_eager_gpu_r1b.err; cut -c1-400 gpurun_out/bench_torch_eager_gpu_r1b.json; tail -2 gpurun_out/bench_torch_eager_gpu_r1b.err
