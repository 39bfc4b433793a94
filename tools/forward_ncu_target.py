"""ncu target: ONE forward of the bench network (B from argv, default 64) inside cudaProfilerStart/Stop.

    ncu --profile-from-start off --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/forward_sol.csv python tools/forward_ncu_target.py
"""
import sys

import torch

sys.path.insert(0, ".")
import bench
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = bench.workload_config()
torch.manual_seed(0)
model = utils.create_model(cfg).cuda().eval()
model._engine.ensure_packed(torch.device("cuda", 0))
plan = model._engine.plan(B, 160, 160, 3, 3)
plan.in0.normal_(); plan.in1.uniform_(); plan.labels.fill_(500.0)
for _ in range(2):
    plan.rec.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.rec.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
