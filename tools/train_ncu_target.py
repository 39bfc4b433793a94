"""ncu target: ONE training step (forward + backward launch lists, eager) of BASELINE configs[3] (ddpm_paired_SR3 nf128,
64 px, batch 50), inside cudaProfilerStart/Stop after two warm-up steps.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_ncu_target.py
"""
import math
import os
import sys

import torch

os.environ["CSD_NO_GRAPH"] = "1"          # eager launch lists: ncu sees every kernel of the step
sys.path.insert(0, ".")
import bench
from conditional_score_diffusion_b200 import losses, sde_lib
from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.TRAIN_BATCH
cfg = bench.train_config()
torch.manual_seed(0)
model = utils.create_model(cfg).cuda().train()
sde = sde_lib.cVESDE(5e-3, math.sqrt(3 * bench.TRAIN_IMAGE ** 2), 1000)
loss_fn = losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                         likelihood_weighting=True)
x = torch.rand(B, 3, bench.TRAIN_IMAGE, bench.TRAIN_IMAGE, device="cuda")
y = torch.rand_like(x)
for _ in range(2):
    model.zero_grad()
    loss_fn(model, (y, x)).backward()
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.zero_grad()
loss_fn(model, (y, x)).backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
