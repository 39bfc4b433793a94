"""ncu target: ONE predictor-corrector step of the bench workload (2 network evaluations + update kernels, B from
argv, default 64), run eagerly inside cudaProfilerStart/Stop after two warm-up steps.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/step_ncu_target.py
"""
import sys

import torch

sys.path.insert(0, ".")
import bench
from conditional_score_diffusion_b200 import sde_lib
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
from conditional_score_diffusion_b200.sampling.fused import FusedPCSampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = bench.workload_config()
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = utils.create_model(cfg).to(dev).eval()
sde = {"x": sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000),
       "y": sde_lib.VESDE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)}
shape = (B, 3, bench.IMAGE, bench.IMAGE)
fs = FusedPCSampler(model, sde, shape, "reverse_diffusion", "langevin", bench.SNR, bench.PC_STEPS, 1, False, True, True,
                    bench.EPS, conditional=True)
fs._setup(dev)
fs.plan.use_graph = False
fs.y.copy_(torch.rand(*shape, device=dev))
fs.x.copy_(torch.randn(*shape, device=dev) * cfg.model.sigma_max_x)
fs.draw_noise = True
fs.step_idx.zero_()
for _ in range(2):
    fs._step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fs._step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
