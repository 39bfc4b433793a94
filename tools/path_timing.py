"""use_path=True at the bench shape (config 2, B = 64, 160 px): per-step wall time of the per-step Python loop vs the
fused CUDA-graph loop."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from conditional_score_diffusion_b200 import sampling, sde_lib  # noqa: E402
from conditional_score_diffusion_b200.models import ncsnpp  # noqa: E402,F401  (registers the model names)
from conditional_score_diffusion_b200.models import utils as mutils  # noqa: E402
from conditional_score_diffusion_b200.sampling import conditional  # noqa: E402


def main():
    cfg = bench.workload_config()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = mutils.create_model(cfg).to(dev).eval()
    B, img = 64, cfg.data.image_size
    sde = {"x": sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000),
           "y": sde_lib.VESDE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)}
    y = torch.rand(B, 3, img, img, device=dev)
    steps = 10
    for fused_on in (False, True):
        conditional.FUSED_PATH = fused_on
        fn = sampling.get_pc_conditional_sampler(sde, (B, 3, img, img), sampling.get_predictor("conditional_reverse_diffusion"),
                                                 sampling.get_corrector("conditional_langevin"), 0.075, steps, 1,
                                                 continuous=True, denoise=True, use_path=True, eps=1e-5)
        fn(model, y)                      # warm-up (plans, graphs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, _ = fn(model, y)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        print(f"use_path fused={fused_on}: {dt * 1e3:.2f} ms per PC step, finite={bool(torch.isfinite(out).all())}")


if __name__ == "__main__":
    main()
