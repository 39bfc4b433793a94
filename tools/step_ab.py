"""Time one PC step of the bench workload (config 2, B = 64, 160 px) under the CURRENT environment switches:
K graph replays between CUDA events. For A/B pairs run in one gpurun call, e.g.

    CSD_FIR_STAGES=1 CSD_FIR_DOWN_QUADS=1 python tools/step_ab.py before
    python tools/step_ab.py after

Prints one line: label, ms per PC step, SM clock seen by nvidia-smi during the run."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from conditional_score_diffusion_b200 import sde_lib  # noqa: E402
from conditional_score_diffusion_b200.models import ncsnpp  # noqa: E402,F401  (registers the model names)
from conditional_score_diffusion_b200.models import utils as mutils  # noqa: E402
from conditional_score_diffusion_b200.sampling.fused import FusedPCSampler  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "run"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    cfg = bench.workload_config()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = mutils.create_model(cfg).to(dev).eval()
    if precision != "bf16":
        model.set_precision(precision)
    B, img = bench.BATCH_PER_GPU, bench.IMAGE
    sde = {"x": sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000),
           "y": sde_lib.VESDE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)}
    shape = (B, 3, img, img)
    fs = FusedPCSampler(model, sde, shape, "reverse_diffusion", "langevin", bench.SNR, 1000, 1, False, True, True,
                        bench.EPS, conditional=True)
    fs._setup(dev)
    fs.y.copy_(torch.rand(*shape, device=dev))
    graph = fs._graph(draw_noise=True)
    fs.x.copy_(torch.randn(*shape, device=dev) * cfg.model.sigma_max_x)
    fs.step_idx.zero_()
    for _ in range(5):
        graph.replay()
    fs.step_idx.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    finite = bool(torch.isfinite(fs.x).all().item())
    print(f"[step_ab] {label}: {ms:.3f} ms per PC step ({precision}, {steps} steps, finite={finite}, clocks/power under load: {clk})",
          flush=True)


if __name__ == "__main__":
    main()
