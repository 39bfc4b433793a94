"""ncu target: the one-launch GroupNorm+SiLU at the 20 px / 10 px / 5 px levels of config 2 (B = 64), bf16."""
import sys

import torch

sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as K

dev = "cuda"
for (hw, c) in ((400, 192), (100, 288), (25, 288)):
    x = torch.randn(64, hw, c, device=dev).to(torch.bfloat16)
    out = torch.empty_like(x)
    gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
    for _ in range(2):
        K.gn_fused(x, c, None, 0, gamma, beta, out, min(c // 4, 32))
    torch.cuda.synchronize()
