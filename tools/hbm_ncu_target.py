"""ncu target for the two HBM kernels changed in the last session of round 2: the up x2 FIR on [64, 80, 80, 96] bf16
(whole 96-channel rows per CTA; CSD_FIR_CHUNK64=1 gives the earlier 64 + 32 channel chunks) and the tap-stacked heads'
shift-sum on [64, 160, 160, 56] (CSD_SHIFT_SUM_LEGACY=1: the per-pixel kernel). Each op runs twice (warm-up + measured)
between cudaProfilerStart/Stop:   ncu --set full --profile-from-start off -k regex:"fir_tma|tap_shift" ..."""
import sys

import torch

sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as K

dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(64, 80, 80, 96, device=dev).to(torch.bfloat16)
up = torch.empty(64, 160, 160, 96, device=dev, dtype=torch.bfloat16)
xd = torch.randn(64, 160, 160, 96, device=dev).to(torch.bfloat16)
dn = torch.empty(64, 80, 80, 96, device=dev, dtype=torch.bfloat16)
part = torch.randn(64, 160, 160, 56, device=dev).to(torch.bfloat16)
res = torch.randn(64, 160, 160, 8, device=dev).to(torch.bfloat16)
bias = torch.randn(6, device=dev)
out = torch.empty(64, 160, 160, 8, device=dev, dtype=torch.bfloat16)
ops = [lambda: K.fir_resample(x, up, "up", [1, 3, 3, 1]), lambda: K.fir_resample(xd, dn, "down", [1, 3, 3, 1]),
       lambda: K.tap_shift_sum(part, 6, bias, res, out)]
for op in ops:
    op()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for op in ops:
    op()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
