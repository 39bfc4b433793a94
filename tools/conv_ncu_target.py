"""ncu target: a few launches of the 3x3 96->96 160x160 convolution in per-tap and halo modes."""
import sys
import torch
sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as k
B, H, cin, cout = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 160, 96, 96
a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
out = torch.empty(B, H, H, cout, device="cuda", dtype=torch.bfloat16)
for kw in (dict(halo=False), dict(halo=True, mt=2), dict(halo=False), dict(halo=True, mt=2)):
    k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, **kw)
torch.cuda.synchronize()
