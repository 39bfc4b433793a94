"""ncu target: the 3x3 convolutions that dominate the forward, in the mode the engine uses for them."""
import sys
import torch
sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as k
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (H, cin, cout) in ((160, 96, 96), (80, 192, 192), (40, 192, 192)):
    a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
    wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
    out = torch.empty(B, H, H, cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H)
torch.cuda.synchronize()
