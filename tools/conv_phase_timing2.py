"""GPU probe: epilogue duration with and without co-resident CTAs issuing MMAs."""
import os, sys, torch
sys.path.insert(0, ".")
ts = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["CSD_DEBUG_TS"] = hex(ts.data_ptr())
from conditional_score_diffusion_b200 import kernels as k
names = ["start", "setup done", "-", "first full", "mma issued", "epi start", "epi end", "dealloc", "mid"]
for (B, H, W) in [(1, 16, 8), (1, 160, 120), (4, 160, 160), (64, 160, 160)]:
    cin = cout = 96
    a = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
    out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    for rep in range(2):
        ts.zero_()
        k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=W, tile=(8, 16, 1))
        torch.cuda.synchronize()
    t = ts.tolist()
    print((B, H, W), "tiles", B * H * W // 128, {names[i]: t[i] - t[0] for i in (1, 3, 8, 4, 5, 6, 7)})
