"""GPU probe: phase timestamps (SM clock) of one mid-grid CTA of the conv kernels."""
import os, sys, torch
sys.path.insert(0, ".")
ts = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["CSD_DEBUG_TS"] = hex(ts.data_ptr())
from conditional_score_diffusion_b200 import kernels as k
B, H, cin, cout = 64, 160, 96, 96
a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
out = torch.empty(B, H, H, cout, device="cuda", dtype=torch.bfloat16)
names = ["start", "setup done", "-", "first full", "mma issued", "epi start", "epi end", "dealloc", "mid"]
for label, kw in [("tap", dict(halo=False)), ("halo mt=2", dict(halo=True, mt=2)), ("transposed", dict(transposed=True))]:
    for rep in range(2):
        ts.zero_()
        k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, **kw)
        torch.cuda.synchronize()
    t = ts.tolist()
    print(label, {names[i]: t[i] - t[0] for i in (1, 3, 8, 4, 5, 6, 7)})
