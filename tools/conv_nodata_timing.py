"""GPU probe: which load stream limits the transposed halo conv? Times the kernel with the steady-state
weight loads (bit 0) and / or activation loads (bit 1) skipped (results are garbage in those runs)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as k
B = 64
for (H, cin, cout) in ((160, 96, 96), (160, 192, 96), (80, 192, 192)):
    a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
    wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
    out = torch.empty(B, H, H, cout, device="cuda", dtype=torch.bfloat16)
    for nodata in (0, 1, 2, 3):
        os.environ["CSD_DEBUG_NODATA"] = str(nodata)
        for _ in range(3):
            k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * H * H * cout * cin * 9
        print(f"{H}x{H} {cin}->{cout} nodata={nodata}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
    os.environ["CSD_DEBUG_NODATA"] = "0"
