"""GPU probe: weight-gradient kernels in isolation (direct MN-major kernel vs pixel-major re-layout + GEMM).
usage: python tools/wgrad_bench.py [batch h w cin cout]"""
import sys

import torch

sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as K

args = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else [50, 64, 64, 128, 128]
B, H, W, CIN, COUT = args
a = torch.randn(B, H, W, CIN, device="cuda").to(torch.bfloat16)
g = torch.randn(B, H, W, COUT, device="cuda").to(torch.bfloat16)
flops = 2.0 * B * H * W * CIN * COUT * 9


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


splits = K.wgrad_direct_splits(B, H, W, COUT, CIN, 9)
partial = torch.empty(splits, 9, COUT, CIN, device="cuda")
dw = torch.zeros(COUT, CIN, 3, 3, device="cuda")
ms = timed(lambda: K.wgrad_direct(g, 0, COUT, a, 0, CIN, 9, partial, splits))
print(f"direct   B={B} {H}x{W} {CIN}->{COUT}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF/s  splits {splits}")
ms_r = timed(lambda: K.wgrad_reduce(partial, splits, 9, COUT, CIN, 1.0, dw, CIN * 9, 9, 1))
print(f"reduce   {ms_r * 1e3:8.1f} us")
geom = K.pixmajor_geometry(B, H, W)
a_pm = K.pixmajor_alloc(geom, CIN, 3, "cuda")
g_pm = K.pixmajor_alloc(geom, COUT, 1, "cuda")
p2 = torch.empty(geom.splits, 9, COUT, CIN, device="cuda")
ms_t = timed(lambda: (K.nhwc_to_pixmajor(a, 0, CIN, geom, a_pm), K.nhwc_to_pixmajor(g, 0, COUT, geom, g_pm)))
ms_g = timed(lambda: K.wgrad_gemm(g_pm, COUT, a_pm, CIN, 9, geom, p2))
print(f"pixmajor re-layout {ms_t * 1e3:8.1f} us + GEMM {ms_g * 1e3:8.1f} us  ({flops / ms_g / 1e9:7.1f} TF/s GEMM only)")
