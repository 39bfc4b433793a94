"""GPU probe: per-iteration timestamps of the per-tap kernel's MMA thread (before / after the `full` wait) and of
producer 0's issues, for one mid-grid CTA. Needs a build with CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS."""
import os, sys, torch
sys.path.insert(0, ".")
ts = torch.zeros(192, dtype=torch.int64, device="cuda")
os.environ["CSD_DEBUG_TS"] = hex(ts.data_ptr())
from conditional_score_diffusion_b200 import kernels as k
dev = "cuda"
for (B, H, cin, cout) in ((64, 5, 288, 288), (64, 20, 192, 192)):
    a = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device=dev) / 30).to(torch.bfloat16)
    n16 = k.ceil_to(cout, 16)
    n_tile = n16 if n16 <= 256 else k.ceil_to((n16 + 1) // 2, 16)
    npad = -(-cout // n_tile) * n_tile
    wt = k.pack_conv_weight(w, n_pad=npad)
    out = torch.empty(B, H, H, cout, device=dev, dtype=torch.bfloat16)
    for rep in range(3):
        ts.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, n_tile=n_tile, transposed=False)
        e1.record()
        torch.cuda.synchronize()
    t = ts.tolist()
    t0 = t[0]
    mma = [tuple(t[16 + 4 * i + j] - t0 for j in range(4)) for i in range(12) if t[16 + 4 * i + 3]]
    prod = [t[64 + i] - t0 for i in range(96) if t[64 + i]]
    print(f"H={H} {cin}->{cout} n_tile={n_tile}: {e0.elapsed_time(e1)*1e3:.0f} us; setup={t[1]-t0} accum_ready={t[5]-t0} epi_end={t[6]-t0} dealloc={t[7]-t0}")
    print("  mma (loop top, after wait+fence, after MMA issue, after commit):", mma)
    print("  producer0 issue times:", prod[:12])
