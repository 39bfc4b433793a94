"""GPU probe: phase timestamps (SM clock) of one mid-grid CTA of the conv kernels.
Needs a build with CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS."""
import os, sys, torch
sys.path.insert(0, ".")
ts = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["CSD_DEBUG_TS"] = hex(ts.data_ptr())
from conditional_score_diffusion_b200 import kernels as k
names = {1: "setup done", 3: "first A full", 8: "second A full", 4: "mma issued", 5: "accum ready", 9: "staged",
         6: "epi end", 7: "dealloc"}
for (B, H, cin, cout, with_res) in ((64, 160, 96, 96, False), (64, 160, 96, 96, True), (64, 160, 192, 96, False),
                                    (64, 80, 192, 192, False)):
    a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
    wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
    out = torch.empty(B, H, H, cout, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(B, H, H, cout, device="cuda").to(torch.bfloat16) if with_res else None
    for label, kw in [("tap", dict(transposed=False)), ("transposed", dict(transposed=True))]:
        for rep in range(3):
            ts.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, res=res, res_pitch=cout, **kw)
            e1.record()
            torch.cuda.synchronize()
        t = ts.tolist()
        print(f"H={H} cin={cin} cout={cout} res={with_res} {label}: {e0.elapsed_time(e1)*1e3:.0f} us; cycles since CTA start:",
              {names[i]: t[i] - t[0] for i in (1, 3, 8, 4, 5, 9, 6, 7) if t[i]})
