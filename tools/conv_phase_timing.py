"""GPU probe: role timestamps (SM clock) of the 6th tile of one mid-grid CTA of the persistent transposed conv.
Needs a build with CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS."""
import math, os, sys, torch
sys.path.insert(0, ".")
ts = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["CSD_DEBUG_TS"] = hex(ts.data_ptr())
from conditional_score_diffusion_b200 import kernels as k
names = {10: "prod tile start", 11: "prod A0 slot free", 8: "xform A0 full", 9: "xform A0 done", 12: "mma wants acc",
         0: "mma acc free", 1: "mma A0 go", 13: "mma A1 go", 2: "mma issued", 3: "epi waits", 4: "epi acc full",
         5: "epi staged", 6: "epi stored", 7: "epi stats"}
dev = "cuda"
DT = torch.float32 if os.environ.get("CSD_PRECISION") == "tf32" else torch.bfloat16   # tf32: the fp32 / kind::tf32 instance
if DT == torch.float32:
    names[6] = "epi pass A staged"
SHAPES = ((64, 160, 96, 96, False), (64, 160, 96, 96, True), (64, 160, 192, 96, True), (64, 80, 192, 192, True),
          (64, 80, 96, 96, False), (64, 40, 192, 192, False), (64, 40, 192, 192, True))
if os.environ.get("CSD_DEBUG_NODATA", "0") != "0":
    SHAPES = tuple(sh for sh in SHAPES if not sh[4])     # the fused prologue cannot run without its loads
for (B, H, cin, cout, full) in SHAPES:
    a = torch.randn(B, H, H, cin, device=dev).to(DT)
    wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device=dev) / 30).to(DT), dtype=DT)
    out = torch.empty(B, H, H, cout, device=dev, dtype=DT)
    kw = {}
    seg = (a, cin, 0, cin, 9)
    if full:   # what the engine launches: fused GroupNorm prologue, temb, residual, statistics
        coef = torch.rand(B, cin, 2, device=dev)
        seg = (a, cin, 0, cin, 9, coef, True)
        tiles = B * math.ceil(H / k.transposed_tile_rows(H)) * math.ceil(H / 8) * 2
        kw = dict(temb=torch.randn(B, cout + 32, device=dev), temb_pitch=cout + 32,
                  bias=torch.randn(cout + 32, device=dev), stat_partials=torch.empty(tiles, cout, 2, device=dev))
    for rep in range(3):
        ts.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k.conv_gemm([seg], wt, cout, out, batch=B, h=H, w=H, transposed=True, **kw)
        e1.record()
        torch.cuda.synchronize()
    t = ts.tolist()
    t0 = min(v for v in t if v)
    ev = sorted((t[i] - t0, names[i]) for i in names if t[i])
    print(f"H={H} cin={cin} cout={cout} full={full}: {e0.elapsed_time(e1)*1e3:.0f} us;", ", ".join(f"{n}@{c}" for c, n in ev))
