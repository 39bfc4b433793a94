"""GPU probe: persistent transposed conv with the operand loads switched off (CSD_DEBUG_NODATA bit 0 = weights,
bit 1 = activations; results are garbage) - separates the MMA issue rate from the data feed."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ".")
    from conditional_score_diffusion_b200 import kernels as k
    for (B, H, cin, cout) in ((64, 160, 96, 96), (64, 160, 192, 96), (64, 80, 192, 192), (64, 80, 96, 96), (64, 40, 192, 192), (64, 40, 384, 192)):
        a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
        wt = k.pack_conv_weight((torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(torch.bfloat16))
        out = torch.empty(B, H, H, cout, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, transposed=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            k.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=B, h=H, w=H, transposed=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"  nodata={os.environ.get('CSD_DEBUG_NODATA','0')} {H}x{H} {cin}->{cout}: {ms*1e3:.0f} us {2.0*B*H*H*cin*cout*9/ms/1e9:.0f} TF/s")
else:
    for nd in (sys.argv[2:] if False else os.environ.get("CSD_NODATA_LIST", "0,1,2,3").split(",")):
        env = dict(os.environ, CSD_DEBUG_NODATA=nd)
        subprocess.run([sys.executable, __file__, "run"], env=env, check=False)
