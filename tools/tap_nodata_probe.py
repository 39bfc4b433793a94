"""GPU probe: the per-tap conv/GEMM kernel on the <= 20 px shapes with the steady-state operand loads switched off
(CSD_DEBUG_NODATA bit 0 = weights, bit 1 = activations; results are garbage) and with a capped ring depth
(CSD_TAP_STAGES) - separates MMA issue, TMA feed and ring depth for the latency-bound small levels."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "run":
    import torch
    sys.path.insert(0, ".")
    from conditional_score_diffusion_b200 import kernels as k
    tag = (f"nodata={os.environ.get('CSD_DEBUG_NODATA','0')} stages={os.environ.get('CSD_TAP_STAGES','max')} "
           f"chunk={'32' if os.environ.get('CSD_TAP_CHUNK32') else '64'}")
    for (B, H, cin, cout, taps) in ((64, 5, 288, 288, 9), (64, 10, 288, 288, 9), (64, 20, 192, 192, 9), (64, 20, 384, 192, 9),
                                    (64, 40, 192, 192, 1), (64, 160, 96, 6, 9)):
        a = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
        kk = 3 if taps == 9 else 1
        w = (torch.randn(cout, cin, kk, kk, device="cuda") / 30).to(torch.bfloat16)
        n16 = k.ceil_to(cout, 16)
        n_tile = n16 if n16 <= 256 else k.ceil_to((n16 + 1) // 2, 16)
        npad = -(-cout // n_tile) * n_tile
        wt = k.pack_conv_weight(w, n_pad=npad)
        out = torch.empty(B, H, H, k.ceil_to(cout, 8), device="cuda", dtype=torch.bfloat16)
        args = dict(batch=B, h=H, w=H, n_tile=n_tile, transposed=False)
        if taps == 1:
            args["pad"] = 0
        for _ in range(3):
            k.conv_gemm([(a, cin, 0, cin, taps)], wt, cout, out, **args)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g):
                for _ in range(20):
                    k.conv_gemm([(a, cin, 0, cin, taps)], wt, cout, out, **args)
        torch.cuda.synchronize()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        stages = taps * -(-cin // 32)
        print(f"  {tag} {H}x{H} {cin}->{cout} taps={taps}: {us:.1f} us/launch, {stages} stages, "
              f"{us * 1.9e3 / stages:.0f} cyc/stage, {2.0*B*H*H*cin*cout*taps/us/1e6:.0f} TF/s", flush=True)
else:
    cfgs = (("0", None, False),) if "--quick" in sys.argv else (("0", None, False), ("0", None, True), ("3", None, False), ("0", "2", False))
    for nd, st, c32 in cfgs:
        env = dict(os.environ, CSD_DEBUG_NODATA=nd)
        if st:
            env["CSD_TAP_STAGES"] = st
        if c32:
            env["CSD_TAP_CHUNK32"] = "1"
        subprocess.run([sys.executable, __file__, "run"], env=env, check=False)
