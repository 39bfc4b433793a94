"""Graph-timed probe of the one-launch GroupNorm (+SiLU) at the small levels of config 2 (B = 64).
Prints microseconds per launch from a CUDA graph of 64 back-to-back launches (warm L2, no launch gaps) - the
number a PC step actually pays. Plan knobs are read by the library per launch: CSD_GNF_MINVEC, CSD_GNF_THREADS."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conditional_score_diffusion_b200 import kernels as K  # noqa: E402


def time_graph(fn, reps=64, rounds=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(rounds):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def main():
    dev = "cuda"
    B = 64
    for dtype in (torch.bfloat16, torch.float32):
        for (hw, c) in ((400, 192), (400, 288), (100, 288), (25, 288), (100, 192)):
            if not K.gn_fused_supported(c, 0, hw, min(c // 4, 32), B, dtype):
                print(f"{dtype} hw={hw} c={c}: unsupported")
                continue
            # several distinct tensors so that consecutive launches do not hit the same lines
            xs = [torch.randn(B, hw, c, device=dev).to(dtype) for _ in range(4)]
            outs = [torch.empty_like(x) for x in xs]
            gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
            i = [0]

            def fn():
                j = i[0] % 4
                i[0] += 1
                K.gn_fused(xs[j], c, None, 0, gamma, beta, outs[j], min(c // 4, 32))
            us = time_graph(fn)
            nbytes = 2 * xs[0].numel() * xs[0].element_size()
            print(f"{str(dtype):16s} hw={hw:4d} c={c:4d}: {us:7.2f} us/launch  {nbytes / us / 1e3:7.1f} GB/s "
                  f"(env MINVEC={os.environ.get('CSD_GNF_MINVEC', '-')} THREADS={os.environ.get('CSD_GNF_THREADS', '-')})")


if __name__ == "__main__":
    main()
