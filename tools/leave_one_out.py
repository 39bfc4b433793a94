"""What does each class of launches cost INSIDE the graph replay of one network evaluation? ncu times every kernel alone
with cold caches; here the forward's launch list is captured with one class of launches left out and the replay time is
compared with the full list (bench workload, B = 64). The outputs of a reduced list are meaningless - only the time is
used."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from conditional_score_diffusion_b200 import kernels as K  # noqa: E402
from conditional_score_diffusion_b200.models import ncsnpp  # noqa: E402,F401
from conditional_score_diffusion_b200.models import utils as mutils  # noqa: E402


def main():
    cfg = bench.workload_config()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = mutils.create_model(cfg).to(dev).eval()
    B, img = bench.BATCH_PER_GPU, bench.IMAGE
    eng = model._engine
    eng.ensure_packed(dev, force_refresh=True)
    plan = eng.plan(B, img, img, 3, 3)
    plan.in0.normal_()
    plan.in1.uniform_()
    plan.labels.fill_(500.0)
    ops = plan.rec.ops

    def cls(op):
        fn, args, kw = op
        name = fn.__name__
        if fn is K.conv_gemm:
            return "conv_transposed" if kw.get("transposed") else "conv_per_tap"
        return name

    classes = {}
    for op in ops:
        classes.setdefault(cls(op), 0)
        classes[cls(op)] += 1

    def _unused_timed(skip_cls, reps=20):
        def run():
            for op in ops:
                if cls(op) == skip_cls:
                    continue
                op[0](*op[1], **op[2])
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # graphs first, then interleaved timing rounds (the clock drifts by ~3 % over a minute under the power cap)
    names = [None] + [c for c, n in sorted(classes.items(), key=lambda kv: -kv[1]) if n >= 3]
    graphs = {}
    for c in names:
        def run(c=c):
            for op in ops:
                if c is not None and cls(op) == c:
                    continue
                op[0](*op[1], **op[2])
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        graphs[c] = g
    rounds, reps = 6, 10
    acc = {c: [] for c in names}
    for r in range(rounds):
        for c in names:
            g = graphs[c]
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            acc[c].append(e0.elapsed_time(e1) / reps)
    med = {c: sorted(v)[len(v) // 2] for c, v in acc.items()}
    full = med[None]
    print(f"[loo] full forward: {full:.3f} ms (rounds: {' '.join(f'{x:.3f}' for x in acc[None])}), {len(ops)} launches")
    for c in names[1:]:
        n = classes[c]
        d = [a - b for a, b in zip(acc[None], acc[c])]      # same-round differences
        dm = sorted(d)[len(d) // 2]
        print(f"[loo] {c:22s} {n:3d} launches: costs {dm:.3f} ms inside the graph (min {min(d):.3f}, max {max(d):.3f}) = "
              f"{dm / n * 1e3:.1f} us per launch", flush=True)


if __name__ == "__main__":
    main()
