"""GPU probe: per-op CUDA-event timing of one forward of the bench network (aggregated by kernel and shape)."""
import collections
import sys

import torch

sys.path.insert(0, ".")
import bench
from conditional_score_diffusion_b200 import kernels as K
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = bench.workload_config()
torch.manual_seed(0)
model = utils.create_model(cfg).cuda().eval()
import os
if os.environ.get("CSD_PRECISION"):
    model.set_precision(os.environ["CSD_PRECISION"])
model._engine.ensure_packed(torch.device("cuda", 0))
plan = model._engine.plan(B, 160, 160, 3, 3)
plan.in0.normal_(); plan.in1.uniform_(); plan.labels.fill_(500.0)
for _ in range(2):
    plan.rec.run()
torch.cuda.synchronize()
evs = []
for fn, a, kw in plan.rec.ops:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(*a, **kw); e1.record()
    evs.append((fn, a, kw, e0, e1))
torch.cuda.synchronize()
by_kind = collections.defaultdict(lambda: [0, 0.0])
by_shape = collections.defaultdict(lambda: [0, 0.0, 0.0])
for fn, a, kw, e0, e1 in evs:
    ms = e0.elapsed_time(e1)
    name = getattr(fn, "__name__", str(fn))
    by_kind[name][0] += 1; by_kind[name][1] += ms
    if fn is K.conv_gemm:
        segs, _, n, _ = a
        k_real = sum(sg[4] * sg[3] for sg in segs)
        pix = kw["batch"] * kw["h"] * kw["w"] * kw.get("z_batches", 1)
        key = ("conv", kw["h"], kw["w"], n, k_real, len(segs), kw.get("z_batches", 1))
        by_shape[key][0] += 1; by_shape[key][1] += ms; by_shape[key][2] += 2.0 * pix * n * k_real
    else:
        t0 = a[0] if a else None
        key = (name,) + tuple(t0.shape) if torch.is_tensor(t0) else (name,)
        by_shape[key][0] += 1; by_shape[key][1] += ms
total = sum(v[1] for v in by_kind.values())
print(f"total {total:.3f} ms over {len(evs)} ops (B={B})")
for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:24s} n={v[0]:4d}  {v[1]:8.3f} ms  {100 * v[1] / total:5.1f}%")
print("top shapes:")
for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:90]:
    tf = f"{v[2] / v[1] / 1e9:7.1f} TF/s" if v[2] else ""
    print(f"  {str(k):70s} n={v[0]:3d} {v[1]:8.3f} ms {tf}")
