"""GPU debug helper: run the bench network's launch list op by op with a sync after each; report the first failure."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from conditional_score_diffusion_b200 import kernels as K
from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = bench.workload_config()
torch.manual_seed(0)
model = utils.create_model(cfg).cuda().eval()
model._engine.ensure_packed(torch.device("cuda", 0))
plan = model._engine.plan(B, 160, 160, 3, 3)
plan.in0.normal_(); plan.in1.uniform_(); plan.labels.fill_(500.0)
for i, (fn, a, kw) in enumerate(plan.rec.ops):
    try:
        fn(*a, **kw)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED op", i, getattr(fn, "__name__", fn), str(e)[:200])
        if fn is K.conv_gemm:
            segs = a[0]
            print("  segs:", [(tuple(s[0].shape), s[1], s[2], s[3], s[4], None if len(s) < 6 or s[5] is None else tuple(s[5].shape)) for s in segs])
            print("  n:", a[2], "out:", tuple(a[3].shape), {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in kw.items()})
        break
else:
    print("all", len(plan.rec.ops), "ops ran")
