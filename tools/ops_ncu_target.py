"""ncu target for the kernels the north star names one by one (run under `ncu --set full -k regex:<kernel>`):
  * upfirdn2d_tile_kernel   - the module-surface `op.upfirdn2d` (NCHW fp32 planes): up x2, down x2 and 1:1 on [64,96,80,80]
  * fused_bias_act_kernel   - `op.fused_leaky_relu` on [64, 96, 160, 160]
  * attn_core_kernel        - the fused attention core at (L, C) = (400, 192), (100, 288), (25, 288), B = 64
  * conv_gemm_kernel<32,true> - the tf32 per-tap convolution, 3x3 96 -> 96 at 160 px (B = 16) and 20 px 192 -> 192 (B = 64)
  * norm_pair_kernel        - the cluster-reduced Langevin norms on [64, 3, 160, 160]
  * fir_tma_kernel / gn_fused_kernel / gn_apply_kernel / tap_shift_sum_kernel - the NHWC HBM kernels at bench shapes
Each op runs twice (warm-up + measured) between cudaProfilerStart/Stop.
"""
import math
import sys

import torch

sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import engine as E, kernels as K
from conditional_score_diffusion_b200.models import layerspp
from conditional_score_diffusion_b200.op import fused_leaky_relu, upfirdn2d

dev = torch.device("cuda", 0)
torch.manual_seed(0)
ops = []

x = torch.randn(64, 96, 80, 80, device=dev)
k4 = torch.tensor([1.0, 3.0, 3.0, 1.0], device=dev)
k2d = torch.outer(k4, k4) / 64
ops.append(lambda: upfirdn2d(x, k2d * 4, up=2, pad=(2, 1)))
ops.append(lambda: upfirdn2d(x, k2d, down=2, pad=(1, 1)))
ops.append(lambda: upfirdn2d(x, k2d, pad=(2, 2)))

xa = torch.randn(64, 96, 160, 160, device=dev)
bias = torch.randn(96, device=dev)
ops.append(lambda: fused_leaky_relu(xa, bias))


def attn_case(c, hw, b):
    blk = layerspp.AttnBlockpp(c, skip_rescale=True, init_scale=1.0).to(dev)

    class _Net(torch.nn.Module):
        pass
    net = _Net()
    net.all_modules = torch.nn.ModuleList([blk])
    eng = E.NetEngine(net)
    eng.device = dev
    pk = eng._pack_attn(blk, dev)
    rec = E.Recorder()
    bops = E.BlockOps(dev, E.BufferPool(dev), rec, torch.zeros(1 << 20, device=dev))
    a = E.Act(torch.randn(b, hw, hw, c, device=dev).to(torch.bfloat16), c)
    bops.attention(pk, a, True)
    return rec.run


for c, hw in ((192, 20), (288, 10), (288, 5)):
    ops.append(attn_case(c, hw, 64))


def tf32_conv(b, hw, cin, cout):
    a = torch.randn(b, hw, hw, cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(9 * cin)
    wt = K.pack_conv_weight(w, dtype=torch.float32)
    out = torch.empty(b, hw, hw, cout, device=dev)
    return lambda: K.conv_gemm([(a, cin, 0, cin, 9)], wt, cout, out, batch=b, h=hw, w=hw, transposed=False)


ops.append(tf32_conv(16, 160, 96, 96))
ops.append(tf32_conv(64, 20, 192, 192))

g, z, norms = xa[:, :3].contiguous(), torch.randn(64, 3, 160, 160, device=dev), torch.empty(128, device=dev)
ops.append(lambda: K.langevin_norms(g, z, norms))

act = torch.randn(64, 80, 80, 96, device=dev).to(torch.bfloat16)
up = torch.empty(64, 160, 160, 96, device=dev, dtype=torch.bfloat16)
dn = torch.empty(64, 40, 40, 96, device=dev, dtype=torch.bfloat16)
ops.append(lambda: K.fir_resample(act, up, "up", [1, 3, 3, 1]))
ops.append(lambda: K.fir_resample(act, dn, "down", [1, 3, 3, 1]))
small = torch.randn(64, 20, 20, 192, device=dev).to(torch.bfloat16)
gam, bet = torch.ones(192, device=dev), torch.zeros(192, device=dev)
so = torch.empty_like(small)
ops.append(lambda: K.gn_fused(small, 192, None, 0, gam, bet, so, 32, 1e-6, True))
big = torch.randn(64, 160, 160, 96, device=dev).to(torch.bfloat16)
sums = torch.empty(64, 96, 2, device=dev)
bo = torch.empty_like(big)
g96, b96 = torch.ones(96, device=dev), torch.zeros(96, device=dev)
ops.append(lambda: (K.gn_chan_stats(big, 96, sums), K.gn_apply(big, 96, sums, None, 0, None, g96, b96, bo, 24, 1e-6, True)))
part = torch.randn(64, 160, 160, 56, device=dev).to(torch.bfloat16)
res8 = torch.randn(64, 160, 160, 8, device=dev).to(torch.bfloat16)
o8 = torch.empty_like(res8)
b6 = torch.zeros(32, device=dev)
ops.append(lambda: K.tap_shift_sum(part, 6, b6, res8, o8))

for f in ops:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in ops:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ops_ncu_target: done")
