"""GPU probe: run upfirdn2d geometries in separate processes (a faulting kernel poisons the context)."""
import subprocess
import sys

CASES = [
    (6, 8, 8, 2, 1, 2, 1), (6, 160, 160, 2, 1, 2, 1), (6, 64, 68, 2, 1, 2, 1), (6, 20, 20, 2, 1, 2, 1),
    (6, 160, 160, 1, 2, 1, 1), (6, 160, 160, 1, 1, 2, 2), (6, 8, 8, 1, 2, 1, 1), (2, 37, 53, 2, 1, 2, 1),
]

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from conditional_score_diffusion_b200 import kernels as k
from oracle import ops
planes, h, w, up, down, p0, p1 = map(int, sys.argv[1:8])
x = torch.randn(1, planes, h, w)
kk = torch.rand(4, 4)
out = k.upfirdn2d_planes(x.cuda().reshape(planes, h, w), kk.cuda(), up, up, down, down, p0, p1, p0, p1)
torch.cuda.synchronize()
ref = ops.upfirdn2d(x, kk, up, down, (p0, p1))[0]
print("max err", (out.cpu() - ref).abs().max().item())
'''

for c in CASES:
    r = subprocess.run([sys.executable, "-c", CHILD] + [str(v) for v in c], capture_output=True, text=True)
    tail = (r.stdout + r.stderr).strip().splitlines()[-1:] if (r.stdout + r.stderr).strip() else ["<no output>"]
    print(c, "rc", r.returncode, tail[0][:160])
