"""CPU oracle for the score-network / PC-sampler hot path.

TEST INFRASTRUCTURE ONLY. Nothing under `conditional_score_diffusion_b200/` may import this
package: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` use it, and only as the checker or the timed CPU baseline.

The oracle is a plain PyTorch-fp32 restatement (CPU, functional style over a flat parameter
dict) of the reference algorithm; every function cites the reference file:line it follows.
It is pinned against golden vectors produced by importing the real reference
(`tests/golden/make_golden.py`, run in the build container where /root/reference is mounted;
the vectors are committed under tests/golden/). The reference's own tests hold no golden
vectors, so those generated fixtures are the pin.
"""
