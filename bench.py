#!/usr/bin/env python
"""Headline benchmark: sampled images/sec, NCSN++ 160 px VE-SDE, predictor-corrector 1000 steps.

Workload (BASELINE.json configs[1], SURVEY.md §8d "Config 2"): celebA_ours_NDV_160.py in its NCSN++
form (`ncsnpp_paired`, nf 96, ch_mult (1,1,2,2,3,3), attention at 20/10/5, 6 in / 6 out channels,
160x160), conditional PC sampler = conditional Langevin corrector + conditional reverse-diffusion
predictor, snr 0.15, {'x': cVESDE(5e-3, 277.13, 1000), 'y': VESDE(5e-3, 0.5, 1000)}, batch 64 per GPU,
synthetic y = rand, random-init weights with init_scale=1 (init_scale=0 weights make the Langevin
step size overflow).

A "step" is ONE PC step (corrector + predictor = 2 network evaluations + the update kernels) over
the whole per-GPU batch. images/sec = batch * gpus / (seconds per step * 1000): every one of the
1000 steps of a PC-1000 trajectory has the same cost, so K timed steps give the per-image throughput
of the full trajectory (default K = 200; --steps 1000 times a complete PC-1000 trajectory).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Prints ONE JSON line (see the keys in main()). `--impl reference` times the UNMODIFIED reference
(baseline/_ref, installed by baseline/install_ref.py: sampling/conditional.py driving models/ncsnpp.py,
PyTorch fp32 on all host cores) on a bounded sample of the same workload - batch 2, a few PC steps,
extrapolated to PC-1000; it falls back to the oracle port (oracle/, `kind: "port"`) only where
baseline/_ref did not travel. The default line also carries the same reference on the GPU (`stock_gpu`:
eager PyTorch, cuDNN TF32, its own upfirdn2d CUDA op, B = 64) and `vs_stock_gpu`.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PC_STEPS = 1000
BATCH_PER_GPU = 64
IMAGE = 160
SNR = 0.15
EPS = 1e-5
METRIC = "sampled images/sec NCSN++ 160px VE PC-1000"


def workload_config():
    """BASELINE configs[1] in its NCSN++ form (conditional_score_diffusion_b200/workloads.py)."""
    from conditional_score_diffusion_b200 import workloads
    return workloads.config2_ncsnpp_paired_160()


_SAVED_STDOUT = []


def quiet_stdout():
    """Point file descriptor 1 at stderr until emit() restores it: native libraries (NCCL's version banner) then cannot
    put extra lines in front of the one JSON line the driver parses."""
    if not _SAVED_STDOUT:
        sys.stdout.flush()
        _SAVED_STDOUT.append(os.dup(1))
        os.dup2(2, 1)


def emit(line):
    """Print the result line on the real stdout."""
    sys.stdout.flush()
    if _SAVED_STDOUT:
        os.dup2(_SAVED_STDOUT.pop(), 1)
    print(json.dumps(line))
    sys.stdout.flush()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_steps(cfg, sample_batch, steps, warmup, time_budget_s):
    """Time `steps` PC steps of the oracle (CPU, torch fp32, all host threads) at batch `sample_batch`."""
    import torch
    from oracle import ncsnpp as o_net
    from oracle import sampling as o_samp
    from oracle import sde as o_sde
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401  (weights + layout only)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = utils.create_model(cfg)   # parameter container only: the oracle evaluates the state dict on CPU
    params = {k: v.detach() for k, v in model.state_dict().items()}
    o = o_net.model_options(cfg)
    spec = o_net.build_spec(o)
    sx = o_sde.VE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000)
    sy = o_sde.VE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)
    model_fn = lambda d, l: o_net.forward_paired(params, o, d["x"], d["y"], l, spec)
    score_fn = o_sde.score_fn_conditional_pair(model_fn, sx, sy, True)
    shape = (sample_batch, 3, cfg.data.image_size, cfg.data.image_size)
    y = torch.rand(*shape)
    x = sx.prior_sampling(shape)
    ts = torch.linspace(1.0, EPS, PC_STEPS)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            vec_t = torch.ones(sample_batch) * ts[i]
            y_t = y + torch.randn_like(y) * sy.sigma(vec_t)[:, None, None, None]
            grad = score_fn(x, y_t, vec_t)
            x, _ = o_samp.langevin_update(sx, grad, x, vec_t, torch.randn_like(x), SNR)
            y_t = y + torch.randn_like(y) * sy.sigma(vec_t)[:, None, None, None]
            score = score_fn(x, y_t, vec_t)
            x, _ = o_samp.reverse_diffusion_update(sx, score, x, vec_t, torch.randn_like(x))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
                if sum(times) > time_budget_s:
                    break
    s_per_step = sum(times) / len(times)
    return {"s_per_step": s_per_step, "steps_timed": len(times), "cores": cores,
            "images_per_s": sample_batch / (s_per_step * PC_STEPS)}


def bench_state_dict(cfg):
    """The bench network's weights (seed 0, init_scale = 1) as a CPU state dict: the reference's own modules load it
    strictly (same parameter names and shapes = the checkpoint contract)."""
    import torch
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401  (parameter container only)
    torch.manual_seed(0)
    model = utils.create_model(cfg)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def reference_cpu_steps(cfg, sample_batch, steps, warmup):
    """Time the UNMODIFIED reference (baseline/_ref: sampling/conditional.py:47-228 driving models/ncsnpp.py) on the
    host cores; falls back to the oracle port when baseline/_ref did not travel. Returns (result, kind)."""
    import torch
    from baseline import ref_harness
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if not ref_harness.available():
        r = cpu_oracle_steps(cfg, sample_batch, steps, warmup, time_budget_s=150.0)
        return r, "port"
    torch.manual_seed(1)
    y = torch.rand(sample_batch, 3, cfg.data.image_size, cfg.data.image_size)
    r = ref_harness.conditional_pc_time(cfg, bench_state_dict(cfg), y, steps, warmup, "cpu", SNR, EPS)
    return {"s_per_step": r["s_per_step"], "steps_timed": r["steps"], "cores": cores, "finite": r["finite"],
            "images_per_s": sample_batch / (r["s_per_step"] * PC_STEPS)}, "reference"


def stock_gpu_steps(cfg, batch, steps, warmup, device):
    """The denominator of the north star's ">= 10x stock PyTorch + cuDNN": the UNMODIFIED reference on the same B200 -
    its own eager modules, fp32 storage, PyTorch defaults (cuDNN convolutions with TF32 allowed, fp32 matmul), its own
    JIT-built `upfirdn2d` CUDA op (pre-built by baseline/install_ref.py from the reference's sources)."""
    import torch
    from baseline import ref_harness
    if not ref_harness.available():
        return None
    torch.manual_seed(2)
    y = torch.rand(batch, 3, cfg.data.image_size, cfg.data.image_size)
    r = ref_harness.conditional_pc_time(cfg, bench_state_dict(cfg), y, steps, warmup, device, SNR, EPS)
    # what the reference's own precision class (fp32 storage, cuDNN TF32 convolutions) costs against exact fp32: the
    # same network / weights / inputs as network_parity(), evaluated by the unmodified reference on the GPU
    par = None
    try:
        from oracle import ncsnpp as o_net
        sd = bench_state_dict(cfg)
        x, yy, labels = parity_inputs(cfg)
        ref = o_net.forward_paired(sd, o_net.model_options(cfg), x, yy, labels)
        rm = ref_harness.create_model(cfg, sd, device)
        with torch.no_grad():
            out = rm({"x": x.to(device), "y": yy.to(device)}, labels.to(device))
        mx = max((out[k].cpu() - ref[k]).abs().max().item() / ref[k].abs().max().item() for k in ("x", "y"))
        l2 = max(((out[k].cpu() - ref[k]).norm() / ref[k].norm()).item() for k in ("x", "y"))
        par = {"max_rel": mx, "l2_rel": l2, "what": "unmodified reference on cuda (TF32 convolutions) vs CPU fp32 oracle, "
                                                    "same network / inputs as `parity`"}
        del rm
    except Exception as e:  # noqa: BLE001
        par = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    torch.cuda.empty_cache()
    ms = r["s_per_step"] * 1e3
    return {"ms_per_step": ms, "parity_vs_fp32": par, "images_per_s": batch / (r["s_per_step"] * PC_STEPS), "batch": batch, "steps": steps,
            "warmup_steps": warmup, "finite_output": r["finite"], "dtype": "f32 storage, cuDNN TF32 convolutions (PyTorch defaults)",
            "what": "unmodified reference (baseline/_ref) on cuda: sampling/conditional.py PC loop, models/ncsnpp.py eager "
                    "modules, the reference's own upfirdn2d CUDA extension"}


def torch_eager_gpu_steps(cfg, batch, steps, warmup):
    """Context number, not a contract arm: the oracle's plain-PyTorch restatement of the reference path run on
    cuda:0 with PyTorch defaults (fp32 storage, cuDNN convolutions with TF32 allowed, fp32 matmul, eager launches) -
    the closest stand-in for "stock PyTorch + cuDNN" that can travel to the GPU box (the reference itself is not
    installed there). Its FIR resampling goes through F.conv2d like the reference's own `upfirdn2d_native` branch
    (op/upfirdn2d.py:159-200), not through the reference's JIT-built CUDA op."""
    import torch
    from oracle import ncsnpp as o_net
    from oracle import sampling as o_samp
    from oracle import sde as o_sde
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    torch.manual_seed(0)
    model = utils.create_model(cfg)
    torch.set_default_device("cuda")
    params = {k: v.detach().cuda() for k, v in model.state_dict().items()}
    o = o_net.model_options(cfg)
    spec = o_net.build_spec(o)
    sx = o_sde.VE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000)
    sy = o_sde.VE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)
    model_fn = lambda d, l: o_net.forward_paired(params, o, d["x"], d["y"], l, spec)
    score_fn = o_sde.score_fn_conditional_pair(model_fn, sx, sy, True)
    shape = (batch, 3, cfg.data.image_size, cfg.data.image_size)
    y = torch.rand(*shape)
    x = torch.randn(*shape) * cfg.model.sigma_max_x
    ts = torch.linspace(1.0, EPS, PC_STEPS)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                e0.record()
            vec_t = torch.ones(batch) * ts[i]
            y_t = y + torch.randn_like(y) * sy.sigma(vec_t)[:, None, None, None]
            grad = score_fn(x, y_t, vec_t)
            x, _ = o_samp.langevin_update(sx, grad, x, vec_t, torch.randn_like(x), SNR)
            y_t = y + torch.randn_like(y) * sy.sigma(vec_t)[:, None, None, None]
            score = score_fn(x, y_t, vec_t)
            x, _ = o_samp.reverse_diffusion_update(sx, score, x, vec_t, torch.randn_like(x))
        e1.record()
        torch.cuda.synchronize()
    torch.set_default_device("cpu")
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "images_per_s": batch / (ms * 1e-3 * PC_STEPS), "batch": batch, "steps": steps,
            "what": "oracle restatement of the reference path in eager PyTorch on cuda:0 (fp32 storage, cuDNN TF32 "
                    "convolutions, F.conv2d FIR); context only"}


def run_torch_eager_gpu(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = torch_eager_gpu_steps(workload_config(), BATCH_PER_GPU, max(3, min(args.steps, 20)), 2)
    print(json.dumps({"impl": "torch_eager_gpu", "metric": METRIC, "value": r["images_per_s"], "unit": "images/s",
                      "ms_per_step": r["ms_per_step"], "n_gpus": 1, "steps": r["steps"], "dtype": "f32/tf32",
                      "config": {"workload": "same as the b200 arm", "batch_per_gpu": r["batch"]}, "note": r["what"]}))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config()
    sample_b = 2
    steps = max(1, min(args.steps, 20))
    r, kind = reference_cpu_steps(cfg, sample_b, steps, 1)
    what = ("the unmodified reference (baseline/_ref: sampling/conditional.py + models/ncsnpp.py, torch CPU fp32)"
            if kind == "reference" else "oracle port of the reference path (torch CPU fp32)")
    sample = (f"batch {sample_b} of {BATCH_PER_GPU}, {r['steps_timed']} PC steps timed after 1 warm-up step, {what}, "
              f"{r['cores']} threads, extrapolated x1000 steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["images_per_s"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": r["steps_timed"], "warmup": 1, "ms_per_step": r["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "celebA_ours_NDV_160 as ncsnpp_paired nf96, conditional PC-1000, 160x160",
                   "batch_per_step": sample_b, "pc_steps_per_image": PC_STEPS},
        "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"], "kind": kind,
                         "sample": sample},
        "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def identity_segments(eng):
    """wt data_ptr -> per-segment flag 'identity weights' (a residual riding as a K segment): those MACs are not
    algorithmic work of the reference's network and are left out of every FLOP figure."""
    from conditional_score_diffusion_b200.engine import PackedConv
    table = {}

    def add(obj):
        if isinstance(obj, PackedConv):
            table[obj.wt.data_ptr()] = [any(s.kind == "eye" for s in seg) for seg in obj.segs]

    eng._walk_packed(add)
    return table


def conv_profile(plan):
    """CUDA-event time and algorithmic FLOPs of every tcgen05 conv/GEMM launch of one forward."""
    import torch
    from conditional_score_diffusion_b200 import kernels as K
    eye = identity_segments(plan.eng)
    ev = []
    torch.cuda.synchronize()
    for fn, a, kw in plan.rec.ops:
        if fn is K.conv_gemm:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **kw)
            e1.record()
            segs, _, n, _ = a
            pixels = kw["batch"] * kw["h"] * kw["w"] * kw.get("z_batches", 1)
            flags = eye.get(a[1].data_ptr(), [])
            k_real = sum(sg[4] * sg[3] for i, sg in enumerate(segs) if not (i < len(flags) and flags[i]))
            head = bool(kw.get("transposed")) and segs[0][4] == 1       # tap-stacked output head (1x1 to 9 * Cout rows)
            ev.append((e0, e1, 2.0 * pixels * n * k_real, (kw["h"], kw["w"], n, k_real), bool(kw.get("transposed")), head))
        else:
            fn(*a, **kw)
    torch.cuda.synchronize()
    rows = [(e0.elapsed_time(e1), fl, shp, tp, head) for e0, e1, fl, shp, tp, head in ev]
    return rows


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of one PC step (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def parity_inputs(cfg):
    import torch
    g = torch.Generator().manual_seed(77)
    hw = cfg.data.image_size
    x = torch.randn(2, 3, hw, hw, generator=g) * 20.0
    y = torch.rand(2, 3, hw, hw, generator=g)
    labels = torch.tensor([700.0, 150.0])
    return x, y, labels


def network_parity(cfg, model, dev, precision="bf16"):
    """max / L2 relative error of the benchmarked network (the bench weights, B = 2, sigma-scaled inputs) against the
    CPU oracle restatement of models/ncsnpp.py (pinned to the unmodified reference by tests/test_oracle_golden.py)."""
    import torch
    from oracle import ncsnpp as o_net
    x, y, labels = parity_inputs(cfg)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    o = o_net.model_options(cfg)
    with torch.no_grad():
        ref = o_net.forward_paired(sd, o, x, y, labels)
        out = model({"x": x.to(dev), "y": y.to(dev)}, labels.to(dev))
    res = {"config": "configs[1] network (ncsnpp_paired 160 px nf 96), bench weights, B=2, vs CPU fp32 oracle",
           "precision": precision}
    mx = l2 = 0.0
    for k in ("x", "y"):
        got, r = out[k].float().cpu(), ref[k]
        mx = max(mx, (got - r).abs().max().item() / (r.abs().max().item() + 1e-30))
        l2 = max(l2, ((got - r).norm() / (r.norm() + 1e-30)).item())
    res["max_rel"], res["l2_rel"] = mx, l2
    return res


def tf32_line(cfg, model, sde, shape, dev, steps, B, world):
    """The same PC step on the reference-precision plan (model.set_precision('tf32')): device-timed like `value`
    (rank 0 only; a context line beside the bf16 headline, same workload, same batch)."""
    import torch
    from conditional_score_diffusion_b200.sampling.fused import FusedPCSampler
    model.set_precision("tf32")
    fs = FusedPCSampler(model, sde, shape, "reverse_diffusion", "langevin", SNR, PC_STEPS, 1, False, True, True, EPS,
                        conditional=True)
    fs._setup(dev)
    fs.y.copy_(torch.rand(*shape, device=dev))
    graph = fs._graph(draw_noise=True)
    fs.x.copy_(torch.randn(*shape, device=dev) * cfg.model.sigma_max_x)
    fs.step_idx.zero_()
    for _ in range(3):
        graph.replay()
    fs.step_idx.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    finite = bool(torch.isfinite(fs.x).all().item())
    par = network_parity(cfg, model, dev, "tf32")
    gb = fs.plan.pool.nbytes() / 1e9
    del fs, graph
    torch.cuda.empty_cache()
    return {"dtype": "tf32 operands, fp32 activations in HBM (the reference's precision class)", "ms_per_step": ms,
            "value": B / (ms * 1e-3 * PC_STEPS), "unit": "images/s per GPU", "steps": steps, "finite_output": finite,
            "parity": par, "plan_buffers_gb": gb}


def _pc_steps_ms(model, sde, shape, conditional, snr, n_scales, steps, dev, eps=EPS):
    """ms per PC step of a fused sampler (device-timed graph replays, inputs resident)."""
    import torch
    from conditional_score_diffusion_b200.sampling.fused import FusedPCSampler
    fs = FusedPCSampler(model, sde, shape, "reverse_diffusion", "langevin", snr, n_scales, 1, False, True, True, eps,
                        conditional=conditional)
    fs._setup(dev)
    if conditional:
        fs.y.copy_(torch.rand(*shape, device=dev))
    graph = fs._graph(draw_noise=True)
    c_sde = sde["x"] if isinstance(sde, dict) else sde
    fs.x.copy_(torch.randn(*shape, device=dev) * c_sde.sigma_max)
    fs.step_idx.zero_()
    for _ in range(3):
        graph.replay()
    fs.step_idx.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(fs.x).all().item())
    gb = fs.plan.pool.nbytes() / 1e9
    return e0.elapsed_time(e1) / steps, finite, gb


def other_configs(dev, steps=5):
    """Short device-timed records of the other BASELINE.json configs (rank 0, after the headline): configs[2]
    (ddpm_paired 128 px inpainting net, conditional PC), configs[4] (NCSN++ 256 px nf 128, unconditional PC-2000) and
    configs[3] (ddpm_paired_SR3 training step at 64 px as shipped and at 128 px). Context for the reader and the driver,
    not part of `value`."""
    import torch
    from conditional_score_diffusion_b200 import losses, optim, sde_lib, workloads
    from conditional_score_diffusion_b200.models import ddpm, ncsnpp, utils  # noqa: F401
    out = {}

    def guard(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001 - context records must not take the headline down
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    def cfg3():
        cfg = workloads.config3_ddpm_paired_128()
        torch.manual_seed(3)
        m = utils.create_model(cfg).to(dev).eval()
        sde = {"x": sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000),
               "y": sde_lib.VESDE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)}
        B = 64
        ms, finite, gb = _pc_steps_ms(m, sde, (B, 3, 128, 128), True, 0.15, 1000, steps, dev)
        return {"workload": "configs[2]: inpainting/celebA_ours_DV.py as shipped (ddpm_paired nf96, 128x128), conditional "
                            "PC-1000, batch 64 per GPU (512 over 8 GPUs)", "ms_per_pc_step": ms,
                "images_per_s_per_gpu": B / (ms * 1e-3 * 1000), "finite_output": finite, "plan_buffers_gb": gb}

    def cfg5():
        cfg = workloads.config5_ncsnpp_256()
        torch.manual_seed(5)
        m = utils.create_model(cfg).to(dev).eval()
        sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
        B = 16
        ms, finite, gb = _pc_steps_ms(m, sde, (B, 3, 256, 256), False, cfg.sampling.snr, cfg.model.num_scales, steps, dev)
        return {"workload": "configs[4]: church_ncsnpp_continuous (NCSN++ nf128, 7 levels, attention at 16x16, 256x256), "
                            "unconditional PC-2000, batch 16 per GPU (128 over 8 GPUs)", "ms_per_pc_step": ms,
                "images_per_s_per_gpu": B / (ms * 1e-3 * cfg.model.num_scales), "finite_output": finite,
                "plan_buffers_gb": gb}

    def train(image, B):
        def run():
            cfg = workloads.config4_ddpm_sr3_64(image)
            torch.manual_seed(4)
            m = utils.create_model(cfg).to(dev).train()
            sde = sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000)
            loss_fn = losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                                     likelihood_weighting=True)
            opt = optim.FusedAdamEMA(m.parameters(), lr=cfg.optim.lr, betas=(cfg.optim.beta1, 0.999), eps=cfg.optim.eps,
                                     weight_decay=cfg.optim.weight_decay, grad_clip=cfg.optim.grad_clip, ema_decay=0.999,
                                     warmup=cfg.optim.warmup, model=m)
            x, y = torch.rand(B, 3, image, image, device=dev), torch.rand(B, 3, image, image, device=dev)

            def step():
                opt.zero_grad()
                loss = loss_fn(m, (y, x))
                loss.backward()
                opt.step()
                return loss

            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return {"workload": f"configs[3]: edges2shoes_SR3 training step (ddpm_paired_SR3 nf128, {image}x{image}, dropout "
                                f"0.1, SR3 loss, clip 1.0, fused Adam + EMA), batch {B} per GPU", "ms_per_step": ms,
                    "images_per_s_per_gpu": B / (ms * 1e-3), "finite_loss": bool(math.isfinite(loss.item()))}
        return run

    guard("config3_ddpm_paired_128_pc", cfg3)
    guard("config5_ncsnpp_256_pc", cfg5)
    guard("config4_train_64", train(64, 50))
    guard("config4_train_128", train(128, 25))
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from conditional_score_diffusion_b200 import _lib, sampling, sde_lib
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback "
                           "(use --impl reference for the CPU oracle timing)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        quiet_stdout()      # NCCL prints its version banner on stdout at the first collective: stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config()
    B, K_steps, W = BATCH_PER_GPU, args.steps, max(args.warmup, 3)

    torch.manual_seed(0)
    model = utils.create_model(cfg).to(dev).eval()
    if world > 1:
        # the only collective of the sampling path: one weight broadcast at init
        flat = torch.cat([p.data.reshape(-1) for p in model.parameters()])
        dist.broadcast(flat, src=0)
        off = 0
        for p in model.parameters():
            p.data.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    sde = {"x": sde_lib.cVESDE(cfg.model.sigma_min_x, cfg.model.sigma_max_x, 1000),
           "y": sde_lib.VESDE(cfg.model.sigma_min_y, cfg.model.sigma_max_y, 1000)}
    shape = (B, 3, IMAGE, IMAGE)
    torch.manual_seed(1000 + rank)

    def make_sampler(p_steps):
        return sampling.get_pc_conditional_sampler(
            sde, shape, sampling.get_predictor("conditional_reverse_diffusion"),
            sampling.get_corrector("conditional_langevin"), SNR, p_steps, 1, continuous=True, denoise=True, eps=EPS)

    # ---- device-resident timing: inputs already in HBM, K graph replays of one PC step ----
    from conditional_score_diffusion_b200.sampling.fused import FusedPCSampler
    fs = FusedPCSampler(model, sde, shape, "reverse_diffusion", "langevin", SNR, PC_STEPS, 1, False, True, True, EPS,
                        conditional=True)
    fs._setup(dev)
    y_dev = torch.rand(*shape, device=dev)
    fs.y.copy_(y_dev)
    launches0 = _lib.lib().csd_launch_count()
    fs.draw_noise = True
    fs.x.copy_(torch.randn(*shape, device=dev) * cfg.model.sigma_max_x)
    fs.step_idx.zero_()
    # one step issued launch by launch (the form the step graph is captured from: the second network evaluation reuses
    # the first one's time embedding) counts this library's launches per PC step
    own_graph, fs.plan.use_graph = fs.plan.use_graph, False
    fs._step()
    fs.plan.use_graph = own_graph
    launches_per_step = _lib.lib().csd_launch_count() - launches0
    graph = fs._graph(draw_noise=True)
    fs.x.copy_(torch.randn(*shape, device=dev) * cfg.model.sigma_max_x)
    fs.step_idx.zero_()
    for _ in range(W):
        graph.replay()
    fs.step_idx.zero_()             # the K timed steps walk the first K entries of the PC-1000 schedule
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(min(K_steps, PC_STEPS)):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k_eff = min(K_steps, PC_STEPS)
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / k_eff
    clock_info = clocks.stop() if rank == 0 else None
    finite = bool(torch.isfinite(fs.x).all().item())

    # ---- end to end through the public API: host y in pinned memory -> samples back on the host ----
    sampler = make_sampler(k_eff)
    y_host = torch.rand(*shape).pin_memory()
    out_host = torch.empty(*shape).pin_memory()
    sampler(model, y_host.to(dev, non_blocking=True))      # builds the plan/graph for this p_steps (untimed)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    samples, _ = sampler(model, y_host.to(dev, non_blocking=True))   # H2D of y; prior drawn on host + H2D inside
    out_host.copy_(samples, non_blocking=True)                       # D2H of the result
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s_per_step = t.item() / k_eff
    bytes_img = 3 * IMAGE * IMAGE * 4 * B
    e2e = {"value": B * world / (e2e_s_per_step * PC_STEPS), "unit": "images/s",
           "h2d_bytes_per_step": 2 * bytes_img / k_eff, "d2h_bytes_per_step": bytes_img / k_eff,
           "note": f"public get_pc_conditional_sampler call with p_steps={k_eff}: y and the host-drawn prior "
                   f"go H2D once per call, samples D2H once per call (bytes amortised over the {k_eff} steps)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tcgen05 conv/GEMM), CUDA events per launch ----
    peaks = measured_peaks()
    rows = conv_profile(fs.plan)
    rows = conv_profile(fs.plan)     # second pass: warm
    # dominant kernel = the persistent transposed tcgen05 conv (conv_halo_tp_kernel); achieved = algorithmic FLOP
    # of its launches / their CUDA-event time, per launch on average; the other tcgen05 kernel is reported beside it
    # (the 3/6-channel output heads also run in that kernel, with 8 of the 128 M rows stored: by construction they
    # cannot approach the tensor roofline, so they are reported beside the >= 32-channel layers, not averaged in)
    tp = [r for r in rows if r[3] and r[2][2] >= 32 and not r[4]]
    heads = [r for r in rows if r[3] and (r[2][2] < 32 or r[4])]
    tap = [r for r in rows if not r[3]]
    heads_ms, heads_fl = sum(r[0] for r in heads), sum(r[1] for r in heads)
    tp_ms, tp_fl = sum(r[0] for r in tp), sum(r[1] for r in tp)
    tap_ms, tap_fl = sum(r[0] for r in tap), sum(r[1] for r in tap)
    top = max(rows, key=lambda r: r[0])
    achieved = tp_fl / (tp_ms * 1e-3) / 1e12
    # per-shape table of the dominant kernel: launches of one forward grouped by (H, W, C_out, K without identity rows)
    shapes = {}
    for ms_, fl_, shp_, _, _ in tp:
        e = shapes.setdefault(tuple(shp_), [0, 0.0, 0.0])
        e[0] += 1
        e[1] += ms_
        e[2] += fl_
    per_shape = [{"h_w_n_k": list(k), "launches": v[0], "ms": round(v[1], 4),
                  "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1),
                  "frac": round(v[2] / (v[1] * 1e-3) / 1e12 / peaks["bf16_tflops"], 3)}
                 for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])]
    roofline = {"bound": "tensor", "kernel": "conv_halo_tp_kernel (persistent tcgen05 implicit-GEMM 3x3 conv, fused "
                                             "GroupNorm+SiLU prologue)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": ncu_traffic("conv_halo_tp_kernel"),
                "traffic_unit": "DRAM bytes per launch (ncu, profiles/ncu_traffic.json)",
                "peak_source": peaks["source"], "launches_per_forward": len(tp),
                "avg_launch_ms": tp_ms / max(1, len(tp)), "algorithmic_gflop_per_launch": tp_fl / 1e9 / max(1, len(tp)),
                "ms_per_forward_in_kernel": tp_ms, "share_of_step": 2 * tp_ms / ms_per_step,
                "top_launch": {"ms": top[0], "tflops": top[1] / (top[0] * 1e-3) / 1e12, "h_w_n_k": top[2]},
                "per_shape": per_shape,
                "other_tcgen05_kernel": {"kernel": "conv_gemm_kernel (per-tap implicit GEMM: <= 20 px levels, 1x1, "
                                                   "stride 2, attention GEMMs)",
                                         "launches_per_forward": len(tap), "ms_per_forward": tap_ms,
                                         "achieved_tflops": tap_fl / (tap_ms * 1e-3) / 1e12,
                                         "share_of_step": 2 * tap_ms / ms_per_step},
                "output_heads_in_same_kernel": {"launches_per_forward": len(heads), "ms_per_forward": heads_ms,
                                                "achieved_tflops": heads_fl / (heads_ms * 1e-3) / 1e12 if heads else None,
                                                "note": "conv3x3 -> 6 channels as a tap-stacked 1x1 convolution (54 of 128 M "
                                                        "rows) + csd_tap_shift_sum"},
                "all_conv_gemm": {"ms_per_forward": tp_ms + tap_ms + heads_ms,
                                  "algorithmic_gflop_per_forward": (tp_fl + tap_fl + heads_fl) / 1e9,
                                  "achieved_tflops": (tp_fl + tap_fl + heads_fl) / ((tp_ms + tap_ms + heads_ms) * 1e-3) / 1e12}}

    plan_gb = fs.plan.pool.nbytes() / 1e9
    # ---- parity of the benchmarked network (same weights, B = 2) against the CPU oracle ----
    parity = network_parity(cfg, model, dev)

    # ---- the reference-precision plan next to the bf16 one: fp32 activations in HBM, tcgen05 kind::tf32 operands ----
    # (context legs - tf32 plan, stock reference on the GPU, other configs - run at N = 1 only: they are rank-0 work and
    #  would only lengthen the multi-GPU runs, whose job is the scaling of `value`)
    tf32 = None
    if not args.no_tf32 and world == 1:
        try:
            tf32 = tf32_line(cfg, model, sde, shape, dev, min(k_eff, 10), B, world)
        except Exception as e:  # noqa: BLE001
            tf32 = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        model.set_precision("bf16")

    # ---- stock PyTorch + cuDNN on the same GPU: the unmodified reference, B = 64 ----
    stock = None
    if not args.no_stock_gpu and world == 1:
        try:
            stock = stock_gpu_steps(cfg, B, 5, 2, dev)
        except Exception as e:  # noqa: BLE001 - the leg is context; the b200 line must still print
            stock = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    # ---- the other BASELINE configs, briefly ----
    extras = None
    if not args.no_extras and world == 1:
        del fs, graph
        torch.cuda.empty_cache()
        extras = other_configs(dev)

    # ---- CPU baseline: the unmodified reference on the host cores, bounded sample ----
    cpu, cpu_kind = reference_cpu_steps(cfg, 2, 3, 1)
    cpu_what = "unmodified reference (baseline/_ref)" if cpu_kind == "reference" else "oracle port"
    cpu_baseline = {"value": cpu["images_per_s"], "unit": "images/s", "cores": cpu["cores"], "kind": cpu_kind,
                    "sample": f"batch 2 of {B}, {cpu['steps_timed']} PC steps after 1 warm-up step, {cpu_what} (torch CPU "
                              f"fp32, {cpu['cores']} threads), extrapolated x1000 steps"}

    line = {
        "metric": METRIC, "value": B * world / (ms_per_step * 1e-3 * PC_STEPS), "unit": "images/s",
        "n_gpus": world, "steps": k_eff, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "celebA_ours_NDV_160 as ncsnpp_paired nf96 (48.5 M params), conditional PC sampler "
                               "(conditional_langevin + conditional_reverse_diffusion, snr 0.15), 160x160",
                   "batch_per_gpu": B, "global_batch": B * world, "pc_steps_per_image": PC_STEPS,
                   "step": "one PC step = 2 score-network evaluations + update kernels over the batch",
                   "parallelism": f"batch sharded over {world} GPU(s), one weight broadcast, no per-step collective",
                   "l2": "activations per step (several GB) exceed the 126 MB L2; no flush needed",
                   "weights": "random init, init_scale=1", "finite_output": finite},
        "e2e": e2e, "gpu_launches": int(launches_per_step * k_eff), "launches_per_step": int(launches_per_step),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clock_info,
        "plan_buffers_gb": plan_gb, "parity": parity,
    }
    if extras is not None:
        line["other_configs"] = extras
    if tf32 is not None:
        line["tf32"] = tf32
    if stock is not None:
        line["stock_gpu"] = stock
        if "ms_per_step" in stock:
            line["vs_stock_gpu"] = stock["ms_per_step"] / ms_per_step
            if tf32 is not None and "ms_per_step" in tf32:
                tf32["vs_stock_gpu"] = stock["ms_per_step"] / tf32["ms_per_step"]
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# training workload (BASELINE configs[3]: edges2shoes_SR3.py as shipped - ddpm_paired_SR3, nf 128, 64 px, batch 50 per
# GPU, SR3 loss with likelihood weighting, Adam 2e-4, clip 1.0, DDP gradient all-reduce). Not the default line.
# ------------------------------------------------------------------------------------------------
TRAIN_METRIC = "training images/sec ddpm_paired_SR3 64px (edges2shoes_SR3), fwd + bwd + clip + Adam + EMA"
TRAIN_BATCH, TRAIN_IMAGE = 50, 64


def train_config():
    from types import SimpleNamespace as NS
    c = NS()
    c.training = NS(continuous=True)
    c.data = NS(image_size=TRAIN_IMAGE, effective_image_size=TRAIN_IMAGE, num_channels=6, centered=False)
    c.model = NS(name="ddpm_paired_SR3", nf=128, ch_mult=(1, 1, 2, 2), num_res_blocks=2, attn_resolutions=(16, 8),
                 dropout=0.1, resamp_with_conv=True, conditional=True, nonlinearity="swish", input_channels=6,
                 output_channels=3, num_scales=1000)
    c.optim = NS(weight_decay=0, optimizer="Adam", lr=2e-4, beta1=0.9, eps=1e-8, warmup=2500, grad_clip=1.0)
    return c


def cpu_oracle_train_steps(cfg, sample_batch, steps, time_budget_s):
    """Oracle port of the training step on host cores: autograd through the CPU restatement + torch Adam."""
    import torch
    from oracle import ddpm as o_ddpm, losses as o_loss, sde as o_sde
    from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    cfg.model.dropout = 0.0
    model = utils.create_model(cfg)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    o = o_ddpm.model_options(cfg)
    hw = cfg.data.image_size
    sx = o_sde.VE(5e-3, math.sqrt(3 * hw * hw), 1000)
    trainable = [v for v in params.values() if v.requires_grad]
    opt = torch.optim.Adam(trainable, lr=2e-4)
    shadow = [v.detach().clone() for v in trainable]
    x, y = torch.rand(sample_batch, 3, hw, hw), torch.rand(sample_batch, 3, hw, hw)
    score = lambda d, t: o_ddpm.forward_paired_sr3(params, o, d["x"], d["y"], t * 999) / sx.sigma(t)[:, None, None, None]
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        t = torch.rand(sample_batch) * (1 - 1e-5) + 1e-5
        loss = o_loss.sr3_loss(score, sx, y, x, t, torch.randn_like(x), True, True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([v for v in params.values() if v.requires_grad], 1.0)
        opt.step()
        with torch.no_grad():
            torch._foreach_lerp_(shadow, trainable, 1.0 - 0.999)
        if i >= 1:
            times.append(time.perf_counter() - t0)
            if sum(times) > time_budget_s:
                break
    s = sum(times) / len(times)
    return {"s_per_step": s, "steps_timed": len(times), "cores": cores, "images_per_s": sample_batch / s}


def run_train(args):
    import torch
    import torch.distributed as dist
    from conditional_score_diffusion_b200 import _lib, distributed as D, losses, optim, sde_lib
    from conditional_score_diffusion_b200.models import ddpm, utils  # noqa: F401
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = train_config()
    if args.impl == "reference":
        if rank == 0:
            r = cpu_oracle_train_steps(cfg, 2, max(1, min(args.steps, 3)), 120.0)
            sample = f"batch 2 of {TRAIN_BATCH}, {r['steps_timed']} steps after 1 warm-up, oracle port (torch CPU fp32 autograd)"
            print(json.dumps({"impl": "reference", "metric": TRAIN_METRIC, "value": r["images_per_s"], "unit": "images/s",
                              "n_gpus": args.gpus, "steps": r["steps_timed"], "warmup": 1,
                              "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": "edges2shoes_SR3 training step, ddpm_paired_SR3 nf128 64x64", "batch_per_step": 2},
                              "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"],
                                               "kind": "port", "sample": sample},
                              "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        quiet_stdout()
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = utils.create_model(cfg).to(dev).train()
    if world > 1:
        D.broadcast_parameters(model, src=0)
        if args.overlap:
            # bucketed gradient all-reduce overlapped with the reverse launch list (distributed.OverlappedGradSync).
            # Opt-in: measured at N = 2 (profiles/ddp_overlap_r2.json) the whole 115 MB all-reduce is 0.28 ms exposed
            # after the backward pass and 0.44 ms when split into 36 spans over 4 segments - NVLink makes the collective
            # too short for the extra launches to pay
            D.enable_gradient_overlap(model, segments=4)
    sde = sde_lib.cVESDE(5e-3, math.sqrt(3 * TRAIN_IMAGE * TRAIN_IMAGE), 1000)
    loss_fn = losses.get_general_sde_loss_fn(sde, train=True, conditional=True, reduce_mean=True, continuous=True,
                                             likelihood_weighting=True)
    if args.torch_optim:     # the reference's own tail: clip_grad_norm_ + optim.Adam + per-tensor EMA (models/ema.py)
        opt = losses.get_optimizer(cfg, model.parameters())
        optimize_fn = losses.optimization_manager(cfg)
        shadow = [p.detach().clone() for p in model.parameters() if p.requires_grad]
    else:                    # fused tail over the flat parameter buffer (optim.FusedAdamEMA: 2 launches)
        opt = optim.FusedAdamEMA(model.parameters(), lr=cfg.optim.lr, betas=(cfg.optim.beta1, 0.999), eps=cfg.optim.eps,
                                 weight_decay=cfg.optim.weight_decay, grad_clip=cfg.optim.grad_clip, ema_decay=0.999,
                                 warmup=cfg.optim.warmup, model=model)
    B, K_steps, W = TRAIN_BATCH, args.steps, max(args.warmup, 3)
    torch.manual_seed(1000 + rank)
    # pinned host batches (a data loader's output): H2D copies are inside the e2e region
    host = [(torch.rand(B, 3, TRAIN_IMAGE, TRAIN_IMAGE).pin_memory(), torch.rand(B, 3, TRAIN_IMAGE, TRAIN_IMAGE).pin_memory())
            for _ in range(4)]
    xd, yd = host[0][0].to(dev), host[0][1].to(dev)

    def step(i, x, y):
        opt.zero_grad()
        loss = loss_fn(model, (y, x))
        loss.backward()
        if world > 1:
            D.allreduce_gradients(model, average=True)
        if args.torch_optim:
            optimize_fn(opt, model.parameters(), step=i + 1)
            with torch.no_grad():
                ps = [p for p in model.parameters() if p.requires_grad]
                torch._foreach_lerp_(shadow, ps, 1.0 - 0.999)
        else:
            opt.step()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i, xd, yd)
    clocks = ClockSampler(local_rank)
    sync()
    n0 = _lib.lib().csd_launch_count()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K_steps):
        loss = step(W + i, xd, yd)
    e1.record()
    sync()
    ck = clocks.stop()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / K_steps, dev)
    # e2e: batch from pinned host memory every step, loss read back every step
    sync()
    t0 = time.perf_counter()
    for i in range(K_steps):
        hx, hy = host[i % len(host)]
        loss = step(W + K_steps + i, hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True))
        loss_host = loss.item()
    sync()
    e2e_ms = D.max_over_ranks((time.perf_counter() - t0) / K_steps * 1e3, dev)
    plan = next(iter(model._engine.train_plans.values()))
    launches_per_step = len(plan.rec.ops) + len(plan.bwd.ops)
    if rank == 0:
        peaks = measured_peaks()
        line = {
            "metric": TRAIN_METRIC, "value": B * world / ms * 1e3, "unit": "images/s", "n_gpus": world, "steps": K_steps,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "edges2shoes_SR3.py as shipped: ddpm_paired_SR3 nf128 ch_mult (1,1,2,2) attn (16,8), 64x64, "
                                   "dropout 0.1, SR3 loss (likelihood weighting), Adam 2e-4, clip 1.0, EMA 0.999",
                       "optimizer": "torch Adam + foreach EMA" if args.torch_optim else "optim.FusedAdamEMA (2 launches)",
                       "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": (f"data parallel over {world} GPU(s), gradient all-reduce of the flat fp32 buffer in 4 "
                                       f"segments overlapped with the backward pass" if (world > 1 and args.overlap)
                                       else f"data parallel over {world} GPU(s), one all-reduce of the flat fp32 gradient "
                                            f"buffer after the backward pass"),
                       "l2": "activations + pixel-major copies per step (GBs) exceed the 126 MB L2", "finite_loss": bool(math.isfinite(loss_host))},
            "e2e": {"value": B * world / e2e_ms * 1e3, "unit": "images/s", "h2d_bytes_per_step": 2 * B * 3 * TRAIN_IMAGE * TRAIN_IMAGE * 4,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * K_steps, "launches_per_step": launches_per_step,
            "clocks": ck, "peaks": peaks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # 200 of the 1000 identical steps
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_eager_gpu"])
    ap.add_argument("--overlap", action="store_true", help="train workload: segment-wise gradient all-reduce overlapped "
                    "with the backward pass instead of one all-reduce after it")
    ap.add_argument("--no-extras", action="store_true", help="skip the short records of the other BASELINE configs")
    ap.add_argument("--no-tf32", action="store_true", help="skip the reference-precision (tf32 plan) context line")
    ap.add_argument("--no-stock-gpu", action="store_true", help="skip the stock PyTorch + cuDNN leg (the unmodified "
                    "reference on the same GPU)")
    ap.add_argument("--torch-optim", action="store_true", help="train workload: torch Adam instead of the fused optimizer")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample = the headline PC-1000 sampling line (default); train = BASELINE configs[3] training step")
    args = ap.parse_args()
    if args.workload == "train":
        if args.steps == 200:
            args.steps = 20
        run_train(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_eager_gpu":
        run_torch_eager_gpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
