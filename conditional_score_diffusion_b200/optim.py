"""Fused training-step tail: gradient clipping + Adam + EMA over ONE flat fp32 parameter buffer.

The reference's step is `clip_grad_norm_` + `optim.Adam.step()` (losses.py:38-52, built by get_optimizer :26-35) followed
by `ExponentialMovingAverage.update` (models/ema.py:64-93, called from lightning_callbacks/callbacks.py:125-126): three
passes of per-tensor kernels over ~600 parameter tensors. `FusedAdamEMA` re-homes the parameters as views of one flat
buffer (same values, same nn.Parameter objects, checkpoints unchanged) and runs the whole tail as two launches
(csd_sumsq_f32, csd_fused_adam_ema_f32). The arithmetic is torch.optim.Adam's (no amsgrad, L2 weight decay) and the
reference EMA's (decay capped by (1 + n) / (10 + n)); `losses.get_optimizer` keeps returning torch's Adam - this class
is opt-in.
"""
import torch

from . import kernels as K


class FusedAdamEMA:
    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_clip=-1.0, ema_decay=None,
                 warmup=0, use_num_updates=True, model=None):
        self.model = model
        self.params = list(params)
        if not self.params:
            raise ValueError("FusedAdamEMA got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("FusedAdamEMA needs fp32 CUDA parameters on one device (libcsd_b200 has no CPU path)")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.grad_clip, self.ema_decay, self.warmup, self.use_num_updates = grad_clip, ema_decay, warmup, use_num_updates
        self.param_groups = [{"lr": lr, "params": self.params}]      # what lr schedulers / warm-up code touch
        # same layout rule as engine_train.TrainPlan.gflat: parameters() order, each padded to 4 elements
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.ema = self.flat.clone() if ema_decay is not None else None
        self.gflat = torch.zeros_like(self.flat)
        self.gnorm_sq = K.reduce_workspace(dev)     # [0] = squared gradient norm (deterministic device-wide sum)
        self.step_count = 0
        self.num_updates = 0
        self._stored = None

    # -- gradients ----------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def _flat_grads(self):
        """The gradients as one flat tensor in this optimizer's layout: zero-copy when autograd handed back views of one
        buffer with the same offsets (what the engine's backward pass produces), gathered otherwise."""
        first = next((i for i, p in enumerate(self.params) if p.grad is not None), None)
        if first is None:
            raise RuntimeError("FusedAdamEMA.step() called without gradients")
        g0 = self.params[first].grad
        base = g0.untyped_storage().data_ptr()
        start = g0.storage_offset() - self.offsets[first]
        ok = start >= 0
        for p, o in zip(self.params, self.offsets):
            g = p.grad
            if g is None:
                if p.requires_grad:
                    ok = False
                continue
            if (g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base
                    or g.storage_offset() - o != start):
                ok = False
                break
        if ok and g0.untyped_storage().nbytes() >= 4 * (start + self.flat.numel()):
            return torch.empty(0, dtype=torch.float32, device=g0.device).set_(g0.untyped_storage(), start,
                                                                              (self.flat.numel(),), (1,))
        self.gflat.zero_()
        for p, o in zip(self.params, self.offsets):
            if p.grad is not None:
                self.gflat[o:o + p.numel()].view(p.shape).copy_(p.grad)
        return self.gflat

    # -- step -----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self):
        """One optimizer + EMA step. Equivalent reference sequence: optimize_fn(optimizer, params, step) (losses.py:38-52)
        then ema.update(params) (losses.py:393)."""
        g = self._flat_grads()
        # warm-up as in losses.py:46-48: lr * min(step / warmup, 1) with `step` counted from 0 (the first update runs
        # at lr = 0). param_groups[0]['lr'] is the BASE rate an external scheduler may change; it is scaled, not replaced
        lr = self.param_groups[0]["lr"]
        if self.warmup > 0:
            lr = lr * min(self.step_count / self.warmup, 1.0)
        self.step_count += 1
        b1, b2 = self.betas
        if self.grad_clip >= 0:
            K.sumsq(g, self.gnorm_sq)
        decay = 0.0
        if self.ema is not None:
            decay = self.ema_decay
            if self.use_num_updates:
                self.num_updates += 1
                decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        K.fused_adam_ema(self.flat, g, self.m, self.v, self.ema, lr, b1, b2, self.eps, self.weight_decay,
                         1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count, self.grad_clip, self.gnorm_sq, decay)
        self._mark_changed()

    def _mark_changed(self):
        """The kernels wrote the parameters behind autograd's back: tell the engine (its packed bf16 operands are keyed
        by the parameters' version counters) or, without a model handle, bump the counters with a no-op in-place add."""
        if self.model is not None and hasattr(self.model, "_engine"):
            self.model._engine.invalidate()
        else:
            torch._foreach_add_(self.params, 0.0)

    # -- EMA access (models/ema.py:95-140) -----------------------------------------------------------------------------
    def ema_parameters(self):
        if self.ema is None:
            raise RuntimeError("EMA is disabled (ema_decay=None)")
        return [self.ema[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def ema_store(self):
        self._stored = self.flat.clone()

    def ema_copy_to(self):
        self.flat.copy_(self.ema)
        self._mark_changed()

    def ema_restore(self):
        self.flat.copy_(self._stored)
        self._mark_changed()

    def state_dict(self):
        return {"step": self.step_count, "num_updates": self.num_updates, "m": self.m, "v": self.v, "ema": self.ema,
                "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        self.step_count, self.num_updates = sd["step"], sd["num_updates"]
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"])
        if self.ema is not None and sd.get("ema") is not None:
            self.ema.copy_(sd["ema"])
        self.param_groups[0]["lr"] = sd["lr"]
