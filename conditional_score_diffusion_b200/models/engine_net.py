"""Base class of the engine-backed score networks (NCSN++ and DDPM families).

The reference's networks subclass pl.LightningModule (models/ncsnpp.py:23,40; models/ddpm.py:24,81) and run
eager ATen ops in forward(). Here a network is a parameter container with the reference's module order and names;
its arithmetic is the planned CUDA launch list of `engine.NetEngine`. Inference only: calling forward with autograd
enabled on parameters that require grad raises (there is no silent PyTorch fallback).
"""
import torch
import torch.nn as nn

try:  # the reference subclasses pl.LightningModule (models/ncsnpp.py:23,40)
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - Lightning is not installed in the build image
    class _Base(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def save_hyperparameters(self, *args, **kwargs):
            pass

        def log(self, *args, **kwargs):
            pass


class EngineNet(_Base):
    """forward() plumbing shared by the engine-backed networks: copy inputs into the plan's static buffers, launch
    (CUDA graph after the first call), clone the outputs."""

    def _check_inference(self, *tensors):
        if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors if torch.is_tensor(t))
                                        or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError(
                "score-network backward (training / likelihood divergence) is not implemented by the B200 engine yet; "
                "call the network under torch.no_grad(). No PyTorch fallback is provided on purpose.")
        if self.training and self.config.model.dropout > 0:
            raise NotImplementedError("dropout (train mode) is not implemented by the B200 engine; use .eval()")

    def _run(self, x0, x1, time_cond, scale0=None, scale1=None, clone=True):
        """x0 [B,c0,H,W] (+ optional x1 [B,c1,H,W], channel-concatenated after x0), time_cond [B]."""
        self._check_inference(x0, x1, time_cond)
        if x0.device.type != "cuda":
            raise RuntimeError("the score network runs on CUDA tensors only (libcsd_b200 has no CPU path)")
        eng = self._engine
        eng.ensure_packed(x0.device)
        b, c0, h, w = x0.shape
        c1 = x1.shape[1] if x1 is not None else 0
        if c0 + c1 != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {c0 + c1}")
        plan = eng.plan(b, h, w, c0, c1)
        plan.in0.copy_(x0)
        if x1 is not None:
            plan.in1.copy_(x1)
        plan.labels.copy_(time_cond.to(torch.float32))
        if scale0 is not None:
            plan.row_scale.copy_(scale0)
        else:
            plan.row_scale.fill_(1.0)
        if len(plan.outputs()) > 1:
            if scale1 is not None:
                plan.row_scale1.copy_(scale1)
            else:
                plan.row_scale1.fill_(1.0)
        plan.launch()
        outs = plan.outputs()
        return [o.clone() for o in outs] if clone else outs

    def forward_scaled(self, x, time_cond, inv_std):
        """forward(x, time_cond) * inv_std[:, None, None, None], fused into the output kernel."""
        return self._run(x, None, time_cond, scale0=inv_std)[0]
