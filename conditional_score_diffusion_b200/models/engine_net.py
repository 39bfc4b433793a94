"""Base class of the engine-backed score networks (NCSN++ and DDPM families).

The reference's networks subclass pl.LightningModule (models/ncsnpp.py:23,40; models/ddpm.py:24,81) and run
eager ATen ops in forward(). Here a network is a parameter container with the reference's module order and names;
its arithmetic is the planned CUDA launch list of `engine.NetEngine`; with autograd enabled the network is one
autograd node whose backward is the planned reverse launch list of `engine_train.TrainPlan` (no PyTorch fallback).
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


class _NetFunction(torch.autograd.Function):
    """The whole score network as ONE autograd node: forward = the training plan's launch list, backward = its
    reverse launch list (engine_train.TrainPlan). Parameters are passed as inputs so `loss.backward()` fills their
    `.grad` exactly as it does for the reference's eager module (losses.py:345-407)."""

    @staticmethod
    def forward(ctx, net, plan, x0, x1, time_cond, scale0, scale1, *params):
        net._load_inputs(plan, x0, x1, time_cond, scale0, scale1)
        if plan.dropout_p > 0:
            plan.dropout_seed.random_()       # torch's CUDA generator: torch.manual_seed makes the masks reproducible
        plan.launch()
        ctx.plan = plan
        ctx.has_x1 = x1 is not None
        ctx.n_params = len(params)
        plan.forward_serial = getattr(plan, "forward_serial", 0) + 1
        ctx.serial = plan.forward_serial
        return tuple(o.clone() for o in plan.outputs())

    @staticmethod
    def backward(ctx, *gouts):
        plan = ctx.plan
        if ctx.serial != plan.forward_serial:
            raise RuntimeError("the score network was evaluated again before backward(): its stored activations were "
                               "overwritten (one live autograd graph per network and batch shape)")
        for buf, g in zip(plan.gouts, gouts):
            if g is None:
                buf.zero_()
            else:
                buf.copy_(g)
        plan.run_backward()
        gx0 = gx1 = None
        if plan.want_input:
            gx0 = plan.gin[0].clone()
            if ctx.has_x1:
                gx1 = plan.gin[1].clone()
        pg = plan.param_grads() if plan.want_params else [None] * ctx.n_params
        return (None, None, gx0, gx1, None, None, None) + tuple(pg)


class EngineNet(_Base):
    """forward() plumbing shared by the engine-backed networks: copy inputs into the plan's static buffers, launch
    (CUDA graph after the first call), clone the outputs. With autograd enabled the call goes through
    `_NetFunction` (planned backward pass); otherwise through the inference plan."""

    def _needs_grad(self, *tensors):
        if not torch.is_grad_enabled():
            return False, False
        want_in = any(t.requires_grad for t in tensors if torch.is_tensor(t))
        want_p = any(p.requires_grad for p in self.parameters())
        return want_p, want_in

    def _load_inputs(self, plan, x0, x1, time_cond, scale0, scale1):
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

    def _run(self, x0, x1, time_cond, scale0=None, scale1=None, clone=True):
        """x0 [B,c0,H,W] (+ optional x1 [B,c1,H,W], channel-concatenated after x0), time_cond [B]."""
        if x0.device.type != "cuda":
            raise RuntimeError("the score network runs on CUDA tensors only (libcsd_b200 has no CPU path)")
        want_p, want_in = self._needs_grad(x0, x1)
        p_drop = float(self.config.model.dropout) if self.training else 0.0
        if p_drop > 0 and not (want_p or want_in):
            raise NotImplementedError("train-mode dropout without autograd (torch.no_grad() on a .train() network) is "
                                      "not planned by the B200 engine; call .eval() for inference")
        eng = self._engine
        # inference entry: weights may have been swapped through `.data` (EMA copy_to/restore) - always re-pack
        eng.ensure_packed(x0.device, force_refresh=not (want_p or want_in))
        b, c0, h, w = x0.shape
        c1 = x1.shape[1] if x1 is not None else 0
        if c0 + c1 != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {c0 + c1}")
        if want_p or want_in:
            plan = eng.train_plan(b, h, w, c0, c1, want_params=want_p, want_input=want_in, dropout=p_drop)
            params = [p for p in self.parameters()]
            outs = _NetFunction.apply(self, plan, x0, x1, time_cond, scale0, scale1, *params)
            return list(outs)
        plan = eng.plan(b, h, w, c0, c1)
        self._load_inputs(plan, x0, x1, time_cond, scale0, scale1)
        plan.launch()
        outs = plan.outputs()
        return [o.clone() for o in outs] if clone else outs

    def forward_scaled(self, x, time_cond, inv_std):
        """forward(x, time_cond) * inv_std[:, None, None, None], fused into the output kernel."""
        return self._run(x, None, time_cond, scale0=inv_std)[0]

    def set_precision(self, precision):
        """'bf16' (default: bf16 activations + bf16 tensor-core operands, the fast plan) or 'tf32' (fp32 activations in
        HBM + tf32 tensor-core operands: the reference's own precision class - fp32 storage everywhere,
        sampling/unconditional.py:206, cuDNN TF32 convolutions under PyTorch's defaults). Inference plans only; cached
        samplers re-plan on their next call. Returns self."""
        self._engine.set_precision(precision)
        return self

    @property
    def precision(self):
        return self._engine.precision


class SqueezeBlock(nn.Module):
    """models/ncsnpp.py:403-416: space-to-depth by 2 (channel order c*4 + dy*2 + dx) and its inverse. A pure
    permutation of the data: layout plumbing around the engine call, done with tensor views."""

    def forward(self, z, reverse=False):
        B, C, H, W = z.shape
        if not reverse:
            z = z.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4)
            return z.reshape(B, 4 * C, H // 2, W // 2)
        z = z.reshape(B, C // 4, 2, 2, H, W).permute(0, 1, 4, 2, 5, 3)
        return z.reshape(B, C // 4, H * 2, W * 2)


class SqueezeSRMixin:
    """forward of the *_2xSR wrappers (models/ncsnpp.py:418-433, models/ddpm.py:300-314): x is squeezed to the
    resolution of y, concatenated, and the x part of the output is un-squeezed."""

    def _init_sr(self, config):
        self.squeeze_block = SqueezeBlock()

    def _sr_forward(self, input_dict, labels, inv_std=None):
        x = self.squeeze_block(input_dict["x"]).contiguous()
        y = input_dict["y"]
        s0 = inv_std["x"] if inv_std is not None else None
        s1 = inv_std["y"] if inv_std is not None else None
        ox, oy = self._run(x, y, labels, scale0=s0, scale1=s1)
        return {"x": self.squeeze_block(ox, reverse=True), "y": oy}

    def forward(self, input_dict, labels):
        return self._sr_forward(input_dict, labels)

    def forward_scaled(self, input_dict, labels, inv_std):
        return self._sr_forward(input_dict, labels, inv_std)


class ResizeSRMixin:
    """forward of the *_KxSR wrappers (models/ncsnpp.py:435-449, models/ddpm.py:316-331): y is bilinearly resized to
    the target resolution before the network and its score back to the low resolution after it, with the same
    torchvision Resize transforms the reference constructs (pre/post-processing of the condition, not network work)."""

    def _init_sr(self, config):
        from torchvision.transforms import Resize
        from torchvision.transforms.functional import InterpolationMode
        self.resize_to_GT = Resize(config.data.target_resolution, interpolation=InterpolationMode.BILINEAR)
        self.resize_to_LQ = Resize(config.data.target_resolution // config.data.scale,
                                   interpolation=InterpolationMode.BILINEAR)

    def _sr_forward(self, input_dict, labels, inv_std=None):
        x = input_dict["x"]
        y = self.resize_to_GT(input_dict["y"]).contiguous()
        s0 = inv_std["x"] if inv_std is not None else None
        s1 = inv_std["y"] if inv_std is not None else None
        ox, oy = self._run(x, y, labels, scale0=s0, scale1=s1)
        return {"x": ox, "y": self.resize_to_LQ(oy)}

    def forward(self, input_dict, labels):
        return self._sr_forward(input_dict, labels)

    def forward_scaled(self, input_dict, labels, inv_std):
        return self._sr_forward(input_dict, labels, inv_std)
