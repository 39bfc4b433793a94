"""DDPM U-Net score network (reference: models/ddpm.py:80-213, 275-298).

This is the network most shipped configs select (`model.name = 'ddpm_paired'`). Same registry
names ('ddpm', 'ddpm_paired', 'ddpm_paired_SR3', 'ddpm_2xSR', 'ddpm_KxSR'), constructor (`DDPM(config)`), `all_modules` order and state-dict
keys as the reference; forward() runs the planned CUDA launch list of `engine.NetEngine` (the same kernels as
NCSN++: GroupNorm(32)+SiLU fused into the 3x3 convolutions where they run in the transposed mode, NIN shortcut as an
extra K segment, nearest-neighbour x2 as a 2-tap FIR, pad(0,1,0,1)+stride-2 conv through TMA zero fill).

Not carried over: 'ddpm_multi_speed_haar' (broken in the reference: un-imported InvertibleDownsampling2D,
models/ddpm.py:219). With autograd enabled the network is differentiable (engine_train.TrainPlan).
"""
import functools

import torch.nn as nn

from . import layers, utils
from ..engine import NetEngine
from .engine_net import EngineNet

ResnetBlockDDPM = layers.ResnetBlockDDPM
Upsample = layers.Upsample
Downsample = layers.Downsample
conv3x3 = layers.ddpm_conv3x3
get_act = layers.get_act
default_initializer = layers.default_init


@utils.register_model(name="ddpm")
class DDPM(EngineNet):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.act = act = get_act(config)
        m = config.model
        self.nf = nf = m.nf
        ch_mult = m.ch_mult
        self.num_res_blocks = num_res_blocks = m.num_res_blocks
        self.attn_resolutions = attn_resolutions = m.attn_resolutions
        dropout = m.dropout
        self.resamp_with_conv = resamp_with_conv = m.resamp_with_conv
        self.num_resolutions = num_resolutions = len(ch_mult)
        self.all_resolutions = all_resolutions = [config.data.effective_image_size // (2 ** i)
                                                  for i in range(num_resolutions)]
        AttnBlock = functools.partial(layers.AttnBlock)
        self.conditional = conditional = m.conditional
        if not conditional:
            # the reference's unconditional branch never defines `modules` (models/ddpm.py:98-112)
            raise NotImplementedError("DDPM(conditional=False) is broken in the reference and not supported")
        ResnetBlock = functools.partial(ResnetBlockDDPM, act=act, temb_dim=4 * nf, dropout=dropout)
        modules = [nn.Linear(nf, nf * 4)]
        modules[0].weight.data = default_initializer()(modules[0].weight.data.shape)
        nn.init.zeros_(modules[0].bias)
        modules.append(nn.Linear(nf * 4, nf * 4))
        modules[1].weight.data = default_initializer()(modules[1].weight.data.shape)
        nn.init.zeros_(modules[1].bias)

        self.centered = config.data.centered
        self.in_channels = input_channels = m.input_channels
        self.out_channels = output_channels = m.output_channels

        modules.append(conv3x3(input_channels, nf))
        hs_c = [nf]
        in_ch = nf
        for i_level in range(num_resolutions):
            for _ in range(num_res_blocks):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch, out_ch=out_ch))
                in_ch = out_ch
                if all_resolutions[i_level] in attn_resolutions:
                    modules.append(AttnBlock(channels=in_ch))
                hs_c.append(in_ch)
            if i_level != num_resolutions - 1:
                modules.append(Downsample(channels=in_ch, with_conv=resamp_with_conv))
                hs_c.append(in_ch)

        in_ch = hs_c[-1]
        modules.append(ResnetBlock(in_ch=in_ch))
        modules.append(AttnBlock(channels=in_ch))
        modules.append(ResnetBlock(in_ch=in_ch))

        for i_level in reversed(range(num_resolutions)):
            for _ in range(num_res_blocks + 1):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch + hs_c.pop(), out_ch=out_ch))
                in_ch = out_ch
            if all_resolutions[i_level] in attn_resolutions:
                modules.append(AttnBlock(channels=in_ch))
            if i_level != 0:
                modules.append(Upsample(channels=in_ch, with_conv=resamp_with_conv))

        assert not hs_c
        modules.append(nn.GroupNorm(num_channels=in_ch, num_groups=32, eps=1e-6))
        modules.append(conv3x3(in_ch, output_channels, init_scale=0.0))
        self.all_modules = nn.ModuleList(modules)
        self.arch = "ddpm"
        self.embedding_type = "positional"      # read by models.utils.get_score_fn for unconditional VE labels
        self._engine = NetEngine(self)

    def forward(self, x, labels):
        return self._run(x, None, labels)[0]


@utils.register_model(name="ddpm_paired_SR3")
class DDPM_paired_SR3(DDPM):
    """models/ddpm.py:275-285: cat(x, y) in, score of x out (output_channels = x channels)."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)

    def forward(self, input_dict, labels):
        return self._run(input_dict["x"], input_dict["y"], labels)[0]

    def forward_scaled(self, input_dict, labels, inv_std):
        return self._run(input_dict["x"], input_dict["y"], labels, scale0=inv_std)[0]


@utils.register_model(name="ddpm_paired")
class DDPM_paired(DDPM):
    """models/ddpm.py:287-298: cat(x, y) in, output split back into {'x', 'y'}."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)

    def forward(self, input_dict, labels):
        ox, oy = self._run(input_dict["x"], input_dict["y"], labels)
        return {"x": ox, "y": oy}

    def forward_scaled(self, input_dict, labels, inv_std):
        ox, oy = self._run(input_dict["x"], input_dict["y"], labels, scale0=inv_std["x"], scale1=inv_std["y"])
        return {"x": ox, "y": oy}


from .engine_net import ResizeSRMixin, SqueezeSRMixin, SqueezeBlock  # noqa: E402,F401


@utils.register_model(name="ddpm_2xSR")
class DDPM_2xSR(SqueezeSRMixin, DDPM):
    """models/ddpm.py:300-314."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)
        self._init_sr(config)


@utils.register_model(name="ddpm_KxSR")
class ddpm_KxSR(ResizeSRMixin, DDPM):
    """models/ddpm.py:316-331."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)
        self._init_sr(config)
