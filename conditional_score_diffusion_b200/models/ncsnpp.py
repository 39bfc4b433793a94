"""NCSN++ score network (reference: models/ncsnpp.py:39-449).

Same registry names ('ncsnpp', 'ncsnpp_paired', 'ncsnpp_2xSR', 'ncsnpp_KxSR'), constructor
(`NCSNpp(config)`), attributes read by callers (`embedding_type`, `device`, `all_modules`) and
state-dict keys (`all_modules.<i>.<Sub>.<param>`: the positional order of `all_modules` is the
checkpoint contract, models/ncsnpp.py:236-382). `forward(x, time_cond)` returns a fresh NCHW fp32
tensor like the reference; internally it runs the planned CUDA launch list of `engine.NetEngine`.

With autograd enabled the network is one autograd node backed by the planned backward pass
(engine_train.TrainPlan); there is no PyTorch fallback.
"""
import functools

import torch
import torch.nn as nn

from . import layers, layerspp, utils
from ..engine import NetEngine

from .engine_net import EngineNet as _Base

ResnetBlockDDPM = layerspp.ResnetBlockDDPMpp
ResnetBlockBigGAN = layerspp.ResnetBlockBigGANpp
Combine = layerspp.Combine
conv3x3 = layerspp.conv3x3
conv1x1 = layerspp.conv1x1
get_act = layers.get_act
default_initializer = layers.default_init


@utils.register_model(name="ncsnpp")
class NCSNpp(_Base):
    """NCSN++ model."""

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
        resamp_with_conv = m.resamp_with_conv
        self.num_resolutions = num_resolutions = len(ch_mult)
        self.all_resolutions = all_resolutions = [config.data.effective_image_size // (2 ** i)
                                                  for i in range(num_resolutions)]
        self.conditional = conditional = m.conditional
        self.fir = fir = m.fir
        self.fir_kernel = fir_kernel = m.fir_kernel
        self.skip_rescale = skip_rescale = m.skip_rescale
        self.resblock_type = resblock_type = m.resblock_type.lower()
        self.progressive = progressive = m.progressive.lower()
        self.progressive_input = progressive_input = m.progressive_input.lower()
        self.embedding_type = embedding_type = m.embedding_type.lower()
        self.centered = config.data.centered
        init_scale = m.init_scale
        assert progressive in ["none", "output_skip", "residual"]
        assert progressive_input in ["none", "input_skip", "residual"]
        assert embedding_type in ["fourier", "positional"]
        self.combine_method = combine_method = m.progressive_combine.lower()
        combiner = functools.partial(Combine, method=combine_method)

        modules = []
        if embedding_type == "fourier":
            assert config.training.continuous, "Fourier features are only used for continuous training."
            modules.append(layerspp.GaussianFourierProjection(embedding_size=nf, scale=m.fourier_scale))
            embed_dim = 2 * nf
        else:
            embed_dim = nf
        if conditional:
            for d_in in (embed_dim, nf * 4):
                lin = nn.Linear(d_in, nf * 4)
                lin.weight.data = default_initializer()(lin.weight.shape)
                nn.init.zeros_(lin.bias)
                modules.append(lin)

        AttnBlock = functools.partial(layerspp.AttnBlockpp, init_scale=init_scale, skip_rescale=skip_rescale)
        Upsample = functools.partial(layerspp.Upsample, with_conv=resamp_with_conv, fir=fir, fir_kernel=fir_kernel)
        if progressive == "output_skip":
            self.pyramid_upsample = layerspp.Upsample(fir=fir, fir_kernel=fir_kernel, with_conv=False)
        elif progressive == "residual":
            pyramid_upsample = functools.partial(layerspp.Upsample, fir=fir, fir_kernel=fir_kernel, with_conv=True)
        Downsample = functools.partial(layerspp.Downsample, with_conv=resamp_with_conv, fir=fir, fir_kernel=fir_kernel)
        if progressive_input == "input_skip":
            self.pyramid_downsample = layerspp.Downsample(fir=fir, fir_kernel=fir_kernel, with_conv=False)
        elif progressive_input == "residual":
            pyramid_downsample = functools.partial(layerspp.Downsample, fir=fir, fir_kernel=fir_kernel,
                                                   with_conv=True)
        if resblock_type == "ddpm":
            ResnetBlock = functools.partial(ResnetBlockDDPM, act=act, dropout=dropout, init_scale=init_scale,
                                            skip_rescale=skip_rescale, temb_dim=nf * 4)
        elif resblock_type == "biggan":
            ResnetBlock = functools.partial(ResnetBlockBigGAN, act=act, dropout=dropout, fir=fir,
                                            fir_kernel=fir_kernel, init_scale=init_scale,
                                            skip_rescale=skip_rescale, temb_dim=nf * 4)
        else:
            raise ValueError(f"resblock type {resblock_type} unrecognized.")

        # ---- downsampling path (models/ncsnpp.py:140-175) ----
        channels = config.data.num_channels
        if progressive_input != "none":
            input_pyramid_ch = channels
        modules.append(conv3x3(channels, nf))
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
                if resblock_type == "ddpm":
                    modules.append(Downsample(in_ch=in_ch))
                else:
                    modules.append(ResnetBlock(down=True, in_ch=in_ch))
                if progressive_input == "input_skip":
                    modules.append(combiner(dim1=input_pyramid_ch, dim2=in_ch))
                    if combine_method == "cat":
                        in_ch *= 2
                elif progressive_input == "residual":
                    modules.append(pyramid_downsample(in_ch=input_pyramid_ch, out_ch=in_ch))
                    input_pyramid_ch = in_ch
                hs_c.append(in_ch)

        # ---- middle (ncsnpp.py:177-180) ----
        in_ch = hs_c[-1]
        modules.append(ResnetBlock(in_ch=in_ch))
        modules.append(AttnBlock(channels=in_ch))
        modules.append(ResnetBlock(in_ch=in_ch))

        # ---- upsampling path (ncsnpp.py:182-227) ----
        pyramid_ch = 0
        for i_level in reversed(range(num_resolutions)):
            for _ in range(num_res_blocks + 1):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch + hs_c.pop(), out_ch=out_ch))
                in_ch = out_ch
            if all_resolutions[i_level] in attn_resolutions:
                modules.append(AttnBlock(channels=in_ch))
            if progressive != "none":
                if i_level == num_resolutions - 1:
                    modules.append(nn.GroupNorm(num_groups=min(in_ch // 4, 32), num_channels=in_ch, eps=1e-6))
                    if progressive == "output_skip":
                        modules.append(conv3x3(in_ch, channels, init_scale=init_scale))
                        pyramid_ch = channels
                    else:
                        modules.append(conv3x3(in_ch, in_ch, bias=True))
                        pyramid_ch = in_ch
                else:
                    if progressive == "output_skip":
                        modules.append(nn.GroupNorm(num_groups=min(in_ch // 4, 32), num_channels=in_ch, eps=1e-6))
                        modules.append(conv3x3(in_ch, channels, bias=True, init_scale=init_scale))
                        pyramid_ch = channels
                    else:
                        modules.append(pyramid_upsample(in_ch=pyramid_ch, out_ch=in_ch))
                        pyramid_ch = in_ch
            if i_level != 0:
                if resblock_type == "ddpm":
                    modules.append(Upsample(in_ch=in_ch))
                else:
                    modules.append(ResnetBlock(in_ch=in_ch, up=True))
        assert not hs_c
        if progressive != "output_skip":
            modules.append(nn.GroupNorm(num_groups=min(in_ch // 4, 32), num_channels=in_ch, eps=1e-6))
            modules.append(conv3x3(in_ch, channels, init_scale=init_scale))
        self.all_modules = nn.ModuleList(modules)
        self.arch = "ncsnpp"
        self.in_channels = self.out_channels = channels
        self._engine = NetEngine(self)

    def forward(self, x, time_cond):
        return self._run(x, None, time_cond)[0]

    def forward_scaled(self, x, time_cond, inv_std):
        """forward(x, time_cond) * inv_std[:, None, None, None], fused into the output kernel."""
        return self._run(x, None, time_cond, scale0=inv_std)[0]


@utils.register_model(name="ncsnpp_paired")
class NCSNpp_paired(NCSNpp):
    """models/ncsnpp.py:390-401: cat(x, y) on channels, output split back into {'x', 'y'}."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)

    def forward(self, input_dict, labels):
        ox, oy = self._run(input_dict["x"], input_dict["y"], labels)
        return {"x": ox, "y": oy}

    def forward_scaled(self, input_dict, labels, inv_std):
        ox, oy = self._run(input_dict["x"], input_dict["y"], labels, scale0=inv_std["x"], scale1=inv_std["y"])
        return {"x": ox, "y": oy}


from .engine_net import ResizeSRMixin, SqueezeSRMixin, SqueezeBlock  # noqa: E402,F401


@utils.register_model(name="ncsnpp_2xSR")
class NCSNpp_2xSR(SqueezeSRMixin, NCSNpp):
    """models/ncsnpp.py:418-433."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)
        self._init_sr(config)


@utils.register_model(name="ncsnpp_KxSR")
class NCSNpp_KxSR(ResizeSRMixin, NCSNpp):
    """models/ncsnpp.py:435-449."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config)
        self._init_sr(config)
