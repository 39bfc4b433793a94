"""Oracle restatement of the DDPM U-Net score network forward pass (CPU, fp32, functional).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows models/ddpm.py:80-213 (module order and forward), models/layers.py:567-675 (AttnBlock, Upsample,
Downsample, ResnetBlockDDPM) and :524-538 (timestep embedding) over a flat parameter dict whose keys are the
reference state-dict names (`all_modules.<i>.<Sub>.<param>`). GroupNorm always has 32 groups here (the NCSN++
layers use min(C//4, 32)).
"""
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from .ncsnpp import _nin, _silu, timestep_embedding


def model_options(config):
    m, d = config.model, config.data
    return SimpleNamespace(nf=m.nf, ch_mult=tuple(m.ch_mult), num_res_blocks=m.num_res_blocks,
                           attn_resolutions=tuple(m.attn_resolutions), resamp_with_conv=m.resamp_with_conv,
                           image_size=d.effective_image_size, centered=d.centered,
                           input_channels=m.input_channels, output_channels=m.output_channels)


def _gn32(params, prefix, x):
    return F.group_norm(x, 32, params[prefix + ".weight"], params[prefix + ".bias"], eps=1e-6)


def _conv(params, prefix, x, padding, stride=1):
    return F.conv2d(x, params[prefix + ".weight"], params[prefix + ".bias"], padding=padding, stride=stride)


def resblock(params, idx, x, temb):
    """ResnetBlockDDPM.forward (models/layers.py:656-675); dropout is the identity in eval mode."""
    pre = f"all_modules.{idx}"
    h = _silu(_gn32(params, pre + ".GroupNorm_0", x))
    h = _conv(params, pre + ".Conv_0", h, 1)
    h = h + F.linear(_silu(temb), params[pre + ".Dense_0.weight"], params[pre + ".Dense_0.bias"])[:, :, None, None]
    h = _silu(_gn32(params, pre + ".GroupNorm_1", h))
    h = _conv(params, pre + ".Conv_1", h, 1)
    if (pre + ".NIN_0.W") in params:
        x = _nin(params, pre + ".NIN_0", x)
    return x + h


def attn_block(params, idx, x):
    """AttnBlock.forward (models/layers.py:577-590)."""
    pre = f"all_modules.{idx}"
    b, c, hh, ww = x.shape
    h = _gn32(params, pre + ".GroupNorm_0", x)
    q = _nin(params, pre + ".NIN_0", h).reshape(b, c, hh * ww)
    k = _nin(params, pre + ".NIN_1", h).reshape(b, c, hh * ww)
    v = _nin(params, pre + ".NIN_2", h).reshape(b, c, hh * ww)
    probs = torch.softmax(torch.einsum("bci,bcj->bij", q, k) * (int(c) ** (-0.5)), dim=-1)
    h = torch.einsum("bij,bcj->bci", probs, v).reshape(b, c, hh, ww)
    return x + _nin(params, pre + ".NIN_3", h)


def forward(params, o, x, labels):
    """DDPM.forward (models/ddpm.py:150-213). x [B, input_channels, H, W], labels [B]."""
    m = 0
    temb = timestep_embedding(labels, o.nf)
    temb = F.linear(temb, params["all_modules.0.weight"], params["all_modules.0.bias"])
    temb = F.linear(_silu(temb), params["all_modules.1.weight"], params["all_modules.1.bias"])
    m = 2
    h = x if o.centered else 2 * x - 1.0
    hs = [_conv(params, f"all_modules.{m}", h, 1)]
    m += 1
    n_res = len(o.ch_mult)
    for lvl in range(n_res):
        for _ in range(o.num_res_blocks):
            h = resblock(params, m, hs[-1], temb)
            m += 1
            if h.shape[-1] in o.attn_resolutions:
                h = attn_block(params, m, h)
                m += 1
            hs.append(h)
        if lvl != n_res - 1:
            if o.resamp_with_conv:      # Downsample: F.pad (0,1,0,1) + stride-2 VALID conv (models/layers.py:619-623)
                hs.append(_conv(params, f"all_modules.{m}.Conv_0", F.pad(hs[-1], (0, 1, 0, 1)), 0, stride=2))
            else:
                hs.append(F.avg_pool2d(hs[-1], 2))
            m += 1
    h = hs[-1]
    h = resblock(params, m, h, temb); m += 1
    h = attn_block(params, m, h); m += 1
    h = resblock(params, m, h, temb); m += 1
    for lvl in reversed(range(n_res)):
        for _ in range(o.num_res_blocks + 1):
            h = resblock(params, m, torch.cat([h, hs.pop()], dim=1), temb)
            m += 1
        if h.shape[-1] in o.attn_resolutions:
            h = attn_block(params, m, h)
            m += 1
        if lvl != 0:                    # Upsample: nearest x2 (+ conv) (models/layers.py:600-604)
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            if o.resamp_with_conv:
                h = _conv(params, f"all_modules.{m}.Conv_0", h, 1)
            m += 1
    assert not hs
    h = _silu(_gn32(params, f"all_modules.{m}", h)); m += 1
    h = _conv(params, f"all_modules.{m}", h, 1); m += 1
    return h


def forward_paired(params, o, x, y, labels):
    """DDPM_paired.forward (models/ddpm.py:292-298)."""
    out = forward(params, o, torch.cat([x, y], dim=1), labels)
    c = x.shape[1]
    return {"x": out[:, :c], "y": out[:, c:]}


def forward_paired_sr3(params, o, x, y, labels):
    """DDPM_paired_SR3.forward (models/ddpm.py:280-285)."""
    return forward(params, o, torch.cat([x, y], dim=1), labels)
