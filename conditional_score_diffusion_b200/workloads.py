"""The BASELINE.json workloads as plain config objects (what the reference's config files hand to
`mutils.create_model`), shared by bench.py and the real-shape parity tests.

Each builder cites the reference config it restates. Only the fields the hot path reads are present
(`config.model.*`, `config.data.*`, `config.training.continuous`, `config.sampling.*`, `config.optim.*`).
Random-init weights use init_scale=1: with the shipped init_scale=0 the last convolutions are 1e-10-scaled and the
Langevin step size overflows.
"""
import math
from types import SimpleNamespace as NS


def config2_ncsnpp_paired_160(image=160, nf=96, ch_mult=(1, 1, 2, 2, 3, 3), attn=(20, 10, 5), num_res_blocks=2,
                              name="ncsnpp_paired"):
    """configs/ve/inverse_problems/super_resolution/celebA_ours_NDV_160.py:83-144 with model.name='ncsnpp_paired'
    (the north-star form); name='ddpm_paired' gives the file as shipped."""
    c = NS(
        training=NS(continuous=True),
        data=NS(image_size=image, effective_image_size=image, num_channels=6, centered=False),
        model=NS(name=name, nf=nf, ch_mult=ch_mult, num_res_blocks=num_res_blocks,
                 attn_resolutions=attn, dropout=0.1, resamp_with_conv=True, conditional=True, fir=True,
                 fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type="biggan", progressive="output_skip",
                 progressive_input="input_skip", progressive_combine="sum", embedding_type="positional",
                 init_scale=1.0, fourier_scale=16, nonlinearity="swish", num_scales=1000,
                 input_channels=6, output_channels=6,
                 sigma_max_x=math.sqrt(3 * image * image), sigma_max_y=0.5, sigma_min_x=5e-3, sigma_min_y=5e-3),
        sampling=NS(method="pc", predictor="conditional_reverse_diffusion", corrector="conditional_langevin",
                    n_steps_each=1, noise_removal=True, probability_flow=False, snr=0.15),
    )
    return c


def config3_ddpm_paired_128():
    """configs/ve/inverse_problems/inpainting/celebA_ours_DV.py:81-144 as shipped: ddpm_paired, nf 96,
    ch_mult (1,1,2,2,3,3), attention at 16/8/4, 128x128, sigma_max_x = sigma_max_y = sqrt(3*128^2)."""
    c = config2_ncsnpp_paired_160(image=128, attn=(16, 8, 4), name="ddpm_paired")
    c.model.sigma_max_y = math.sqrt(3 * 128 * 128)
    return c


def config4_ddpm_sr3_64(image=64):
    """configs/ve/inverse_problems/image_to_image_translation/edges2shoes_SR3.py:85-138 as shipped: ddpm_paired_SR3,
    nf 128, ch_mult (1,1,2,2), attention at 16/8, 64x64, 6 -> 3 channels; image=128 is BASELINE.json's
    wording of the same config."""
    c = NS()
    c.training = NS(continuous=True)
    c.data = NS(image_size=image, effective_image_size=image, num_channels=6, centered=False)
    c.model = NS(name="ddpm_paired_SR3", nf=128, ch_mult=(1, 1, 2, 2), num_res_blocks=2, attn_resolutions=(16, 8),
                 dropout=0.1, resamp_with_conv=True, conditional=True, nonlinearity="swish", input_channels=6,
                 output_channels=3, num_scales=1000, sigma_min_x=5e-3, sigma_max_x=math.sqrt(3 * image * image))
    c.optim = NS(weight_decay=0, optimizer="Adam", lr=2e-4, beta1=0.9, eps=1e-8, warmup=2500, grad_clip=1.0)
    return c


def config5_ncsnpp_256(image=256):
    """configs/ve/church_ncsnpp_continuous.py:40-62 on configs/default_lsun_configs.py (+ data.effective_image_size,
    which NCSNpp.__init__ needs): unconditional NCSN++, nf 128, ch_mult (1,1,2,2,2,2,2), attention at
    16x16, Fourier embedding, VESDE(0.01, 380, 2000), snr 0.075."""
    c = NS(
        training=NS(continuous=True),
        data=NS(image_size=image, effective_image_size=image, num_channels=3, centered=False),
        model=NS(name="ncsnpp", nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,),
                 dropout=0.0, resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1],
                 skip_rescale=True, resblock_type="biggan", progressive="output_skip", progressive_input="input_skip",
                 progressive_combine="sum", embedding_type="fourier", init_scale=1.0, fourier_scale=16,
                 nonlinearity="swish", num_scales=2000, sigma_min=0.01, sigma_max=380.0),
        sampling=NS(method="pc", predictor="reverse_diffusion", corrector="langevin", n_steps_each=1,
                    noise_removal=True, probability_flow=False, snr=0.075),
    )
    return c


def randomize_small_params(model, seed=1):
    """Give biases / zero-initialised tensors a non-trivial value (the reference initialises biases to zero, so random
    init alone would leave every bias path untested). Deterministic on CPU; call before `.cuda()`."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for pn, p in model.named_parameters():
            if pn.endswith("bias") or pn.endswith(".b"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif p.abs().max() < 1e-6:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return model
