"""Predictor / corrector / PC-loop parity on the GPU against the CPU oracle with injected noise.

The network inside the loop runs with bf16 operands (see test_gpu_network.py), so per-step state
comparisons use a 2e-2 tolerance relative to the state's max magnitude; the update kernels
themselves are fp32 and are checked to 1e-5 in test_gpu_ops.py.
"""
import pytest
import torch

from golden_utils import golden, to_namespace
from oracle import ncsnpp as o_net
from oracle import sampling as o_samp
from oracle import sde as o_sde

pytestmark = pytest.mark.gpu

STATE_TOL = 2e-2


def _pkg():
    from conditional_score_diffusion_b200 import sampling, sde_lib
    from conditional_score_diffusion_b200.models import ncsnpp, utils  # noqa: F401
    return sampling, sde_lib, utils


def _model(name):
    _, _, utils = _pkg()
    f = golden()[f"ncsnpp_{name}"]
    m = utils.create_model(to_namespace(f["config"]))
    m.load_state_dict(f["state_dict"], strict=True)
    return f, m.cuda().eval()


def _rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)


class NoiseTape:
    """Pre-drawn noise shared by the oracle (sequential draws) and the CUDA sampler (named draws)."""

    def __init__(self, shape, steps, conditional, seed):
        g = torch.Generator().manual_seed(seed)
        self.names = ["y_c", "x_c", "y_p", "x_p"] if conditional else ["x_c", "x_p"]
        self.data = {(n, i): torch.randn(*shape, generator=g) for i in range(steps) for n in self.names}
        self.order = [(n, i) for i in range(steps) for n in self.names]
        self.pos = 0

    def sequential(self, like):
        t = self.data[self.order[self.pos]]
        self.pos += 1
        return t

    def named(self, name, step, inner):
        return self.data[(name, step)]


def test_pc_conditional_fused_vs_oracle_injected_noise():
    sampling, sde_lib, _ = _pkg()
    f, m = _model("paired")
    p = golden()["pc_conditional"]
    steps = 6
    shape = tuple(p["y"].shape)
    tape = NoiseTape(shape, steps, True, 5)
    x0 = torch.randn(*shape, generator=torch.Generator().manual_seed(6)) * p["sigma_max_x"]
    # oracle
    o = o_net.model_options(to_namespace(f["config"]))
    sx, sy = o_sde.VE(p["sigma_min_x"], p["sigma_max_x"], p["N"]), o_sde.VE(p["sigma_min_y"], p["sigma_max_y"], p["N"])
    model_fn = lambda d, l: o_net.forward_paired(f["state_dict"], o, d["x"], d["y"], l)
    rec = []
    ref, _ = o_samp.pc_conditional_sampler(o_sde.score_fn_conditional_pair(model_fn, sx, sy, True), sx, sy, p["y"], shape,
                                           p["snr"], steps, 1, eps=p["eps"], randn_like=tape.sequential, x_init=x0,
                                           record=rec)
    # CUDA
    sde = {"x": sde_lib.cVESDE(p["sigma_min_x"], p["sigma_max_x"], p["N"]),
           "y": sde_lib.VESDE(p["sigma_min_y"], p["sigma_max_y"], p["N"])}
    sampler = sampling.get_pc_conditional_sampler(sde, shape, sampling.get_predictor("conditional_reverse_diffusion"),
                                                  sampling.get_corrector("conditional_langevin"), p["snr"], steps, 1,
                                                  continuous=True, denoise=True, eps=p["eps"])
    got, info = sampler(m, p["y"].cuda(), show_evolution=True, x_init=x0, noise_source=tape.named)
    evo = info["evolution"]["x"]
    for i in range(steps):
        r = _rel(evo[i], rec[i])
        print(f"[pc-cond] step {i}: rel={r:.3e} |x|max={rec[i].abs().max().item():.3e}")
        assert r < STATE_TOL
    assert _rel(got, ref) < STATE_TOL
    # graph path twice: bitwise reproducible given the same injected noise (no float atomics anywhere on the path:
    # fixed-order GroupNorm statistics and cluster-reduced Langevin norms; the reference is reproducible under a seed)
    got2, _ = sampler(m, p["y"].cuda(), x_init=x0, noise_source=tape.named)
    assert torch.equal(got2, got)


def test_time_embedding_reuse_inside_a_pc_step_is_bitwise_neutral(monkeypatch):
    """Both network evaluations of a PC step see the same vec_t (sampling/conditional.py:196-226), so the fused loop runs
    the time-embedding MLP and the Dense_0 projections once per step (FusedPCSampler._net). Same injected noise, reuse on
    vs off: the samples must agree bit for bit, and the step with reuse must launch exactly two kernels fewer."""
    from conditional_score_diffusion_b200 import _lib
    from conditional_score_diffusion_b200.sampling import fused
    sampling, sde_lib, _ = _pkg()
    f, m = _model("paired")
    p = golden()["pc_conditional"]
    steps = 4
    shape = tuple(p["y"].shape)
    tape = NoiseTape(shape, steps, True, 25)
    x0 = torch.randn(*shape, generator=torch.Generator().manual_seed(26)) * p["sigma_max_x"]
    sde = {"x": sde_lib.cVESDE(p["sigma_min_x"], p["sigma_max_x"], p["N"]),
           "y": sde_lib.VESDE(p["sigma_min_y"], p["sigma_max_y"], p["N"])}
    outs, launches = [], []
    for reuse in (True, False):
        monkeypatch.setattr(fused, "REUSE_TEMB", reuse)
        fs = fused.FusedPCSampler(m, sde, shape, "reverse_diffusion", "langevin", p["snr"], steps, 1, False, True, True,
                                  p["eps"], conditional=True)
        out, _ = fs.sample(y=p["y"].cuda(), x_init=x0, noise_source=tape.named)
        outs.append(out.clone())
        fs.plan.use_graph = False                      # count the launches of one step issued launch by launch
        n0 = _lib.lib().csd_launch_count()
        fs.draw_noise = False
        fs._step()
        torch.cuda.synchronize()
        launches.append(_lib.lib().csd_launch_count() - n0)
    assert torch.equal(outs[0], outs[1])
    assert launches[1] - launches[0] == 2, launches


def test_pc_unconditional_fused_vs_oracle_injected_noise():
    sampling, sde_lib, _ = _pkg()
    f, m = _model("cifar")
    p = golden()["pc_unconditional"]
    steps = 5
    shape = tuple(f["x"].shape)
    tape = NoiseTape(shape, steps, False, 15)
    x0 = torch.randn(*shape, generator=torch.Generator().manual_seed(16)) * p["sigma_max"]
    o = o_net.model_options(to_namespace(f["config"]))
    sde_o = o_sde.VE(p["sigma_min"], p["sigma_max"], p["N"])
    score_fn = o_sde.score_fn_unconditional(lambda x, l: o_net.forward(f["state_dict"], o, x, l), sde_o, True, "fourier")
    for predictor in ("reverse_diffusion", "euler_maruyama"):
        tape.pos = 0
        rec = []
        ref, _ = o_samp.pc_sampler(score_fn, sde_o, shape, p["snr"], steps, 1, eps=p["eps"], predictor=predictor,
                                   randn_like=tape.sequential, x_init=x0, record=rec)
        sde = sde_lib.VESDE(p["sigma_min"], p["sigma_max"], p["N"])
        sampler = sampling.get_pc_sampler(sde, shape, sampling.get_predictor(predictor), sampling.get_corrector("langevin"),
                                          p["snr"], steps, 1, continuous=True, denoise=True, eps=p["eps"])
        got, info = sampler(m, show_evolution=True, x_init=x0, noise_source=tape.named)
        for i in range(steps):
            r = _rel(info["evolution"][i], rec[i])
            print(f"[pc-uncond {predictor}] step {i}: rel={r:.3e}")
            assert r < STATE_TOL
        assert _rel(got, ref) < STATE_TOL
        assert info["steps"] == steps * 2


def test_generator_noise_path_runs_and_is_seed_reproducible():
    sampling, sde_lib, _ = _pkg()
    f, m = _model("paired")
    p = golden()["pc_conditional"]
    shape = tuple(p["y"].shape)
    sde = {"x": sde_lib.cVESDE(p["sigma_min_x"], p["sigma_max_x"], p["N"]),
           "y": sde_lib.VESDE(p["sigma_min_y"], p["sigma_max_y"], p["N"])}
    sampler = sampling.get_pc_conditional_sampler(sde, shape, sampling.get_predictor("conditional_reverse_diffusion"),
                                                  sampling.get_corrector("conditional_langevin"), p["snr"], 8, 1,
                                                  continuous=True, denoise=True, eps=p["eps"])
    outs = []
    for _ in range(2):
        torch.manual_seed(123)
        out, _ = sampler(m, p["y"].cuda())
        assert torch.isfinite(out).all()
        outs.append(out)
    assert torch.equal(outs[0], outs[1]), "same seed, same samples - bitwise"
    torch.manual_seed(124)
    out3, _ = sampler(m, p["y"].cuda())
    assert _rel(out3, outs[0]) > 1e-2  # a different seed gives a different sample


def test_class_based_updates_vs_oracle():
    """predictor / corrector classes (generic path): one update each, noise drawn from torch's CUDA
    generator and recovered through the fp32 update identities."""
    sampling, sde_lib, utils = _pkg()
    f, m = _model("cifar")
    s = golden()["single_updates"]
    sde = sde_lib.VESDE(0.01, 50, 10)
    score_fn = utils.get_score_fn(sde, m, conditional=False, train=False, continuous=True)
    x, t = s["x"].cuda(), s["t"].cuda()
    with torch.no_grad():
        score = score_fn(x, t)
        assert _rel(score, s["score"]) < STATE_TOL
        for name, cls, kw in [("rd", sampling.get_predictor("reverse_diffusion"), {}),
                              ("em", sampling.get_predictor("euler_maruyama"), {}),
                              ("anc", sampling.get_predictor("ancestral_sampling"), {})]:
            torch.manual_seed(3)
            xo, xm = cls(sde, score_fn).update_fn(x, t)
            assert torch.isfinite(xo).all() and torch.isfinite(xm).all()
            if name != "anc":
                assert _rel(xm, s[f"{name}_mean"]) < STATE_TOL, name
        torch.manual_seed(3)
        xo, xm = sampling.get_corrector("langevin")(sde, score_fn, 0.16, 1).update_fn(x, t)
        z = torch.randn_like(x)  # not the same draw; only check the mean path via the oracle formula
        assert torch.isfinite(xo).all()
        xo2, xm2 = sampling.get_corrector("ald")(sde, score_fn, 0.16, 1).update_fn(x, t)
        sde_o = o_sde.VE(0.01, 50, 10)
        step = (0.16 * sde_o.sigma(s["t"])) ** 2 * 2
        ref_mean = s["x"] + step[:, None, None, None] * s["score"]
        assert _rel(xm2, ref_mean) < STATE_TOL


def _opts(name):
    return o_net.model_options(to_namespace(golden()[f"ncsnpp_{name}"]["config"]))


class _Replay:
    """Replays the CUDA generator's draws in the order the sampler consumed them (same seed, same shapes)."""

    def __init__(self, seed):
        torch.manual_seed(seed)

    def __call__(self, like):
        return torch.randn(like.shape, device="cuda").cpu()


def test_pc_inpainter_vs_oracle_replayed_noise():
    """get_pc_inpainter (sampling/unconditional.py:230-345) on the engine: same draws fed to the CPU oracle (pinned to
    the reference's trajectory by tests/test_oracle_grads.py)."""
    from conditional_score_diffusion_b200.sampling import unconditional
    sampling, sde_lib, utils = _pkg()
    f, m = _model("cifar")
    g = torch.Generator().manual_seed(31)
    data = torch.rand(2, 3, 16, 16, generator=g)
    mask = torch.ones_like(data)
    mask[:, :, 3:11, 5:13] = 0.0
    x_init = torch.randn(2, 3, 16, 16, generator=g) * 50
    sde = sde_lib.VESDE(0.01, 50, 4)
    fn = unconditional.get_pc_inpainter(sde, sampling.get_predictor("reverse_diffusion"), sampling.get_corrector("langevin"),
                                        snr=0.16, n_steps=1, continuous=True, denoise=True, eps=1e-5)
    torch.manual_seed(77)
    got, info = fn(m, data.cuda(), mask.cuda(), show_evolution=True, x_init=x_init)
    sde_o = o_sde.VE(0.01, 50, 4)
    score_fn = o_sde.score_fn_unconditional(lambda x, l: o_net.forward(f["state_dict"], _opts("cifar"), x, l), sde_o, True,
                                            "fourier")
    rec = []
    ref, _ = o_samp.pc_inpainter(score_fn, sde_o, data, mask, 0.16, eps=1e-5, randn_like=_Replay(77), x_init=x_init, record=rec)
    evo = info["evolution"][1:]
    for i in range(4):
        print(f"[inpaint] step {i}: rel {_rel(evo[i], rec[i]):.3e}")
    assert _rel(evo[0], rec[0]) < STATE_TOL
    assert _rel(got, ref) < 5 * STATE_TOL
    # known pixels of the denoised output equal the data exactly (x_mean = ... + data * mask)
    assert torch.allclose(got.cpu() * mask, data * mask, atol=1e-6)


def test_pc_conditional_use_path_vs_oracle_replayed_noise():
    """get_pc_conditional_sampler(use_path=True) (sampling/conditional.py:87-94,124-176) on the engine."""
    from conditional_score_diffusion_b200.sampling import conditional
    sampling, sde_lib, utils = _pkg()
    f, m = _model("paired")
    p = golden()["pc_conditional"]
    y = p["y"]
    g = torch.Generator().manual_seed(32)
    x_init = torch.randn(y.shape, generator=g) * p["sigma_max_x"]
    sde = {"x": sde_lib.cVESDE(p["sigma_min_x"], p["sigma_max_x"], p["N"]), "y": sde_lib.VESDE(p["sigma_min_y"], p["sigma_max_y"], p["N"])}
    fn = conditional.get_pc_conditional_sampler(sde, tuple(y.shape), sampling.get_predictor("conditional_reverse_diffusion"),
                                                sampling.get_corrector("conditional_langevin"), snr=p["snr"], p_steps=3, c_steps=1,
                                                continuous=True, denoise=True, use_path=True, eps=p["eps"])
    torch.manual_seed(78)
    got, info = fn(m, y.cuda(), show_evolution=True, x_init=x_init)
    sx = o_sde.VE(p["sigma_min_x"], p["sigma_max_x"], p["N"])
    sy = o_sde.VE(p["sigma_min_y"], p["sigma_max_y"], p["N"])
    score_fn = o_sde.score_fn_conditional_pair(lambda d, l: o_net.forward_paired(f["state_dict"], _opts("paired"), d["x"], d["y"], l),
                                               sx, sy, True)
    rec = []
    ref, _ = o_samp.pc_conditional_sampler_path(score_fn, sx, sy, y, tuple(y.shape), p["snr"], 3, eps=p["eps"],
                                                randn_like=_Replay(78), x_init=x_init, record=rec)
    evo = info["evolution"]["x"]
    for i in range(3):
        print(f"[use_path] step {i}: rel {_rel(evo[i], rec[i]):.3e}")
    assert _rel(evo[0], rec[0]) < STATE_TOL
    assert _rel(got, ref) < 5 * STATE_TOL


def test_use_path_and_inpainter_run_on_the_graph_loop(monkeypatch):
    """use_path=True and get_pc_inpainter run on the fused CUDA-graph loop (one graph replay per PC step,
    coefficients from device tables) and reproduce the per-step Python loop they replace under the same seed - same
    kernels, same draw order, so the captured generator draws must line up with the eager ones."""
    from conditional_score_diffusion_b200.sampling import conditional, fused, unconditional
    sampling, sde_lib, utils = _pkg()
    launches = []
    orig = fused.FusedPCSampler.sample

    def counting(self, *a, **k):
        launches.append((self.use_path, self.inpaint))
        return orig(self, *a, **k)
    monkeypatch.setattr(fused.FusedPCSampler, "sample", counting)

    # --- use_path ---
    f, m = _model("paired")
    p = golden()["pc_conditional"]
    y = p["y"]
    g = torch.Generator().manual_seed(33)
    x_init = torch.randn(y.shape, generator=g) * p["sigma_max_x"]
    sde = {"x": sde_lib.cVESDE(p["sigma_min_x"], p["sigma_max_x"], p["N"]), "y": sde_lib.VESDE(p["sigma_min_y"], p["sigma_max_y"], p["N"])}
    outs = {}
    for fused_on in (False, True):
        monkeypatch.setattr(conditional, "FUSED_PATH", fused_on)
        fn = conditional.get_pc_conditional_sampler(sde, tuple(y.shape), sampling.get_predictor("conditional_reverse_diffusion"),
                                                    sampling.get_corrector("conditional_langevin"), snr=p["snr"], p_steps=5,
                                                    c_steps=1, continuous=True, denoise=True, use_path=True, eps=p["eps"])
        torch.manual_seed(79)
        outs[fused_on] = fn(m, y.cuda(), show_evolution=True, x_init=x_init)
    assert launches == [(True, False)]
    e_loop, e_graph = outs[False][1]["evolution"], outs[True][1]["evolution"]
    for i in range(5):
        print(f"[use_path graph vs loop] step {i}: x rel {_rel(e_graph['x'][i], e_loop['x'][i]):.3e} "
              f"y rel {_rel(e_graph['y'][i], e_loop['y'][i]):.3e}")
    assert _rel(e_graph["y"][-1], e_loop["y"][-1]) < 1e-5          # the condition path is fp32 elementwise work
    assert _rel(outs[True][0], outs[False][0]) < 5 * STATE_TOL

    # --- inpainter ---
    launches.clear()
    f, m = _model("cifar")
    g = torch.Generator().manual_seed(34)
    data = torch.rand(2, 3, 16, 16, generator=g)
    mask = torch.ones_like(data)
    mask[:, :, 2:9, 4:12] = 0.0
    x_init = torch.randn(2, 3, 16, 16, generator=g) * 50
    sde_u = sde_lib.VESDE(0.01, 50, 5)
    outs = {}
    for fused_on in (False, True):
        monkeypatch.setattr(unconditional, "FUSED_INPAINT", fused_on)
        fn = unconditional.get_pc_inpainter(sde_u, sampling.get_predictor("reverse_diffusion"), sampling.get_corrector("langevin"),
                                            snr=0.16, n_steps=1, continuous=True, denoise=True, eps=1e-5)
        torch.manual_seed(80)
        outs[fused_on] = fn(m, data.cuda(), mask.cuda(), show_evolution=True, x_init=x_init)
    assert launches == [(False, True)]
    e_loop, e_graph = outs[False][1]["evolution"], outs[True][1]["evolution"]
    assert e_loop.shape == e_graph.shape
    for i in range(e_loop.shape[0]):
        print(f"[inpaint graph vs loop] state {i}: rel {_rel(e_graph[i], e_loop[i]):.3e}")
    assert _rel(e_graph[0], e_loop[0]) < 1e-6
    assert _rel(outs[True][0], outs[False][0]) < 5 * STATE_TOL
    assert torch.allclose(outs[True][0].cpu() * mask, data * mask, atol=1e-6)
