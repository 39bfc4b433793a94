"""Fused predictor-corrector loop: one CUDA graph per PC step.

The reference's loop (sampling/unconditional.py:204-216, sampling/conditional.py:196-226) rebuilds
score-function closures, predictor/corrector objects and an RSDE class on every step, uploads
`vec_t` from the host, and issues ~1 700 eager kernels per step. Here a step is:

    labels / 1/sigma  <- device tables[step]          (1 tiny kernel each)
    [y_t = y + sigma_y z]                              (conditional only, 1 kernel)
    score network                                      (engine launch list)
    Langevin norms + update                            (3 kernels)
    [y_t again], score network, predictor update       (1 kernel)
    step += 1

captured once with torch.cuda.graph and replayed p_steps times: no host arithmetic, no H2D copies
and no Python between launches. The state x lives in the network's own input buffer, so the update
kernels write the next network input in place.

Two more step shapes run on the same graph machinery:
  * use_path (sampling/conditional.py:84-176): the condition walks ONE path y_T -> y_0 through the backward kernel
    p(y_t | y_0, y_{t+tau}); the state y_t lives in the network's second input buffer and is advanced in place from
    per-step coefficient tables; predictor first, then the corrector on the same y_t.
  * inpainting (sampling/unconditional.py:230-345): after the corrector and after the predictor the known pixels are
    replaced by the data at the current noise level (one merge kernel each, coefficients from device tables).

Noise: torch's CUDA generator (`normal_` on static buffers, captured in the graph) in the same draw
order as the reference (y, x, y, x per conditional step); `noise_source` replaces it with injected
tensors for parity tests and reproducible replays.
"""
import os

import torch

from .. import kernels as K
from . import tables


REUSE_TEMB = os.environ.get("CSD_NO_TEMB_REUSE", "0") != "1"   # A/B switch


class FusedPCSampler:
    def __init__(self, model, sde, shape, predictor, corrector, snr, p_steps, c_steps, probability_flow,
                 continuous, denoise, eps, conditional, use_path=False, inpaint=False):
        self.model = model
        self.sde = sde
        self.shape = tuple(shape)
        self.predictor = predictor          # 'reverse_diffusion' | 'euler_maruyama' | 'none'
        self.corrector = corrector          # 'langevin' | 'none'
        self.snr, self.p_steps, self.c_steps = snr, p_steps, c_steps
        self.probability_flow, self.continuous, self.denoise, self.eps = probability_flow, continuous, denoise, eps
        self.conditional = conditional      # {'x','y'} pair with a perturbed condition
        self.use_path = use_path            # conditional pair: y follows one backward-kernel path
        self.inpaint = inpaint              # unconditional: known pixels re-imposed after every update
        if use_path and not conditional:
            raise ValueError("use_path needs the {'x', 'y'} SDE pair")
        if inpaint and conditional:
            raise ValueError("the inpainter runs on an unconditional score model")
        self.c_sde = sde["x"] if isinstance(sde, dict) else sde
        self.graphs = {}
        self.ready = False

    # ---- setup -----------------------------------------------------------------------------------
    def _setup(self, device):
        m = self.model
        m._engine.ensure_packed(device, force_refresh=True)
        b, c, h, w = self.shape
        c1 = c if self.conditional else 0
        if self.conditional and m.in_channels != 2 * c:
            raise ValueError("conditional pair sampler expects x and y with the same channel count")
        self.plan = m._engine.plan(b, h, w, c, c1)
        self.x = self.plan.in0                      # state lives in the network input buffer
        self.score = self.plan.outputs()[0]
        self.x_mean = torch.empty_like(self.x)
        self.noise_x = [torch.empty_like(self.x) for _ in range(2)]      # corrector, predictor
        self.norms = torch.empty(2 * b, device=device, dtype=torch.float32)
        self.step_idx = torch.zeros(1, device=device, dtype=torch.int32)
        ts = tables.time_grid(self.c_sde, self.eps, self.p_steps)
        self.timesteps = ts
        labels, inv = tables.model_time_tables(self.sde, ts, self.continuous, self.conditional or
                                               isinstance(self.sde, dict), m.embedding_type)
        dev = lambda t: None if t is None else t.to(device=device, dtype=torch.float32).contiguous()
        self.t_labels = dev(labels)
        if isinstance(inv, dict):
            self.t_inv_x, self.t_inv_y = dev(inv["x"]), dev(inv["y"])
        else:
            self.t_inv_x, self.t_inv_y = dev(inv), None
        if self.predictor != "none":
            lin, g = tables.predictor_tables(self.c_sde, ts, self.predictor)
            self.t_lin, self.t_g = dev(lin), dev(g)
        self.t_alpha = dev(tables.langevin_alpha_table(self.c_sde, ts))
        if self.conditional:
            self.y = torch.empty_like(self.x)
            self.noise_y = [torch.empty_like(self.x) for _ in range(2)]
            self.t_sigma_y = dev(self.sde["y"].marginal_prob(ts, ts)[1])
        if self.use_path:
            # VESDE.compute_backward_kernel (sde_lib.py:323-339) on the whole time grid: y_t = a y_0 + b y_{t+tau} + s z
            sy = self.sde["y"]
            tau = ts[0] - ts[1]
            s_t = sy.marginal_prob(ts, ts)[1] ** 2
            s_tau = sy.marginal_prob(ts, ts + tau)[1] ** 2
            self.t_pa, self.t_pb = dev((s_tau - s_t) / s_tau), dev(s_t / s_tau)
            self.t_ps = dev(torch.sqrt(s_t * (s_tau - s_t) / s_tau))
            one = torch.ones(1)
            self.std_T = dev(sy.marginal_prob(one, one * (ts[0] + tau))[1].expand(b))
            self.vec_a, self.vec_b, self.vec_s = (torch.empty(b, device=device, dtype=torch.float32) for _ in range(3))
            self.y_tmp = torch.empty_like(self.x)
        if self.inpaint:
            mc, ms = self.c_sde.marginal_prob(torch.ones_like(ts), ts)
            self.t_mc, self.t_ms = dev(mc), dev(ms)
            self.vec_mc, self.vec_ms = (torch.empty(b, device=device, dtype=torch.float32) for _ in range(2))
            self.data, self.mask = torch.empty_like(self.x), torch.empty_like(self.x)
            self.noise_m = [torch.zeros_like(self.x) for _ in range(2)]      # merge after corrector, after predictor
        self.ready = True

    # ---- one PC step (recorded into a graph) ---------------------------------------------------------
    def _score(self, which):
        if self.conditional:
            if self.draw_noise:
                self.noise_y[which].normal_()
            K.ve_perturb(self.y, self.noise_y[which], self.plan.in1, self.t_sigma_y, self.step_idx, 0)
        self._net()

    def _net(self):
        """One score-network evaluation. Every evaluation of a PC step sees the same vec_t (corrector first, then the
        predictor, sampling/unconditional.py:216-220, sampling/conditional.py:196-226), so only the first one runs
        the time-embedding MLP and the Dense_0 projections; the others reuse what it wrote."""
        self.plan.launch(reuse_time_embedding=REUSE_TEMB and not self._temb_stale)
        self._temb_stale = False

    def _corrector(self, fresh_condition):
        p = self.plan
        for k in range(self.c_steps):
            if k == 0 and fresh_condition:
                self._score(0)
            else:
                self._net()  # same condition for every inner Langevin iteration (and, on a path, as the predictor)
            if self.draw_noise:
                self.noise_x[0].normal_()
            K.langevin_norms(self.score, self.noise_x[0], self.norms)
            K.langevin_update(self.x, self.score, self.noise_x[0], self.norms, self.x, self.x_mean, self.snr,
                              self.t_alpha, self.step_idx, 0)

    def _predictor(self, fresh_condition):
        if fresh_condition:
            self._score(1)
        else:
            self._net()
        if self.draw_noise:
            self.noise_x[1].normal_()
        if self.predictor == "reverse_diffusion":
            K.reverse_diffusion_update(self.x, self.score, self.noise_x[1], self.x, self.x_mean, self.t_lin,
                                       self.t_g, self.probability_flow, self.step_idx, 0)
        else:
            K.euler_maruyama_update(self.x, self.score, self.noise_x[1], self.x, self.x_mean, self.t_lin, self.t_g,
                                    -1.0 / self.c_sde.N, self.probability_flow, self.step_idx, 0)

    def _merge(self, which):
        """Inpainting: x <- x (1 - mask) + (mean_coef data + std z) mask, x_mean likewise without the noise
        (sampling/unconditional.py:259-275), in place on the network's input buffer."""
        if self.draw_noise:
            self.noise_m[which].normal_()
        K.inpaint_merge(self.x, self.data, self.noise_m[which], self.mask, self.x, self.x_mean, self.vec_mc, self.vec_ms)

    def _step(self):
        p = self.plan
        self._temb_stale = True
        K.broadcast_table(p.labels, self.t_labels, self.step_idx, 0)
        K.broadcast_table(p.row_scale, self.t_inv_x, self.step_idx, 0)
        if self.t_inv_y is not None:
            K.broadcast_table(p.row_scale1, self.t_inv_y, self.step_idx, 0)
        if self.use_path:
            # y_t ~ p(y_t | y_0, y_{t+tau}) in place on the network's condition input, then predictor and corrector on it
            K.broadcast_table(self.vec_a, self.t_pa, self.step_idx, 0)
            K.broadcast_table(self.vec_b, self.t_pb, self.step_idx, 0)
            K.broadcast_table(self.vec_s, self.t_ps, self.step_idx, 0)
            if self.draw_noise:
                self.noise_y[1].normal_()
            K.sde_perturb(self.y, p.in1, self.y_tmp, self.vec_a, self.vec_b)
            K.sde_perturb(self.y_tmp, self.noise_y[1], p.in1, None, self.vec_s)
            if self.predictor != "none":
                self._predictor(False)
            if self.corrector == "langevin":
                self._corrector(False)
            K.step_advance(self.step_idx)
            return
        if self.inpaint:
            K.broadcast_table(self.vec_mc, self.t_mc, self.step_idx, 0)
            K.broadcast_table(self.vec_ms, self.t_ms, self.step_idx, 0)
        if self.corrector == "langevin":
            self._corrector(True)
        if self.inpaint:
            self._merge(0)
        if self.predictor != "none":
            self._predictor(True)
        if self.inpaint:
            self._merge(1)
        K.step_advance(self.step_idx)

    def _graph(self, draw_noise):
        g = self.graphs.get(draw_noise)
        if g is None:
            self.draw_noise = draw_noise
            rng_state = torch.cuda.get_rng_state(self.x.device)   # warm-up must not consume the caller's stream
            for _ in range(2):               # warm-up outside capture (lazy init, the plan's own warm-up)
                self.x.normal_()
                for n in self.noise_x + (self.noise_y if self.conditional else []) + (self.noise_m if self.inpaint else []):
                    n.normal_()
                self.step_idx.zero_()
                self._step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()
            torch.cuda.set_rng_state(rng_state, self.x.device)
            self.graphs[draw_noise] = g
        self.draw_noise = draw_noise
        return g

    # ---- public ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, y=None, x_init=None, noise_source=None, show_evolution=False, use_graph=True, data=None,
               mask=None):
        """Run the full loop. noise_source(name, step, inner) -> tensor replaces the generator (names 'y_c', 'x_c',
        'y_p', 'x_p'; on a path 'y_T' (step -1) and 'y_p'; inpainting adds 'z_c', 'z_p' for the merge draws).
        data, mask: the known image and its mask (inpainting only)."""
        device = self.model.device
        if self.ready:
            # the network may have new weights (load_state_dict, optimizer step, EMA swap through `.data`) or new
            # parameter storage (.to(), a fused optimizer re-homing them) since the graphs were captured: re-pack from
            # the live parameters on every call; if the engine rebuilt its plans, re-plan and re-capture
            b, c, h, w = self.shape
            eng = self.model._engine
            eng.ensure_packed(device, force_refresh=True)
            if eng.plan(b, h, w, c, c if self.conditional else 0) is not self.plan:
                self.ready = False
                self.graphs = {}
        if not self.ready:
            self._setup(device)
        if self.conditional:
            if y is None:
                raise ValueError("conditional sampler needs the condition y")
            self.y.copy_(y.to(device=device, dtype=torch.float32))
        if x_init is None:
            x_init = self.c_sde.prior_sampling(self.shape)      # CPU randn like the reference (sde_lib.py:341-347)
        injected = noise_source is not None
        if injected and self.c_steps != 1:
            raise NotImplementedError("noise injection supports c_steps == 1")
        g = self._graph(draw_noise=not injected) if use_graph else None
        self.draw_noise = not injected
        self.x.copy_(x_init.to(device=device, dtype=torch.float32))
        self.step_idx.zero_()
        evolution = {"x": [], "y": []}
        if self.use_path:
            # y_{T+tau} = y + sigma_y(T + tau) z starts the path (sampling/conditional.py:141-143)
            z_T = noise_source("y_T", -1, 0) if injected else torch.randn_like(self.y)
            K.sde_perturb(self.y, z_T.to(device=device, dtype=torch.float32).contiguous(), self.plan.in1, None, self.std_T)
        if self.inpaint:
            if data is None or mask is None:
                raise ValueError("the inpainter needs data and mask")
            self.data.copy_(data.to(device=device, dtype=torch.float32))
            self.mask.copy_(mask.to(device=device, dtype=torch.float32).expand_as(self.data))
            # x = data * mask + prior * (1 - mask) (sampling/unconditional.py:300): the merge kernel with coefficient 1, std 0
            self.vec_ms.zero_()
            K.inpaint_merge(self.x, self.data, self.noise_m[0], self.mask, self.x, self.x_mean, None, self.vec_ms)
            if show_evolution:
                evolution["x"].append(self.x.cpu())
        for i in range(self.p_steps):
            if injected:
                if self.use_path:
                    self.noise_y[1].copy_(noise_source("y_p", i, 0))
                elif self.conditional:
                    self.noise_y[0].copy_(noise_source("y_c", i, 0))
                    self.noise_y[1].copy_(noise_source("y_p", i, 0))
                if self.inpaint:
                    self.noise_m[0].copy_(noise_source("z_c", i, 0))
                    self.noise_m[1].copy_(noise_source("z_p", i, 0))
                if self.corrector == "langevin":
                    self.noise_x[0].copy_(noise_source("x_c", i, 0))
                if self.predictor != "none":
                    self.noise_x[1].copy_(noise_source("x_p", i, 0))
            if g is not None:
                g.replay()
            else:
                self._step()
            if show_evolution:
                evolution["x"].append(self.x.cpu())
                if self.conditional:
                    evolution["y"].append(self.plan.in1.cpu())
        out = (self.x_mean if self.denoise else self.x).clone()
        return out, evolution
