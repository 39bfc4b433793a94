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

Noise: torch's CUDA generator (`normal_` on static buffers, captured in the graph) in the same draw
order as the reference (y, x, y, x per conditional step); `noise_source` replaces it with injected
tensors for parity tests and reproducible replays.
"""
import torch

from .. import kernels as K
from . import tables


class FusedPCSampler:
    def __init__(self, model, sde, shape, predictor, corrector, snr, p_steps, c_steps, probability_flow,
                 continuous, denoise, eps, conditional):
        self.model = model
        self.sde = sde
        self.shape = tuple(shape)
        self.predictor = predictor          # 'reverse_diffusion' | 'euler_maruyama' | 'none'
        self.corrector = corrector          # 'langevin' | 'none'
        self.snr, self.p_steps, self.c_steps = snr, p_steps, c_steps
        self.probability_flow, self.continuous, self.denoise, self.eps = probability_flow, continuous, denoise, eps
        self.conditional = conditional      # {'x','y'} pair with a perturbed condition
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
        self.ready = True

    # ---- one PC step (recorded into a graph) ---------------------------------------------------------
    def _score(self, which):
        if self.conditional:
            if self.draw_noise:
                self.noise_y[which].normal_()
            K.ve_perturb(self.y, self.noise_y[which], self.plan.in1, self.t_sigma_y, self.step_idx, 0)
        self.plan.launch()

    def _step(self):
        p = self.plan
        K.broadcast_table(p.labels, self.t_labels, self.step_idx, 0)
        K.broadcast_table(p.row_scale, self.t_inv_x, self.step_idx, 0)
        if self.t_inv_y is not None:
            K.broadcast_table(p.row_scale1, self.t_inv_y, self.step_idx, 0)
        if self.corrector == "langevin":
            for k in range(self.c_steps):
                if k == 0:
                    self._score(0)
                else:
                    p.launch()   # same perturbed condition for every inner Langevin iteration
                if self.draw_noise:
                    self.noise_x[0].normal_()
                K.langevin_norms(self.score, self.noise_x[0], self.norms)
                K.langevin_update(self.x, self.score, self.noise_x[0], self.norms, self.x, self.x_mean, self.snr,
                                  self.t_alpha, self.step_idx, 0)
        if self.predictor != "none":
            self._score(1)
            if self.draw_noise:
                self.noise_x[1].normal_()
            if self.predictor == "reverse_diffusion":
                K.reverse_diffusion_update(self.x, self.score, self.noise_x[1], self.x, self.x_mean, self.t_lin,
                                           self.t_g, self.probability_flow, self.step_idx, 0)
            else:
                K.euler_maruyama_update(self.x, self.score, self.noise_x[1], self.x, self.x_mean, self.t_lin, self.t_g,
                                        -1.0 / self.c_sde.N, self.probability_flow, self.step_idx, 0)
        K.step_advance(self.step_idx)

    def _graph(self, draw_noise):
        g = self.graphs.get(draw_noise)
        if g is None:
            self.draw_noise = draw_noise
            rng_state = torch.cuda.get_rng_state(self.x.device)   # warm-up must not consume the caller's stream
            for _ in range(2):               # warm-up outside capture (lazy init, the plan's own warm-up)
                self.x.normal_()
                for n in self.noise_x + (self.noise_y if self.conditional else []):
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
    def sample(self, y=None, x_init=None, noise_source=None, show_evolution=False, use_graph=True):
        """Run the full loop. noise_source(name, step, inner) -> tensor replaces the generator
        (names 'y_c', 'x_c', 'y_p', 'x_p')."""
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
        for i in range(self.p_steps):
            if injected:
                if self.conditional:
                    self.noise_y[0].copy_(noise_source("y_c", i, 0))
                    self.noise_y[1].copy_(noise_source("y_p", i, 0))
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
