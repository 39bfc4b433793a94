"""Device-resident Dormand-Prince 5(4) integrator.

The reference solves the probability-flow ODE with `scipy.integrate.solve_ivp(method='RK45')` (likelihood.py:99,
sampling/unconditional.py:147): every right-hand side converts the state to a float64 numpy vector on the host and
back. `solve_rk45` restates scipy's RK45 - the same Butcher tableau, initial-step heuristic, RMS error norm, step-size
controller (safety 0.9, factors clamped to [0.2, 10], no growth right after a rejection) and FSAL reuse - with the state
y and the seven stage derivatives resident in HBM (fp32): stage combinations and the error norm are libcsd_b200 kernels
(csd_rk_combine_f32, csd_rk_error_sumsq_f32) and only the 4-byte error norm crosses to the host per step, where the
accept/reject decision is taken exactly as scipy takes it. Algorithm reference: scipy/integrate/_ivp/rk.py (RK45,
rk_step, RungeKutta._step_impl) and _ivp/common.py (select_initial_step); scipy is a dependency of the reference
(`from scipy import integrate`), unpinned there, 1.18 in this image.
"""
import math

import numpy as np
import torch

from . import kernels as K

# Dormand-Prince coefficients (scipy.integrate._ivp.rk.RK45)
_C = [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0]
_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
]
_B = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]
_E = [-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40]
_SAFETY, _MIN_FACTOR, _MAX_FACTOR, _ORDER = 0.9, 0.2, 10.0, 4


def solve_rk45(fun, t0, t_bound, y0, rtol=1e-5, atol=1e-5, max_steps=100000):
    """Integrate dy/dt = fun(t, y) from t0 to t_bound. y0: flat fp32 CUDA tensor [n]; fun(t: float, y) returns a flat fp32
    CUDA tensor [n] (it may return a view of its own buffer: the value is copied into the stage stack). Returns
    (y(t_bound), number of function evaluations)."""
    if not (y0.is_cuda and y0.dtype == torch.float32 and y0.dim() == 1):
        raise RuntimeError("solve_rk45 needs a flat fp32 CUDA state (libcsd_b200 has no CPU path)")
    n = y0.numel()
    dev = y0.device
    direction = 1.0 if t_bound >= t0 else -1.0
    y = y0.clone()
    y_new = torch.empty_like(y)
    y_stage = torch.empty_like(y)
    ks = torch.empty(7, n, device=dev, dtype=torch.float32)
    norm_buf = K.reduce_workspace(dev)                   # [0] = the error norm's sum of squares (deterministic)
    nfev = 0

    def rms(stages, e, h, ya, yb):
        K.rk_error_sumsq(ks, stages, e, h, ya, yb, atol, rtol, norm_buf)
        return math.sqrt(norm_buf[0].item() / n)

    ks[0].copy_(fun(t0, y))
    nfev += 1
    # ---- select_initial_step (scipy/integrate/_ivp/common.py) ----
    ks[1].copy_(y)                                      # row 1 = y0 for the d0 norm
    d0 = rms(2, [0.0, 1.0], 1.0, y, y)
    d1 = rms(1, [1.0], 1.0, y, y)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    K.rk_combine(y, ks, 1, [1.0], h0 * direction, y_stage)
    ks[1].copy_(fun(t0 + h0 * direction, y_stage))
    nfev += 1
    d2 = rms(2, [-1.0, 1.0], 1.0, y, y) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / (_ORDER + 1))
    h_abs = min(100 * h0, h1, abs(t_bound - t0))        # scipy >= 1.10 clamps the first step to the interval length

    t = float(t0)
    steps = 0
    while direction * (t - t_bound) < 0:
        steps += 1
        if steps > max_steps:
            raise RuntimeError("solve_rk45: step limit reached")
        min_step = 10 * abs(np.nextafter(t, direction * np.inf) - t)
        h_abs = max(h_abs, min_step)
        step_accepted, step_rejected = False, False
        while not step_accepted:
            if h_abs < min_step:
                raise RuntimeError("solve_rk45: required step size is less than spacing between numbers")
            h = h_abs * direction
            t_new = t + h
            if direction * (t_new - t_bound) > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = abs(h)
            # ---- rk_step: stages 1..5, the 5th-order solution, and f(t_new, y_new) as the FSAL stage ----
            for s in range(1, 6):
                K.rk_combine(y, ks, s, _A[s], h, y_stage)
                ks[s].copy_(fun(t + _C[s] * h, y_stage))
            K.rk_combine(y, ks, 6, _B, h, y_new)
            ks[6].copy_(fun(t + h, y_new))
            nfev += 6
            err = rms(7, _E, h, y, y_new)
            if err < 1:
                factor = _MAX_FACTOR if err == 0 else min(_MAX_FACTOR, _SAFETY * err ** (-1.0 / (_ORDER + 1)))
                if step_rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                step_accepted = True
            else:
                h_abs *= max(_MIN_FACTOR, _SAFETY * err ** (-1.0 / (_ORDER + 1)))
                step_rejected = True
        t = t_new
        y, y_new = y_new, y
        ks[0].copy_(ks[6])                               # first-same-as-last
    return y, nfev
