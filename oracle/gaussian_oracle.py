"""TEST INFRASTRUCTURE ONLY (never imported by the product path) — CPU restatement of the reference's Gaussian-diffusion
arithmetic, pinned against tests/golden/gaussian.pt (generated from the unmodified reference by
oracle/make_golden_gaussian.py):
  schedules / respacing   gaussian_diffusion.py:72-194, modelizations/utils.py:1-57
  add_noise               gaussian_diffusion.py:313-341   (fp64 tables -> .float() per sample, diffuse/utils.py:6-19)
  DDPM.step               samplers/gaussian_diffusion/ddpm.py (set_steps, _get_p_mean_var, step)
  DDIM.step               samplers/gaussian_diffusion/ddim.py:27-103
All per-sample coefficients are read from float64 tables and cast to float32 BEFORE any arithmetic, exactly like
`extract_into_tensor`; the element-wise math is float32 in the reference's operation order.
"""

from __future__ import annotations

import math

import torch


def variance_schedule(n_steps: int, schedule: str = "linear") -> torch.Tensor:
    if schedule == "linear":
        scale = 1000 / n_steps
        return torch.linspace(scale * 0.0001, scale * 0.02, n_steps, dtype=torch.float64)
    if schedule == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        return torch.tensor([min(1 - f((i + 1) / n_steps) / f(i / n_steps), 0.999) for i in range(n_steps)], dtype=torch.float64)
    raise NotImplementedError(schedule)


def space_timesteps(num_timesteps: int, section_counts, ddim: bool = False) -> set[int]:
    """The non-DDIM branch of the reference. (Its DDIM branch raises for every request that is not the identity —
    the `raise` sits inside the stride loop, utils.py:28-31 — so the only reachable DDIM result is range(num_timesteps).)"""
    if ddim:
        if section_counts == num_timesteps:
            return set(range(num_timesteps))
        raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
    counts = [int(x) for x in section_counts.split(",")] if isinstance(section_counts, str) else [section_counts]
    size_per, extra = num_timesteps // len(counts), num_timesteps % len(counts)
    start, steps = 0, []
    for i, c in enumerate(counts):
        size = size_per + (1 if i < extra else 0)
        if size < c:
            raise ValueError(f"cannot divide section of {size} steps into {c}")
        stride = 1 if c <= 1 else (size - 1) / (c - 1)
        cur = 0.0
        for _ in range(c):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


def respace(training_steps: int, n_steps: int, schedule: str, section_counts, ddim: bool):
    """-> (betas float64, timestep_map) as GaussianDiffusion.set_steps leaves them."""
    if n_steps != training_steps:
        section_counts = section_counts or n_steps
    betas = variance_schedule(training_steps, schedule)
    tmap: list[int] = []
    if section_counts:
        use = space_timesteps(training_steps, section_counts, ddim)
        ab = (1 - betas).cumprod(0)
        last = torch.tensor(1.0)
        nb = []
        for i, a in enumerate(ab):
            if i in use:
                nb.append(torch.ones_like(a) - a / last)
                last = a
                tmap.append(i)
        betas = torch.tensor(nb)  # note: the reference rebuilds the table through a python list -> float32
    return betas, tmap


def _ex(arr: torch.Tensor, t: torch.Tensor, ndim: int) -> torch.Tensor:
    r = arr[t.long()].float()
    return r.view(-1, *([1] * (ndim - 1)))


def add_noise(betas: torch.Tensor, x: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    ab = (torch.ones_like(betas) - betas).cumprod(0)
    return _ex(ab.sqrt(), t, x.dim()) * x + (torch.ones_like(noise) - _ex(ab, t, x.dim())).sqrt() * noise


class Tables:
    def __init__(self, betas: torch.Tensor):
        one = lambda a: torch.ones_like(a)  # noqa: E731
        self.betas = betas
        self.alphas = one(betas) - betas
        self.ab = self.alphas.cumprod(0)
        self.ab_prev = torch.cat([torch.tensor([1.0], dtype=torch.float64), self.ab[:-1]])
        self.sqrt_ab = self.ab.sqrt()
        self.post_var = betas * (one(self.ab_prev) - self.ab_prev) / (one(self.ab) - self.ab)
        self.post_logvar = torch.log(torch.cat([self.post_var[1:2], self.post_var[1:]]))
        self.c1 = betas * self.ab_prev.sqrt() / (one(self.ab) - self.ab)
        self.c2 = (one(self.ab_prev) - self.ab_prev) * self.alphas.sqrt() / (one(self.ab) - self.ab)


def x_start(T: Tables, mean_type: str, pred, xt, t, clamp: bool):
    n = xt.dim()
    if mean_type == "epsilon":
        x0 = (1.0 / _ex(T.sqrt_ab, t, n)) * xt - ((torch.ones_like(pred) - _ex(T.ab, t, n)).sqrt() / _ex(T.sqrt_ab, t, n)) * pred
    elif mean_type == "xstart":
        x0 = pred
    else:
        x0 = (1.0 / _ex(T.c1, t, n)) * pred - (_ex(T.c2, t, n) / _ex(T.c1, t, n)) * xt
    return torch.clamp(x0, -1, 1) if clamp else x0


def ddpm_step(betas, mean_type, var_type, pred, xt, t, noise, clamp=False):
    T = Tables(betas)
    n = xt.dim()
    log_var = None
    if var_type in ("learned", "learned_range"):  # ddpm.py:268-270: model output = [mean prediction | variance head] on dim 1
        pred, log_var = torch.chunk(pred, 2, dim=1)
    x0 = x_start(T, mean_type, pred, xt, t, clamp)
    mean = _ex(T.c1, t, n) * x0 + _ex(T.c2, t, n) * xt
    if var_type == "fixed_small":
        var, lv = _ex(T.post_var, t, n), _ex(T.post_logvar, t, n)
    elif var_type == "fixed_large":
        seq = torch.cat([T.post_var[1:2], T.betas[1:]])
        var, lv = _ex(seq, t, n), _ex(torch.log(seq), t, n)
    elif var_type == "learned":  # ddpm.py:213-215: the model's second channel half is log sigma^2
        lv = log_var
        var = lv.exp()
    else:  # learned_range, ddpm.py:217-223: interpolate between the clipped posterior log-variance and log beta_t
        min_log, max_log = _ex(T.post_logvar, t, n), _ex(T.betas, t, n).log()
        w = (log_var + 1) / 2
        lv = w * max_log + (1 - w) * min_log
        var = lv.exp()
    var, lv = var.expand_as(xt), lv.expand_as(xt)
    mask = (t > 0).float().view(-1, *([1] * (n - 1)))
    x_prev = mean + mask * noise * torch.exp(0.5 * lv)
    vs = var.clamp_min(1e-20)
    logprob = (-((x_prev - mean) ** 2) / (2.0 * vs) - torch.log(2 * torch.pi * vs) * 0.5) * mask
    return {"x_prev": x_prev, "estimated_x0": x0, "x_prev_mean": mean, "x_prev_std": vs.sqrt(), "logprob": logprob}


def ddim_step(betas, mean_type, pred, xt, t, noise, clamp=False, eta=0.0):
    T = Tables(betas)
    n = xt.dim()
    x0 = x_start(T, mean_type, pred, xt, t, clamp)
    one = torch.ones_like(xt)
    eps = ((1 / _ex(T.sqrt_ab, t, n)) * xt - x0) / (1 / _ex(T.ab, t, n) - 1).sqrt()
    abp, ab = _ex(T.ab_prev, t, n), _ex(T.ab, t, n)
    sigma = eta * ((one - abp) / (one - ab)).sqrt() * (one - ab / abp).sqrt()
    mean = x0 * abp.sqrt() + (one - abp - sigma**2).sqrt() * eps
    mask = (t > 0).float().view(-1, *([1] * (n - 1)))
    x_prev = mean + mask * sigma * noise
    out = {"x_prev": x_prev, "estimated_x0": x0, "x_prev_mean": mean}
    if eta > 0:
        out["x_prev_std"] = sigma
        out["logprob"] = -((x_prev - mean) ** 2 / (2 * sigma**2) + torch.log(sigma) + 0.5 * torch.log(torch.tensor(2 * torch.pi)))
    return out


def euler_maruyama_step(eta: float, tmax: float, x_t, v, t_curr: float, t_prev: float, noise=None, x_prev=None):
    """samplers/flow/euler_meruyama.py:24-57 (python-float schedule scalars, fp32 tensors)."""
    sigma = ((t_curr / (1 - min(t_curr, tmax))) ** 0.5) * eta
    mean = x_t - (v + sigma**2 / (2 * t_curr) * (x_t + (1 - t_curr) * v)) * (t_curr - t_prev)
    std = torch.tensor(sigma * (t_curr - t_prev) ** 0.5)
    if x_prev is None:
        x_prev = mean + std * noise
    logprob = -((x_prev - mean) ** 2 / (2 * std**2) + torch.log(std) + 0.5 * torch.log(torch.tensor(2 * torch.pi)))
    return {"x_prev": x_prev, "x_prev_mean": mean, "x_prev_std": std, "estimated_x0": x_t - v * t_curr, "logprob": logprob}
