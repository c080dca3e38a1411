"""TEST INFRASTRUCTURE ONLY — generates tests/golden/gaussian.pt by running the UNMODIFIED reference (CPU, fp32):
GaussianDiffusion schedule tables / respacing / add_noise (gaussian_diffusion.py:72-341), DDPM.step and DDIM.step
(samplers/gaussian_diffusion/ddpm.py, ddim.py) with recorded noise, and a short denoise loop around a toy denoiser.

Run in the build container (where /root/reference is mounted):  python oracle/make_golden_gaussian.py
"""

from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "gaussian.pt")


class ToyDenoiser(torch.nn.Module):
    """eps_hat = a * x + b * sin(t / 100) (per sample), a / b learnable scalars: enough to drive the sampler loops."""

    classifier_free = True

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.3))
        self.b = torch.nn.Parameter(torch.tensor(-0.2))

    def forward(self, x, timesteps, p=0.0, **_):
        s = torch.sin(timesteps.float() / 100.0).view(-1, 1, 1, 1)
        return {"x": self.a * x + self.b * s * (0.5 if p == 1 else 1.0)}


def main():
    import_reference()
    from diffulab.diffuse.modelizations.gaussian_diffusion import GaussianDiffusion
    from diffulab.diffuse.samplers.gaussian_diffusion import DDIM, DDPM

    fx: dict = {"tables": [], "add_noise": [], "steps": [], "denoise": []}
    for kw in (dict(n_steps=1000, schedule="linear"), dict(n_steps=1000, schedule="cosine"), dict(n_steps=250, schedule="linear")):
        gd = GaussianDiffusion(**kw)
        fx["tables"].append({"kw": kw, "betas": gd.betas.clone(), "alphas_bar": gd.alphas_bar.clone(), "timestep_map": list(gd.timestep_map)})
    # (DDIM respacing is not exercised: the reference's space_timesteps(ddim=True) raises inside its first loop
    #  iteration for every n != training steps — modelizations/utils.py:28-31 — so there is nothing to pin.)
    for kw, n, sec in ((dict(n_steps=1000, sampling_method="ddpm"), 50, None),
                       (dict(n_steps=1000, sampling_method="ddpm", schedule="cosine"), 40, "10,15,15")):
        gd = GaussianDiffusion(**kw)
        gd.set_steps(n, schedule=kw.get("schedule", "linear"), section_counts=sec)
        fx["tables"].append({"kw": kw, "set_steps": (n, kw.get("schedule", "linear"), sec), "betas": gd.betas.clone(),
                             "alphas_bar": gd.alphas_bar.clone(), "timestep_map": list(gd.timestep_map)})

    g = torch.Generator().manual_seed(7)
    gd = GaussianDiffusion(n_steps=1000)
    x = torch.randn(6, 3, 8, 8, generator=g)
    noise = torch.randn(6, 3, 8, 8, generator=g)
    t = torch.tensor([0, 1, 17, 500, 998, 999], dtype=torch.int32)
    xt, _ = gd.add_noise(x, t, noise)
    fx["add_noise"].append({"x": x, "noise": noise, "t": t, "xt": xt})

    seed = 100
    for cls, name in ((DDPM, "ddpm"), (DDIM, "ddim")):
        for mean_type in ("epsilon", "xstart", "xprev"):
            for var_type in ("fixed_small", "fixed_large"):
                for clamp in (False, True):
                    for eta in ((0.0, 0.5) if name == "ddim" else (None,)):
                        gd = GaussianDiffusion(n_steps=1000, sampling_method=name, sampler_parameters=dict(mean_type=mean_type, var_type=var_type))
                        pred = torch.randn(5, 3, 8, 8, generator=g)
                        xt_ = torch.randn(5, 3, 8, 8, generator=g)
                        ts = torch.tensor([0, 1, 250, 640, 999], dtype=torch.int32)
                        seed += 1
                        torch.manual_seed(seed)
                        kwargs = {} if eta is None else {"eta": eta}
                        out = gd.sampler.step(model_prediction=pred, timesteps=ts, xt=xt_, clamp_x=clamp, **kwargs)
                        torch.manual_seed(seed)
                        noise_ = torch.randn_like(xt_)
                        fx["steps"].append({"sampler": name, "mean_type": mean_type, "var_type": var_type, "clamp": clamp, "eta": eta,
                                            "pred": pred, "xt": xt_, "t": ts, "noise": noise_, "seed": seed,
                                            "out": {k: v.clone() for k, v in out.items()}})

    model = ToyDenoiser()
    for name, n, gs, sargs in (("ddpm", 20, 0.0, {}), ("ddim", 100, 2.0, {"eta": 0.0}), ("ddim", 100, 0.0, {"eta": 0.3})):
        gd = GaussianDiffusion(n_steps=1000 if name == "ddpm" else n, sampling_method=name)
        if name == "ddpm":
            gd.set_steps(n)
        x0 = torch.randn(3, 1, 8, 8, generator=g)
        torch.manual_seed(1234)
        with torch.no_grad():
            out = gd.denoise(model, {"x": x0.clone()}, use_tqdm=False, clamp_x=True, guidance_scale=gs, sampler_args=sargs,
                             return_intermediates=True)
        fx["denoise"].append({"sampler": name, "n": n, "guidance": gs, "sampler_args": sargs, "x_init": x0, "seed": 1234,
                              "timestep_map": list(gd.timestep_map), "out": {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}})
    # training loss with the toy denoiser
    gd = GaussianDiffusion(n_steps=1000)
    x = torch.randn(4, 1, 8, 8, generator=g)
    noise = torch.randn(4, 1, 8, 8, generator=g)
    t = torch.tensor([3, 400, 777, 999], dtype=torch.int32)
    loss = gd.compute_loss(model, {"x": x.clone()}, t, noise)["loss"]
    loss.backward()
    fx["loss"] = {"x": x, "noise": noise, "t": t, "loss": loss.detach(), "grad_a": model.a.grad.clone(), "grad_b": model.b.grad.clone()}
    # ---- Euler-Maruyama flow sampler (samplers/flow/euler_meruyama.py) + a short stochastic Flow.denoise
    from diffulab.diffuse.modelizations.flow import Flow

    fx["em_steps"] = []
    for eta, n, idx, given in ((0.7, 10, 0, False), (0.7, 10, 4, False), (0.3, 25, 24, False), (1.0, 10, 9, True)):
        fl = Flow(n_steps=n, sampling_method="euler_maruyama", sampler_parameters={"eta": eta})
        x_t = torch.randn(3, 4, 8, 8, generator=g)
        v = torch.randn(3, 4, 8, 8, generator=g)
        t_curr, t_prev = fl.timesteps[idx], fl.timesteps[idx + 1]
        xp_in = torch.randn(3, 4, 8, 8, generator=g) if given else None
        seed += 1
        torch.manual_seed(seed)
        out = fl.sampler.step(x_t, v, t_curr, t_prev, x_prev=xp_in)
        torch.manual_seed(seed)
        noise_ = torch.randn_like(x_t)
        fx["em_steps"].append({"eta": eta, "n": n, "idx": idx, "x_t": x_t, "v": v, "x_prev_in": xp_in, "noise": noise_,
                               "t_curr": t_curr, "t_prev": t_prev, "out": {k: v_.clone() for k, v_ in out.items()}})

    class ToyFlow(ToyDenoiser):
        def forward(self, x, timesteps, p=0.0, **_):
            s = torch.sin(timesteps.float() * 3.0).view(-1, 1, 1, 1)
            return {"x": self.a * x + self.b * s * (0.5 if p == 1 else 1.0)}

    fl = Flow(n_steps=8, sampling_method="euler_maruyama", sampler_parameters={"eta": 0.5})
    x0 = torch.randn(2, 1, 8, 8, generator=g)
    torch.manual_seed(4321)
    out = fl.denoise(ToyFlow(), {"x": x0.clone()}, use_tqdm=False, guidance_scale=1.5, return_intermediates=True)
    fx["em_denoise"] = {"n": 8, "eta": 0.5, "guidance": 1.5, "x_init": x0, "seed": 4321,
                        "out": {k: v_.clone() for k, v_ in out.items()}}
    torch.save(fx, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB;", len(fx["steps"]), "step cases")


if __name__ == "__main__":
    main()
