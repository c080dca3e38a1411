"""TEST INFRASTRUCTURE ONLY — generates tests/golden/gaussian_learned.pt by running the UNMODIFIED reference (CPU, fp32):
DDPM.step with the learned-variance parameterisations ("learned": the model's second channel half is log sigma^2;
"learned_range": it interpolates between log beta_t and the clipped posterior log-variance), reference
diffuse/samplers/gaussian_diffusion/ddpm.py:186-228, 260-330.

Run in the build container (where /root/reference is mounted):  python oracle/make_golden_gaussian_learned.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "gaussian_learned.pt")


def main():
    import_reference()
    from diffulab.diffuse.modelizations.gaussian_diffusion import GaussianDiffusion

    g = torch.Generator().manual_seed(21)
    fx = {"steps": []}
    seed = 500
    for var_type in ("learned", "learned_range"):
        for mean_type in ("epsilon", "xstart", "xprev"):
            for clamp in (False, True):
                gd = GaussianDiffusion(n_steps=1000, sampling_method="ddpm", sampler_parameters=dict(mean_type=mean_type, var_type=var_type))
                pred = torch.randn(5, 6, 8, 8, generator=g)  # 2C channels: [mean prediction | log-variance head]
                pred[:, 3:] = pred[:, 3:].clamp(-1, 1) if var_type == "learned_range" else pred[:, 3:] * 0.5 - 3.0
                xt = torch.randn(5, 3, 8, 8, generator=g)
                ts = torch.tensor([0, 1, 250, 640, 999], dtype=torch.int32)
                seed += 1
                torch.manual_seed(seed)
                out = gd.sampler.step(model_prediction=pred, timesteps=ts, xt=xt, clamp_x=clamp)
                torch.manual_seed(seed)
                noise = torch.randn_like(xt)
                fx["steps"].append({"mean_type": mean_type, "var_type": var_type, "clamp": clamp, "pred": pred, "xt": xt, "t": ts, "noise": noise,
                                    "seed": seed, "out": {k: v.clone() for k, v in out.items()}})
    torch.save(fx, OUT)
    print("wrote", OUT, len(fx["steps"]), "cases")


if __name__ == "__main__":
    main()
