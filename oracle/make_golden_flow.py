"""TEST INFRASTRUCTURE ONLY — generates tests/golden/flow_misc.pt by running the UNMODIFIED reference (CPU, fp32):
Flow.draw_timesteps under fixed seeds for every (logits_normal, shift, prediction_type) combination the shipped configs
use (diffuse/modelizations/flow.py:168-197, 84-99), Flow.set_steps schedules (flow.py:101-135) and Euler trajectories of
a toy denoiser through Flow.denoise with and without guidance (flow.py:410-524) — the a14 / a19 rows of SURVEY.md 8(a).

Run in the build container (where /root/reference is mounted):  python oracle/make_golden_flow.py
"""

from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "flow_misc.pt")


class ToyVelocity(torch.nn.Module):
    """v = a * x + b * cos(3 t) (+ c for the unconditional pass): enough to drive the sampler loops deterministically."""

    classifier_free = True

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.4))
        self.b = torch.nn.Parameter(torch.tensor(-0.3))

    def forward(self, x, timesteps, p=0.0, **_):
        s = torch.cos(3.0 * timesteps.float()).view(-1, 1, 1, 1)
        return {"x": self.a * x + self.b * s + (0.25 if p == 1 else 0.0)}


def main():
    import_reference()
    from diffulab.diffuse.modelizations.flow import Flow

    fx: dict = {"draws": [], "steps": [], "denoise": []}
    for kw in (dict(), dict(logits_normal=True), dict(logits_normal=True, shift=4.63), dict(shift=6.93),
               dict(logits_normal=True, prediction_type="x"), dict(shift=4.63, prediction_type="x")):
        for seed, B in ((0, 7), (1, 128), (2, 1)):
            flow = Flow(n_steps=10, sampling_method="euler", **kw)
            torch.manual_seed(seed)
            t = flow.draw_timesteps(B)
            fx["draws"].append({"kw": kw, "seed": seed, "B": B, "t": t.clone()})
    for n, shift in ((50, None), (50, 6.93), (100, None), (7, 4.63), (1, None)):
        flow = Flow(n_steps=4, sampling_method="euler")
        flow.set_steps(n, shift=shift)
        fx["steps"].append({"n": n, "shift": shift, "timesteps": list(flow.timesteps)})
    model = ToyVelocity()
    g = torch.Generator().manual_seed(11)
    x_init = torch.randn(3, 2, 4, 4, generator=g)
    for n, shift, guidance in ((8, None, 0.0), (8, 6.93, 0.0), (5, None, 2.0), (12, 4.63, 4.0)):
        flow = Flow(n_steps=n, sampling_method="euler")
        flow.set_steps(n, shift=shift)
        out = flow.denoise(model, {"x": x_init.clone()}, use_tqdm=False, guidance_scale=guidance, return_intermediates=True)
        fx["denoise"].append({"n": n, "shift": shift, "guidance": guidance, "x_init": x_init.clone(), "x": out["x"].detach().clone(),
                              "xt": out["xt"].detach().clone(), "estimated_x0": out["estimated_x0"].detach().clone()})
    torch.save(fx, OUT)
    print("wrote", OUT, {k: len(v) for k, v in fx.items()})


if __name__ == "__main__":
    main()
