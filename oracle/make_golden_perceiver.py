"""TEST INFRASTRUCTURE ONLY — generates tests/golden/perceiver.pt by running the UNMODIFIED reference (CPU, fp32):
PerceiverResampler.forward (networks/repa/perceiver_resampler.py:90-252) and RepaLoss with use_resampler=True
(training/losses/repa.py:73-186). The reference's own default path raises IndexError (un-batched position ids indexed as
batched), so the tables are built here exactly as the reference builds them and passed with the batch
dimension its rotary helper expects — same arithmetic, sample-independent tables.

Run in the build container (where /root/reference is mounted):  python oracle/make_golden_perceiver.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "perceiver.pt")


def main():
    import_reference()
    from diffulab.networks.repa.perceiver_resampler import PerceiverResampler
    from diffulab.networks.utils.nn import get_cos_sin_ndim_grid

    fx = {"cases": []}
    for seed, kw, B, N in ((0, dict(dim=64, depth=2, head_dim=16, num_heads=4, ff_mult=2, num_latents=16), 3, 16),
                           (1, dict(dim=128, depth=1, head_dim=32, num_heads=2, ff_mult=4, num_latents=24, rope_axes_dim=[8, 8], rope_base=2000), 2, 64)):
        torch.manual_seed(seed)
        m = PerceiverResampler(**kw)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dim() == 1:
                    p.add_(torch.randn_like(p) * 0.05)
        g = torch.Generator().manual_seed(seed + 10)
        x = torch.randn(B, N, kw["dim"], generator=g).requires_grad_(True)
        hw = int(N**0.5)
        pos = torch.stack(torch.meshgrid(torch.arange(hw), torch.arange(hw), indexing="ij"), -1).view(-1, 2)
        cos, sin = get_cos_sin_ndim_grid(pos, base=m.rope_base, axes_dim=m.rope_axes_dim)
        raised = False
        try:
            m(x.detach())
        except (IndexError, RuntimeError):
            raised = True
        out = m(x, cos_sin=(cos[None].expand(B, -1, -1), sin[None].expand(B, -1, -1)))
        gout = torch.randn(out.shape, generator=g)
        out.backward(gout)
        grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if n in ("latents", "layers.0.0.to_kv.weight", "layers.0.0.to_q.weight",
                                                                                     "layers.0.0.norm_x.weight", "layers.0.1.1.weight", "norm.bias")}
        fx["cases"].append({"kw": kw, "state_dict": {k: v.detach().clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(),
                            "out": out.detach().clone(), "gout": gout, "dx": x.grad.detach().clone(), "grads": grads, "default_path_raises": raised,
                            "rope_axes_dim": list(m.rope_axes_dim), "rope_base": m.rope_base})
    torch.save(fx, OUT)
    print("wrote", OUT, [c["default_path_raises"] for c in fx["cases"]], [float(c["out"].abs().mean()) for c in fx["cases"]])


if __name__ == "__main__":
    main()
