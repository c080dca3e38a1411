"""CPU: the oracle's PerceiverResampler restatement against outputs of the UNMODIFIED reference (tests/golden/perceiver.pt, made
by oracle/make_golden_perceiver.py), including the recorded fact that the reference's default path raises."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "perceiver.pt")


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def test_oracle_perceiver_matches_reference():
    from oracle import dit_oracle as O

    fx = torch.load(GOLDEN, map_location="cpu", weights_only=False)
    O.set_round(None)
    for c in fx["cases"]:
        assert c["default_path_raises"] is True  # the un-batched position-id defect of the reference, kept on record
        kw = c["kw"]
        sd = {k: v.clone().requires_grad_(True) for k, v in c["state_dict"].items()}
        x = c["x"].clone().requires_grad_(True)
        out = O.perceiver_resampler(sd, x, kw["num_heads"], kw["head_dim"], c["rope_axes_dim"], c["rope_base"])
        assert rel_l2(out, c["out"]) < 2e-5
        out.backward(c["gout"])
        assert rel_l2(x.grad, c["dx"]) < 1e-4
        for k, g in c["grads"].items():
            assert rel_l2(sd[k].grad, g) < 1e-4, k
