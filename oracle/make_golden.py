"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference (CPU, fp32).

Run in the build container (where /root/reference is mounted):  python oracle/make_golden.py
Each fixture stores: constructor kwargs, the full state_dict (reference init, then every zero-initialised
modulation / mask-token parameter re-randomised with N(0, 0.02) — otherwise adaLN-Zero makes every block the
identity), the inputs, the uniform draws consumed by torch.rand inside the forward, and the
reference outputs: denoiser output, per-block features, loss dict (flow + REPA), selected gradients, SPRINT
kept indices and a short Euler trajectory with classifier-free guidance.
"""

from __future__ import annotations

import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rerandomize(model: torch.nn.Module, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "modulation" in name.lower() or name == "mask_token" or "qk_norm" in name or ".norm" in name or "_norm_" in name:
                if p.abs().sum() == 0:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                elif "scale" in name or name.endswith("weight") and p.dim() == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)  # LN / RMS scales away from exactly 1
                elif p.dim() == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.02)
            elif name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)


def make_embedder(ref, D_txt: int, L: int, valid: int, seed: int):
    from diffulab.networks.embedders.precomputed import PrecomputedEmbedder

    g = torch.Generator().manual_seed(seed)
    null = torch.randn(L, D_txt, generator=g)
    with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as f:
        torch.save(null, f.name)
        path = f.name
    emb = PrecomputedEmbedder(path, valid)
    os.unlink(path)
    return emb, null


def synth_context(B: int, L: int, D_txt: int, g: torch.Generator):
    lens = torch.randint(2, L + 1, (B,), generator=g)
    return {"embeddings": torch.randn(B, L, D_txt, generator=g), "attn_mask": torch.arange(L)[None, :] < lens[:, None]}


def capture_blocks(model, names):
    feats, handles = {}, []
    for lst in names:
        for i, layer in enumerate(getattr(model, lst)):
            def hook(_m, _i, out, key=f"{lst}.{i}"):
                feats[key] = (out[0] if isinstance(out, tuple) else out).detach().clone()
            handles.append(layer.register_forward_hook(hook))
    return feats, handles


def run_case(name: str, build, kwargs: dict, B: int, C: int, HW: int, mm: bool, train_mode: bool, p: float, seed: int,
             repa: dict | None = None, euler: dict | None = None):
    if _ONLY and name not in _ONLY:
        return
    ref = import_reference()
    from diffulab.diffuse.modelizations.flow import Flow

    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed + 1)
    extra = {}
    null = None
    if mm:
        D_txt, L = 48, 12
        embedder, null = make_embedder(ref, D_txt, L, 5, seed + 2)
        model = build(context_embedder=embedder, **kwargs)
        context = synth_context(B, L, D_txt, g)
        extra["initial_context"] = context
    else:
        model = build(**kwargs)
        extra["y"] = torch.randint(0, kwargs["n_classes"], (B,), generator=g)
    rerandomize(model, seed + 3)
    model.train(train_mode)
    x0 = torch.randn(B, C, HW, HW, generator=g)
    eps = torch.randn(B, C, HW, HW, generator=g)
    t = torch.rand(B, generator=g) * 0.9 + 0.05
    flow = Flow(n_steps=4, sampling_method="euler", shift=euler.get("shift") if euler else None)
    lists = [n for n in ("layers", "deep_layers", "decoder_layers") if hasattr(model, n)]
    feats, handles = capture_blocks(model, lists)

    # replay of the torch.rand draws consumed by the forward (order = call order inside the reference)
    fwd_seed = seed + 7
    draws = {}
    torch.manual_seed(fwd_seed)
    if p > 0:
        draws["context" if mm else "label"] = torch.rand(B)
    is_sprint = hasattr(model, "mask_token")
    if is_sprint and train_mode:
        p_sz = kwargs["patch_size"]
        S = (HW // p_sz) ** 2
        draws["scores"] = torch.rand((B, S), dtype=torch.float32)
        if 0 < p < 1:
            draws["path"] = torch.rand(B)

    losses = []
    if repa is not None:
        from diffulab.training.losses.repa import RepaLoss

        torch.manual_seed(seed + 11)
        rl = RepaLoss(load_dino=False, use_resampler=False, **repa)
        rl.set_model(model)
        losses.append(rl)
        S_img = (HW // kwargs["patch_size"]) ** 2
        dst = torch.randn(B, S_img, repa["embedding_dim"], generator=g)
    model.zero_grad()
    torch.manual_seed(fwd_seed)
    inputs = {"x": x0.clone(), "p": p, **extra}
    loss_dict = flow.compute_loss(model, inputs, t, noise=eps.clone(), extra_losses=losses,
                                  extra_args={"dst_features": dst} if repa is not None else {})
    x_t = inputs["x"].detach().clone()
    total = sum(loss_dict.values())
    total.backward()
    for h in handles:
        h.remove()

    # plain forward output at the same x_t (same draws)
    torch.manual_seed(fwd_seed)
    with torch.no_grad():
        pred = model(x_t, t, p=p, **extra)["x"]

    grad_keys = [k for k in ["conv_proj.weight", "conv_proj_encoder.weight", "layers.0.attention.qkv.weight",
                             "layers.0.attention.qkv_input.weight", "layers.1.modulation.lin.weight",
                             "layers.0.modulation_input.lin.bias", "last_layer.linear.weight", "layers.0.norm_1.weight",
                             "layers.0.attention.qk_norm.query_norm.scale", "mask_token", "fuse.weight",
                             "decoder_layers.0.modulation.lin.weight", "label_embed.embedding.weight",
                             "time_embed.0.weight", "context_embed.weight", "deep_layers.0.modulation.1.weight"]
                 if k in dict(model.named_parameters())]
    params = dict(model.named_parameters())
    grads = {k: (params[k].grad.detach().clone() if params[k].grad is not None else None) for k in grad_keys}

    fixture = {
        "name": name, "kwargs": kwargs, "mm": mm, "train_mode": train_mode, "p": p,
        "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "x0": x0, "eps": eps, "t": t, "x_t": x_t, "draws": draws,
        "context": extra.get("initial_context"), "y": extra.get("y"), "null_embedding": null, "null_valid": 5,
        "pred": pred, "features": feats,
        "losses": {k: v.detach().clone() for k, v in loss_dict.items()}, "grads": grads,
    }
    if repa is not None:
        fixture["repa"] = {"kwargs": repa, "state_dict": {k: v.detach().clone() for k, v in losses[0].state_dict().items()},
                           "dst": dst, "proj_grad": losses[0].proj[0].weight.grad.detach().clone()}
    if euler is not None:
        # short CFG Euler trajectory in eval mode (Flow.denoise flow.py:410-524)
        model.eval()
        flow.set_steps(euler["n_steps"], shift=euler.get("shift"))
        x_init = torch.randn(B, C, HW, HW, generator=g)
        inp = {"x": x_init.clone(), **extra}
        out = flow.denoise(model, inp, use_tqdm=False, guidance_scale=euler["guidance"], return_intermediates=True)
        fixture["euler"] = {**euler, "x_init": x_init, "x_final": out["x"].clone(), "xt": out["xt"].clone(),
                            "timesteps": list(flow.timesteps)}
    os.makedirs(OUT, exist_ok=True)
    torch.save(fixture, os.path.join(OUT, f"{name}.pt"))
    print(name, {k: float(v) for k, v in fixture["losses"].items()}, f"{os.path.getsize(os.path.join(OUT, name + '.pt')) / 1e6:.2f} MB")


_ONLY: set[str] = set(sys.argv[1:])  # optional fixture-name filter: python oracle/make_golden.py sprint_dit_cfg


def main():
    import_reference()
    from diffulab.networks.denoisers.ddt import DDT
    from diffulab.networks.denoisers.mmdit import MMDiT
    from diffulab.networks.denoisers.sprint import SprintDiT

    dit_kw = dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=64, embedding_dim=64, num_heads=2,
                  mlp_ratio=4, patch_size=2, depth=3, n_classes=10, classifier_free=True)
    run_case("dit_small_p0", MMDiT, dit_kw, B=3, C=4, HW=8, mm=False, train_mode=True, p=0.0, seed=0,
             repa=dict(alignment_layer=2, denoiser_dimension=64, hidden_dim=96, embedding_dim=40, coeff=0.5),
             euler=dict(n_steps=4, guidance=4.0, shift=None))
    run_case("dit_small_cfgdrop", MMDiT, dit_kw, B=4, C=4, HW=8, mm=False, train_mode=True, p=0.5, seed=1)
    # head_dim 72 (the DiT-XL/2 head size, not a multiple of 16/64) with partial rotary dims
    run_case("dit_hd72", MMDiT, dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=144, embedding_dim=80,
                                     num_heads=2, mlp_ratio=4, patch_size=2, depth=2, n_classes=7, classifier_free=True,
                                     rope_axes_dim=[32, 32]),
             B=2, C=4, HW=12, mm=False, train_mode=True, p=0.0, seed=2)
    mm_kw = dict(simple_dit=False, input_channels=8, output_channels=8, inner_dim=64, embedding_dim=64, num_heads=2,
                 mlp_ratio=4, patch_size=1, depth=3, n_single_stream_blocks=1, rope_axes_dim=[8, 12, 12], rope_base=2000,
                 classifier_free=True)
    run_case("mmdit_small", MMDiT, mm_kw, B=3, C=8, HW=4, mm=True, train_mode=True, p=0.4, seed=3,
             repa=dict(alignment_layer=2, denoiser_dimension=64, hidden_dim=96, embedding_dim=40, coeff=0.5),
             euler=dict(n_steps=3, guidance=2.0, shift=4.63))
    sp_kw = dict(simple_dit=False, input_channels=8, output_channels=8, inner_dim=64, embedding_dim=64, num_heads=2,
                 mlp_ratio=4, patch_size=1, encoder_depth=2, deep_layers_depth=2, n_single_stream_blocks=2, decoder_depth=2,
                 rope_axes_dim=[8, 12, 12], rope_base=2000, classifier_free=True, drop_rate=0.75)
    run_case("sprint_mm_train", SprintDiT, sp_kw, B=3, C=8, HW=4, mm=True, train_mode=True, p=0.3, seed=4,
             repa=dict(alignment_layer=2, denoiser_dimension=64, hidden_dim=96, embedding_dim=40, coeff=0.5),
             euler=dict(n_steps=3, guidance=4.0, shift=6.93))
    run_case("sprint_dit_train", SprintDiT, dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=64,
                                                  embedding_dim=64, num_heads=2, mlp_ratio=4, patch_size=2, encoder_depth=1,
                                                  deep_layers_depth=2, decoder_depth=1, n_classes=10, classifier_free=True,
                                                  drop_rate=0.5),
             B=3, C=4, HW=8, mm=False, train_mode=True, p=0.0, seed=5)
    # label-conditioned SprintDiT sampled with classifier-free guidance: the p = 1 pass skips the deep layers (path-drop
    # guidance, sprint.py:474-475), which a batched [y; null] evaluation must NOT be substituted for
    run_case("sprint_dit_cfg", SprintDiT, dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=64,
                                                embedding_dim=64, num_heads=2, mlp_ratio=4, patch_size=2, encoder_depth=1,
                                                deep_layers_depth=2, decoder_depth=1, n_classes=10, classifier_free=True,
                                                drop_rate=0.5),
             B=3, C=4, HW=8, mm=False, train_mode=True, p=0.3, seed=8, euler=dict(n_steps=3, guidance=4.0, shift=6.93))
    ddt_kw = dict(simple_ddt=False, input_channels=8, output_channels=8, inner_dim=64, num_heads=2, mlp_ratio=4, patch_size=1,
                  encoder_depth=3, n_single_stream_blocks=1, decoder_depth=2, rope_axes_dim=[8, 12, 12], rope_base=1000,
                  classifier_free=True)
    run_case("ddt_mm", DDT, ddt_kw, B=2, C=8, HW=4, mm=True, train_mode=True, p=0.0, seed=6,
             repa=dict(alignment_layer=2, denoiser_dimension=64, hidden_dim=96, embedding_dim=40, coeff=0.5))
    run_case("ddt_simple", DDT, dict(simple_ddt=True, input_channels=4, output_channels=4, inner_dim=64, num_heads=2,
                                     mlp_ratio=4, patch_size=2, encoder_depth=2, decoder_depth=2, n_classes=10,
                                     classifier_free=True),
             B=2, C=4, HW=8, mm=False, train_mode=True, p=0.0, seed=7)


if __name__ == "__main__":
    main()
