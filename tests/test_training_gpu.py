"""Training-side parity on the GPU: an N-step loss curve of `training_step` (fused AdamW,
hand-scheduled backward) against the CPU oracle + torch.optim.AdamW on the same weights / timesteps / noise; the fused
AdamW's torch-compatible state_dict round trip; gradient-less parameters skipped as torch skips `grad is None`; the fused
EMA against a torch restatement of ema_pytorch's rule (the package is absent here: "parity unpinned" for EMA)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


KW = dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=128, embedding_dim=128, num_heads=2, mlp_ratio=4,
          patch_size=2, depth=2, n_classes=10, classifier_free=True)
OCFG = dict(num_heads=2, patch_size=2, output_channels=4, rope_axes_dim=[32, 32], rope_base=10000, frequency_embedding=256, n_classes=10)


def _model(seed=0):
    import diffulab_b200 as dl

    torch.manual_seed(seed)
    m = dl.MMDiT(**KW)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m


def test_loss_curve_matches_oracle_adamw(cuda_device):
    """20 optimisation steps: per-step flow + REPA loss of the CUDA path tracks oracle + torch.optim.AdamW (fp32 CPU) within
    2e-2 relative (bf16 compute vs fp32), and the trained weights stay close (N-step loss-curve parity)."""
    import diffulab_b200 as dl
    from diffulab_b200.training import FusedAdamW, training_step
    from oracle import dit_oracle as O

    model = _model()
    repa = dl.RepaLoss(load_dino=False, alignment_layer=1, denoiser_dimension=128, hidden_dim=96, embedding_dim=40, coeff=0.5)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    rsd = {k: v.detach().clone().requires_grad_(True) for k, v in repa.state_dict().items()}
    ref_opt = torch.optim.AdamW([v for v in list(sd.values()) + list(rsd.values()) if v.requires_grad], lr=1e-3, weight_decay=0.01)
    model, repa = model.cuda().train(), repa.cuda()
    repa.set_model(model)
    diffuser = dl.Diffuser(model, sampling_method="euler", n_steps=4, extra_args={"logits_normal": True}, extra_losses=[repa])
    opt = FusedAdamW(list(model.parameters()) + list(repa.parameters()), lr=1e-3, weight_decay=0.01)
    B, hw, n_steps = 4, 16, 20
    g = torch.Generator().manual_seed(9)
    O.set_round(None)
    curve, ref_curve = [], []
    for i in range(n_steps):
        x0 = torch.randn(B, 4, hw, hw, generator=g)
        y = torch.randint(0, 10, (B,), generator=g)
        dst = torch.randn(B, (hw // 2) ** 2, 40, generator=g)
        # the draws training_step will make: timesteps on the CPU generator, noise with randn_like on the CUDA generator
        torch.manual_seed(1000 + i)
        t = torch.sigmoid(torch.randn(B, dtype=torch.float32))
        eps = torch.randn_like(x0.cuda()).cpu()
        torch.manual_seed(1000 + i)
        out = training_step(diffuser, opt, {"model_inputs": {"x": x0.cuda(), "y": y.cuda()}, "extra": {"dst_features": dst.cuda()}}, 0.0)
        curve.append(sum(float(v.item()) for v in out.values()))
        ref_opt.zero_grad()
        cap: dict = {}
        pred = O.mmdit_forward(sd, OCFG, O.flow_add_noise(x0, t, eps), t, y=y, p=0.0, capture=cap)
        loss = O.flow_loss(pred, x0, eps) + O.repa_loss(rsd, cap["layers.0"], dst, 0.5)
        loss.backward()
        ref_opt.step()
        ref_curve.append(float(loss.detach()))
    for i, (a, b) in enumerate(zip(curve, ref_curve)):
        assert abs(a - b) <= 2e-2 * abs(b), (i, a, b)
    assert ref_curve[-1] < ref_curve[0]  # the run actually optimises
    params = dict(model.named_parameters())
    for k in ("layers.0.attention.qkv.weight", "layers.1.mlp_input.2.weight", "last_layer.linear.weight", "conv_proj.weight"):
        assert rel_l2(params[k].detach(), sd[k].detach()) < 2e-2, k


def test_fused_adamw_state_dict_round_trip(cuda_device):
    """save -> load into a fresh optimizer -> step continues bit-identically (moments and bias-correction step restored);
    the serialised layout is torch's (per-parameter step / exp_avg / exp_avg_sq) and loads into torch.optim.AdamW."""
    from diffulab_b200.training import FusedAdamW

    def params(seed):
        g = torch.Generator().manual_seed(seed)
        return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in ((64, 32), (100,), (7, 3, 2, 2))]

    gg = torch.Generator().manual_seed(5)
    grads = [[torch.randn(p.shape, generator=gg).cuda() for p in params(0)] for _ in range(5)]
    pa = params(0)
    a = FusedAdamW(pa, lr=1e-2, weight_decay=0.05)
    pt = [torch.nn.Parameter(p.detach().clone()) for p in params(0)]
    t = torch.optim.AdamW(pt, lr=1e-2, weight_decay=0.05)

    def run(opt, ps, gs):
        opt.zero_grad()
        for p, g in zip(ps, gs):
            if isinstance(opt, FusedAdamW):
                from diffulab_b200 import blocks as K
                K.gbuf(p).copy_(g)
            else:
                p.grad = g.clone()
        opt.step()

    for i in range(3):
        run(a, pa, grads[i])
        run(t, pt, grads[i])
    state = copy.deepcopy(a.state_dict())
    assert set(state["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(state["state"][0]["step"]) == 3.0
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    b = FusedAdamW(pb, lr=1e-2, weight_decay=0.05)
    b.load_state_dict(state)
    assert b._steps == [3]
    t2 = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in pa], lr=1e-2, weight_decay=0.05)
    t2.load_state_dict(state)  # torch accepts the layout
    for i in range(3, 5):
        run(a, pa, grads[i])
        run(b, pb, grads[i])
        run(t, pt, grads[i])
    for x, y, z in zip(pa, pb, pt):
        assert torch.equal(x.detach(), y.detach())
        assert rel_l2(x.detach(), z.detach()) < 1e-6
    for k in ("exp_avg", "exp_avg_sq"):
        assert rel_l2(b.state[pb[0]][k], t.state[pt[0]][k]) < 1e-6


def test_fused_adamw_skips_gradless_parameters(cuda_device):
    """torch.optim.AdamW leaves parameters whose .grad is None untouched (no weight decay, no moment update); parameters
    with an all-zero gradient DO decay. The reference's last dual block has such unused parameters."""
    from diffulab_b200 import blocks as K
    from diffulab_b200.training import FusedAdamW

    g = torch.Generator().manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in ((40, 8), (33,), (128,))]
    before = [p.detach().clone() for p in ps]
    opt = FusedAdamW(ps, lr=1e-2, weight_decay=0.1)
    pt = [torch.nn.Parameter(p.clone()) for p in before]
    ref = torch.optim.AdamW(pt, lr=1e-2, weight_decay=0.1)
    for _ in range(3):
        opt.zero_grad()
        ref.zero_grad()
        gr = torch.randn(ps[0].shape, generator=g).cuda()
        K.gbuf(ps[0]).copy_(gr)      # real gradient
        K.gbuf(ps[2]).zero_()        # touched, gradient exactly zero -> decays like torch (grad = zeros tensor)
        pt[0].grad, pt[2].grad = gr.clone(), torch.zeros_like(pt[2])
        opt.step()
        ref.step()
    assert torch.equal(ps[1].detach(), before[1])  # never received a gradient: untouched
    assert rel_l2(ps[0].detach(), pt[0].detach()) < 1e-6
    assert rel_l2(ps[2].detach(), pt[2].detach()) < 1e-6 and not torch.equal(ps[2].detach(), before[2])
    assert float(opt.state[ps[1]]["exp_avg"].abs().max()) == 0.0


def _ema_reference(history, beta, update_after_step, update_every, inv_gamma=1.0, power=2.0 / 3.0, min_value=0.0):
    """Plain-torch restatement of ema_pytorch.EMA.update() (0.7.7, as published) over a list of successive online weights."""
    ema, step, initted = None, 0, False
    for w in history:
        s = step
        step += 1
        if not initted:
            ema, initted = w.clone(), True
            continue
        if s % update_every != 0:
            continue
        if s <= update_after_step:
            ema = w.clone()
            continue
        epoch = max(step - update_after_step - 1, 0)
        decay = 0.0 if epoch <= 0 else min(max(1 - (1 + epoch / inv_gamma) ** -power, min_value), beta)
        ema = torch.lerp(ema, w, 1 - decay)
    return ema


@pytest.mark.parametrize("beta,after,every", [(0.999, 0, 10), (0.9999, 0, 1), (0.99, 5, 2)])
def test_fused_ema_matches_restatement(cuda_device, beta, after, every):
    """configs/trainer/default.yaml (update_every 10) and the txt_to_img configs (every 1, rate 0.9999)."""
    from diffulab_b200 import blocks as K
    from diffulab_b200.training import EMA, FusedAdamW

    g = torch.Generator().manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in ((48, 16), (65,))]
    opt = FusedAdamW(ps, lr=5e-2, weight_decay=0.0)
    ema = EMA(opt, beta=beta, update_after_step=after, update_every=every)
    history = []
    for _ in range(25):
        opt.zero_grad()
        for p in ps:
            K.gbuf(p).copy_(torch.randn(p.shape, generator=g).cuda())
        opt.step()  # EMA rides in the AdamW kernel
        history.append(torch.cat([p.detach().reshape(-1).cpu() for p in ps]))
    got = torch.cat([v.reshape(-1).cpu() for v in ema.state_dict().values()])
    ref = _ema_reference(history, beta, after, every)
    assert ema.step == 25
    assert rel_l2(got, ref) < 1e-6
    assert rel_l2(got, history[-1]) > 1e-4  # it is an average, not a copy


def test_device_prefetcher_and_loss_reader(cuda_device):
    """host batches arrive on the device in order and intact while copies run on a side stream; losses are read back one step late"""
    from diffulab_b200.training import DevicePrefetcher, LossReader

    g = torch.Generator().manual_seed(5)
    host = [{"model_inputs": {"x": torch.randn(4, 3, 8, 8, generator=g).pin_memory(), "y": torch.randint(0, 9, (4,), generator=g).pin_memory()},
             "extra": {"dst_features": torch.randn(4, 16, 8, generator=g).pin_memory()}, "tag": i} for i in range(5)]
    reader, seen, sums = LossReader(), [], []
    for b in DevicePrefetcher(host, "cuda"):
        i = b["tag"]
        assert b["model_inputs"]["x"].is_cuda and torch.equal(b["model_inputs"]["x"].cpu(), host[i]["model_inputs"]["x"])
        assert torch.equal(b["model_inputs"]["y"].cpu(), host[i]["model_inputs"]["y"]) and torch.equal(b["extra"]["dst_features"].cpu(), host[i]["extra"]["dst_features"])
        seen.append(i)
        prev = reader.push({"loss": b["model_inputs"]["x"].sum(), "aux": b["extra"]["dst_features"].mean()})
        if prev is not None:
            sums.append(prev)
    sums.append(reader.flush())
    assert seen == list(range(5)) and len(sums) == 5 and reader.flush() is None
    for i, s in enumerate(sums):
        assert abs(s["loss"] - float(host[i]["model_inputs"]["x"].sum())) < 1e-3 and abs(s["aux"] - float(host[i]["extra"]["dst_features"].mean())) < 1e-6
