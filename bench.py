#!/usr/bin/env python
"""Headline benchmark: DiT-XL/2 256px-latent flow-matching (+REPA) training throughput, img/s, on N B200s, and the
50-step Euler sampling sweep — the metric and configuration named in BASELINE.json (SURVEY.md 8d, cfg3 override).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches one rank per GPU for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the reference algorithm on host cores

A "step" is one full optimisation step (draw t -> add noise -> denoiser forward -> flow + REPA loss -> backward ->
bucketed gradient all-reduce -> AdamW) over one batch of synthetic latents of the config's shape.
`value`  : whole-job img/s with the batches already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the public API (`training_step`) with HOST pinned batches: the H2D copies of the
           step's inputs and the D2H read of its losses are inside the timed region.
`roofline`: achieved bf16 TFLOP/s of the dominant kernel (tcgen05 GEMM; algorithmic 2*M*N*K per launch / CUDA-event
           launch time inside the timed region) against the measured cuBLAS peak in MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (restatement of the reference algorithm, oracle/dit_oracle.py) timed on this box's host
           cores on a bounded sample of the same workload. The reference itself is pure Python and is not mounted on
           the GPU box, so both the baseline and `--impl reference` use the oracle port (kind "port").
Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DiT-XL/2 256px latent flow-matching train throughput"
UNIT = "img/s"
CONFIG = os.path.join(ROOT, "configs", "train_imagenet_flow_matching_repa.yaml")
FWD_GFLOP_PER_IMG = 313.33   # matmul FLOPs of one DiT-XL/2 forward
REPA_GFLOP_PER_IMG = 5.0     # projector 1152 -> 1024 -> 1024 -> 1024, forward + backward
TRAIN_GFLOP_PER_IMG = 3 * FWD_GFLOP_PER_IMG + REPA_GFLOP_PER_IMG
# DRAM traffic of the dominant kernel (tcgen05 GEMM), measured once with `ncu --set full` on the bench step (profiles/):
# mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the captured GEMM launches
NCU_GEMM_TRAFFIC_BYTES_PER_LAUNCH = 496.8e6


def load_peaks() -> tuple[dict, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batches(B: int, shape, repa_tokens: int, repa_dim: int, n_classes: int, seed: int, n: int, device=None, pinned=False):
    import torch

    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        x = torch.randn(B, *shape, generator=g)
        y = torch.randint(0, n_classes, (B,), generator=g)
        dst = torch.randn(B, repa_tokens, repa_dim, generator=g)
        if pinned:
            x, y, dst = x.pin_memory(), y.pin_memory(), dst.pin_memory()
        if device is not None:
            x, y, dst = x.to(device), y.to(device), dst.to(device)
        out.append((x, y, dst))
    return out


def make_batch(x, y, dst):
    return {"model_inputs": {"x": x, "y": y}, "extra": {"dst_features": dst}}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated on host cores (oracle port)
# ---------------------------------------------------------------------------------------------------------
def cpu_train_setup(cfg: dict, state_dict=None, repa_sd=None):
    import torch

    from diffulab_b200.config import instantiate

    torch.set_num_threads(os.cpu_count() or 1)
    if state_dict is None:
        model = instantiate(cfg["model"])
        state_dict = model.state_dict()
        repa_sd = instantiate(cfg["repa"]).state_dict()
    sd = {k: v.detach().float().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in state_dict.items()}
    rsd = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in repa_sd.items()}
    params = [v for v in list(sd.values()) + list(rsd.values()) if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=float(cfg["optimizer"]["lr"]), weight_decay=float(cfg["optimizer"]["weight_decay"]),
                            betas=tuple(cfg["optimizer"]["betas"]), eps=float(cfg["optimizer"]["eps"]))
    return sd, rsd, opt


def cpu_train_step(cfg: dict, sd, rsd, opt, x0, y, dst, p: float) -> float:
    """BaseTrainer.training_step (base_trainer.py:138-153) on the oracle: fp32, all host threads."""
    import torch

    from oracle import dit_oracle as O

    m = cfg["model"]
    hd = m["inner_dim"] // m["num_heads"]
    ocfg = dict(num_heads=m["num_heads"], patch_size=m["patch_size"], output_channels=m["output_channels"],
                rope_axes_dim=m.get("rope_axes_dim") or [hd // 2, hd // 2], rope_base=m.get("rope_base", 10000),
                frequency_embedding=256, n_classes=m["n_classes"])
    B = x0.shape[0]
    opt.zero_grad()
    t = torch.sigmoid(torch.randn(B)) if cfg["diffuser"]["extra_args"].get("logits_normal") else torch.rand(B)
    eps = torch.randn_like(x0)
    x_t = O.flow_add_noise(x0, t, eps)
    cap: dict = {}
    pred = O.mmdit_forward(sd, ocfg, x_t, t, y=y, p=p, draws={"label": torch.rand(B)}, capture=cap)
    loss = O.flow_loss(pred, x0, eps)
    r = cfg["repa"]
    loss = loss + O.repa_loss(rsd, cap[f"layers.{r['alignment_layer'] - 1}"], dst, r["coeff"])
    loss.backward()
    opt.step()
    return float(loss.detach())


def run_cpu_arm(args, cfg: dict, as_reference: bool, state_dict=None, repa_sd=None, budget_s: float = 25.0) -> dict:
    import torch

    shape = cfg["synthetic"]["image_shape"]
    sd, rsd, opt = cpu_train_setup(cfg, state_dict, repa_sd)
    p = float(cfg["trainer"]["p_classifier_free_guidance"])
    r = cfg["repa"]

    def batch(B):
        return synth_batches(B, shape, cfg["synthetic"]["repa_tokens"], r["embedding_dim"], cfg["model"]["n_classes"], 99, 1)[0]

    x, y, dst = batch(1)
    t0 = time.perf_counter()
    cpu_train_step(cfg, sd, rsd, opt, x, y, dst, p)  # warm-up + calibration at one image
    t1 = time.perf_counter() - t0
    if as_reference:
        n_steps, n_warm = args.steps, args.warmup
        total_budget = 170.0
        B = int(max(1, min(8, total_budget / max(1e-3, (n_steps + n_warm) * t1))))
    else:
        n_steps, n_warm = 2, 0
        B = int(max(1, min(8, budget_s / max(1e-3, n_steps * t1))))
    x, y, dst = batch(B)
    for _ in range(n_warm):
        cpu_train_step(cfg, sd, rsd, opt, x, y, dst, p)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        cpu_train_step(cfg, sd, rsd, opt, x, y, dst, p)
    dt = time.perf_counter() - t0
    return {"value": B * n_steps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_steps} full train steps (fwd+bwd+AdamW, fp32) of the same DiT-XL/2+REPA config at batch {B} on the host",
            "ms_per_step": 1e3 * dt / n_steps, "batch": B, "steps": n_steps}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import diffulab_b200 as dl
    from diffulab_b200 import _lib, ops
    from diffulab_b200.config import instantiate, load_config
    from diffulab_b200.training import FusedAdamW, GradReducer, training_step

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    _lib.check(_lib.load().dlb_device_check(), "dlb_device_check")

    overrides = [f"dataloader.batch_size={args.batch}"] + (args.override or [])
    cfg = load_config(args.config, overrides)
    B = int(cfg["dataloader"]["batch_size"])  # per-GPU batch (weak scaling: global = B * world)
    shape = cfg["synthetic"]["image_shape"]
    torch.manual_seed(1234 + rank)
    model = instantiate(cfg["model"]).to(device).train()
    repa = instantiate(cfg["repa"]).to(device)
    if world > 1:  # identical initial weights on every rank (DDP broadcasts rank 0's; same effect)
        for t in list(model.parameters()) + list(repa.parameters()):
            dist.broadcast(t.data, 0)
    repa.set_model(model)
    d = cfg["diffuser"]
    diffuser = dl.Diffuser(model, sampling_method=d["sampling_method"], model_type=d["model_type"], n_steps=d["n_steps"],
                           extra_args=d.get("extra_args", {}), extra_losses=[repa])
    opt = instantiate(cfg["optimizer"], params=list(model.parameters()) + list(repa.parameters()))
    assert isinstance(opt, FusedAdamW)
    reducer = GradReducer(stores=opt.stores, bucket_mb=args.bucket_mb) if world > 1 else None
    p_cfg = float(cfg["trainer"]["p_classifier_free_guidance"])
    n_params = sum(p.numel() for p in model.parameters())

    r = cfg["repa"]
    pool = 4
    host = synth_batches(B, shape, cfg["synthetic"]["repa_tokens"], r["embedding_dim"], cfg["model"]["n_classes"], 1234 + rank, pool, pinned=True)
    dev = [(x.to(device), y.to(device), dst.to(device)) for x, y, dst in host]

    def step_resident(i: int):
        return training_step(diffuser, opt, make_batch(*dev[i % pool]), p_cfg, reducer)

    def step_e2e(i: int) -> float:
        x, y, dst = host[i % pool]
        b = make_batch(x.to(device, non_blocking=True), y.to(device, non_blocking=True), dst.to(device, non_blocking=True))
        losses = training_step(diffuser, opt, b, p_cfg, reducer)
        return sum(float(v.item()) for v in losses.values())  # D2H read of the step's result, every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(args.warmup):
        step_resident(i)
    barrier()

    # ---- timed region 1: inputs resident in HBM ----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.launch_count()
    ops.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        last = step_resident(i)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    prof = ops.profile_stop()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    value = B * world * args.steps / (ms / 1e3)
    final_losses = {k: float(v.item()) for k, v in last.items()}

    # ---- timed region 2: end to end through the public API with host batches -------------------------------
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = B * world * args.steps / e2e_s
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- roofline of the dominant kernel (tcgen05 GEMM) ---------------------------------------------------
    peaks, peak_src = load_peaks()
    g_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("gemm_"))
    g_fl = sum(v["work"] for k, v in prof.items() if k.startswith("gemm_"))
    g_calls = sum(v["calls"] for k, v in prof.items() if k.startswith("gemm_"))
    all_ms = sum(v["ms"] for v in prof.values())
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    peak = float(peaks["bf16_tflops_sustained"])
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (fwd + dgrad + wgrad launches)", "achieved": round(achieved, 1),
                "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4), "traffic": NCU_GEMM_TRAFFIC_BYTES_PER_LAUNCH,
                "traffic_source": "dram__bytes_read+write per launch, mean over the 10 GEMM launches of profiles/ncu_top_kernels_r1_final.txt (ncu --set full)",
                "peak_source": f"bf16_tflops_sustained, {peak_src}",
                "launches": int(g_calls), "gemm_ms_per_step": round(g_ms / args.steps, 3),
                "gemm_share_of_kernel_time": round(g_ms / all_ms, 4) if all_ms else None,
                "step_model_flops_frac": round(value / world * TRAIN_GFLOP_PER_IMG * 1e9 / (peak * 1e12), 4)}
    breakdown = {k: {"calls_per_step": v["calls"] / args.steps, "ms_per_step": round(v["ms"] / args.steps, 3),
                     **({"tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1)} if v["work"] > 0 and v["ms"] > 0 else {})}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- 50-step Euler sampling sweep (sharded batch, no collective) --------------------------------------
    sample = None
    if not args.no_sample:
        model.eval()
        sb = args.sample_batch
        yb = torch.randint(0, cfg["model"]["n_classes"], (sb,), device=device)
        diffuser.set_steps(50)
        sample = {}
        for g_scale in (0.0, 4.0):
            x_init = torch.randn(sb, *shape, device=device)
            diffuser.generate({"x": x_init.clone(), "y": yb}, use_tqdm=False, guidance_scale=g_scale)  # warm-up
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            diffuser.generate({"x": x_init.clone(), "y": yb}, use_tqdm=False, guidance_scale=g_scale)
            s1.record()
            barrier()
            sms = max_over_ranks(s0.elapsed_time(s1))
            sample[f"euler50_cfg{g_scale:g}_img_per_s"] = round(sb * world / (sms / 1e3), 2)
        sample["batch_per_gpu"] = sb
        model.train()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sdm = {k: v for k, v in model.state_dict().items()}
        cpu = run_cpu_arm(args, cfg, as_reference=False, state_dict=sdm, repa_sd=repa.state_dict())

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "train_imagenet_flow_matching_repa: DiT-XL/2 (d=1152, depth 28, 16 heads, patch 2) on 32x32x4 "
                                   "latents + REPA (layer 8, 1024-d targets), AdamW, p_cfg=0.1",
                       "per_gpu_batch": B, "global_batch": B * world, "params_M": round(n_params / 1e6, 1), "parallelism": f"dp{world}",
                       "l2": "no explicit flush: every step streams >40 GB of activations and 3.3 GB of weights, far above the 126 MB L2",
                       "train_gflop_per_img": TRAIN_GFLOP_PER_IMG},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 * len(final_losses),
                    "ms_per_step": round(1e3 * e2e_s / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "losses": final_losses,
            "sample": sample, "kernel_breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from diffulab_b200.config import load_config

    cfg = load_config(args.config, args.override or [])
    res = run_cpu_arm(args, cfg, as_reference=True)
    line = {"impl": "reference", "metric": METRIC, "value": round(res["value"], 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(res["ms_per_step"], 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "train_imagenet_flow_matching_repa: DiT-XL/2 + REPA, reference algorithm on host cores (CPU oracle port)",
                       "per_step_sample_batch": res["batch"]},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": round(res["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=CONFIG)
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--bucket-mb", type=float, default=256.0)
    ap.add_argument("--sample-batch", type=int, default=64)
    ap.add_argument("--no-sample", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--override", action="append", help="extra key=value config override (repeatable)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
