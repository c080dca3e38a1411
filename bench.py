#!/usr/bin/env python
"""Headline benchmark: DiT-XL/2 256px-latent flow-matching (+REPA) training throughput, img/s, on N B200s, and the
50-step Euler sampling sweep — the metric and configuration named in BASELINE.json (SURVEY.md 8d, cfg3 override).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches one rank per GPU for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the reference algorithm on host cores
    python bench.py --config {cifar10,txt_to_img,sprint} ...  # the other BASELINE configs at their own shapes

A "step" is one full optimisation step (draw t -> add noise -> denoiser forward -> flow + REPA loss -> backward ->
bucketed gradient all-reduce -> AdamW) over one batch of synthetic inputs of the config's shape.
`value`  : whole-job img/s with the batches already resident in HBM (CUDA events, max over ranks). Nothing but the
           training steps runs inside this region (per-kernel profiling is a SEPARATE pass afterwards).
`e2e`    : the same metric through the public API (`training_step`) with HOST pinned batches: the H2D copies of the
           step's inputs and the D2H read of its losses are inside the timed region. Also carries the 50-step Euler
           sampling numbers (`sample_*`), measured through `Diffuser.generate` (CUDA-graph replayed loop).
`roofline`: achieved bf16 TFLOP/s of the dominant kernel (tcgen05 GEMM; algorithmic 2*M*N*K per launch / CUDA-event
           launch time in the profiled pass) against the measured cuBLAS peak in MEASURED_PEAKS.json; per-shape
           traffic from the committed ncu capture.
`cpu_baseline`: the CPU oracle (restatement of the reference algorithm, oracle/dit_oracle.py) timed on this box's host
           cores on a bounded sample of the same workload, plus `ref_gpu_*`: the same restatement on THIS GPU under bf16
           autocast with F.scaled_dot_product_attention (the reference's PyTorch GPU path — its real competitor),
           eager and optionally torch.compile. The reference itself is pure Python and is not mounted on the GPU box,
           so these arms use the oracle port (kind "port").
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

UNIT = "img/s"
METRICS = {
    "imagenet_repa": "DiT-XL/2 256px latent flow-matching train throughput",
    "cifar10": "DiT (d=512 x10) 32x32 pixel flow-matching train throughput",
    "txt_to_img": "DDT (d=640, 8+4) txt-to-img latent flow-matching train throughput",
    "sprint": "SprintDiT (d=768, 2+8+2) txt-to-img latent flow-matching train throughput",
}
WORKLOADS = {
    "imagenet_repa": "train_imagenet_flow_matching_repa: DiT-XL/2 (d=1152, depth 28, 16 heads, patch 2) on 32x32x4 latents + REPA "
                     "(layer 8, 1024-d targets), AdamW, p_cfg=0.1",
    "cifar10": "train_cifar10_flow_matching: DiT d=512 depth 10 (8 heads, patch 2) on 32x32x3 pixels, AdamW",
    "txt_to_img": "train_imagenet_repa_txt_to_img: DDT d=640 (8 dual-stream encoder + 4 per-token-modulated decoder blocks) on 16x16x128 "
                  "latents, 128 text tokens (ragged masks, 2048-d), REPA (layer 8, 384-d), shift 4.63, p_cfg=0.1",
    "sprint": "train_imagenet_repa_txt_to_img_sprint: SprintDiT d=768 (2 enc + 8 single-stream deep + 2 dec), 75% token drop, 16x16x128 "
              "latents, 128 text tokens, REPA (layer 2, 384-d), shift 4.63, p_cfg=0.1",
}
# DRAM traffic of the dominant kernel per shape, from ONE `ncu --set full` capture of the bench step (profiles/, r2):
# dram__bytes_read.sum + dram__bytes_write.sum per launch next to the algorithmic bytes of that shape (bf16 operands + output).
NCU_GEMM_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_gemm_traffic_r2.json")


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


def config_block(name: str, wl, B: int, world: int, n_params: int | None) -> dict:
    """The `config` object shared by our arm and the reference arm (same keys, same workload string)."""
    return {"workload": WORKLOADS.get(name, name), "per_gpu_batch": B, "global_batch": B * world, "params_M": round(n_params / 1e6, 1) if n_params else None,
            "parallelism": f"dp{world}", "l2": "no explicit flush: every step streams far more activation + weight bytes than the 126 MB L2",
            "train_gflop_per_img": round(wl.flops_per_image(True) / 1e9, 2)}


# ---------------------------------------------------------------------------------------------------------
# oracle arms (CPU baseline, reference-GPU arm): the reference algorithm restated (oracle/dit_oracle.py)
# ---------------------------------------------------------------------------------------------------------
def oracle_cfg_for(wl, device=None) -> dict:
    m = wl.cfg["model"]
    d, H = int(m["inner_dim"]), int(m["num_heads"])
    hd = d // H
    axes = m.get("rope_axes_dim") or [int(hd // (3 if wl.mm else 2))] * (3 if wl.mm else 2)
    cfg = dict(num_heads=H, patch_size=m["patch_size"], output_channels=m.get("output_channels") or m["input_channels"], rope_axes_dim=axes,
               rope_base=m.get("rope_base", 10000), frequency_embedding=m.get("frequency_embedding", 256), n_classes=m.get("n_classes"),
               drop_rate=m.get("drop_rate", 0.75))
    if wl.mm:
        import torch

        null = wl.null_embedding
        cfg["null_embedding"] = null.to(device) if device is not None else null
        mask = torch.arange(null.shape[0]) < int(wl.text["null_valid"])
        cfg["null_mask"] = mask.to(device) if device is not None else mask
    return cfg


def oracle_train_step(wl, ocfg, sd, rsd, opt, b, p: float, device, autocast: bool) -> float:
    """BaseTrainer.training_step (base_trainer.py:138-153) on the oracle: draw t, add noise, forward, flow (+REPA) loss,
    backward, AdamW. fp32 on the host cores, or bf16 autocast + fused SDPA on `device` (the reference's GPU numerics)."""
    import torch

    from oracle import dit_oracle as O

    kind = wl.cfg["model"]["_target_"].rsplit(".", 1)[-1]
    x0 = b["x"].to(device)
    B = x0.shape[0]
    opt.zero_grad()
    ea = wl.cfg["diffuser"].get("extra_args", {})
    t = torch.sigmoid(torch.randn(B)) if ea.get("logits_normal") else torch.rand(B)
    if ea.get("shift"):
        t = O.shift_timestep(t, float(ea["shift"]))
    t = t.to(device)
    eps = torch.randn_like(x0)
    y = b["y"].to(device) if "y" in b else None
    context = {k: v.to(device) for k, v in b["context"].items()} if "context" in b else None
    draws = {"label": torch.rand(B, device=device), "context": torch.rand(B, device=device)}
    with torch.autocast(device.type, dtype=torch.bfloat16, enabled=autocast):
        x_t = O.flow_add_noise(x0, t, eps)
        cap: dict = {}
        common = dict(y=y, context=context, p=p, draws=draws, capture=cap)
        if kind == "SprintDiT":
            C, H, W = wl.shape
            ps = int(wl.cfg["model"]["patch_size"])
            draws["scores"] = torch.rand(B, (H // ps) * (W // ps), device=device)
            draws["path"] = torch.rand(B, device=device)
            pred = O.sprint_forward(sd, ocfg, x_t, t, training=True, **common)
        elif kind == "DDT":
            pred = O.ddt_forward(sd, ocfg, x_t, t, **common)
        else:
            pred = O.mmdit_forward(sd, ocfg, x_t, t, **common)
        loss = O.flow_loss(pred, x0, eps)
        if rsd is not None:
            r = wl.cfg["repa"]
            loss = loss + O.repa_loss(rsd, cap[f"layers.{int(r['alignment_layer']) - 1}"], b["dst"].to(device), float(r["coeff"]))
    loss.backward()
    opt.step()
    return loss.detach()


def oracle_setup(wl, device, state_dict=None, repa_sd=None):
    import torch

    if state_dict is None:
        state_dict = wl.model.state_dict()
        repa_sd = wl.repa.state_dict() if wl.repa is not None else None
    sd = {k: v.detach().to(device=device, dtype=torch.float32 if v.is_floating_point() else v.dtype).clone().requires_grad_(v.is_floating_point())
          for k, v in state_dict.items()}
    rsd = {k: v.detach().float().to(device).clone().requires_grad_(True) for k, v in repa_sd.items()} if repa_sd is not None else None
    params = [v for v in list(sd.values()) + list((rsd or {}).values()) if v.requires_grad]
    o = wl.cfg["optimizer"]
    opt = torch.optim.AdamW(params, lr=float(o["lr"]), weight_decay=float(o["weight_decay"]), betas=tuple(o["betas"]), eps=float(o["eps"]))
    return sd, rsd, opt


def run_cpu_arm(args, wl, as_reference: bool, state_dict=None, repa_sd=None, budget_s: float = 25.0) -> dict:
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    cpu = torch.device("cpu")
    sd, rsd, opt = oracle_setup(wl, cpu, state_dict, repa_sd)
    ocfg = oracle_cfg_for(wl)
    g = torch.Generator().manual_seed(99)
    t0 = time.perf_counter()
    oracle_train_step(wl, ocfg, sd, rsd, opt, wl.batch(1, g), wl.p_cfg, cpu, False)  # warm-up + calibration at one image
    t1 = time.perf_counter() - t0
    if as_reference:
        n_steps, n_warm = args.steps, args.warmup
        B = int(max(1, min(8, 170.0 / max(1e-3, (n_steps + n_warm) * t1))))
    else:
        n_steps, n_warm = 2, 0
        B = int(max(1, min(8, budget_s / max(1e-3, n_steps * t1))))
    b = wl.batch(B, g)
    for _ in range(n_warm):
        oracle_train_step(wl, ocfg, sd, rsd, opt, b, wl.p_cfg, cpu, False)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        oracle_train_step(wl, ocfg, sd, rsd, opt, b, wl.p_cfg, cpu, False)
    dt = time.perf_counter() - t0
    return {"value": B * n_steps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_steps} full train steps (fwd+bwd+AdamW, fp32) of the same config at batch {B} on the host cores",
            "ms_per_step": 1e3 * dt / n_steps, "batch": B, "steps": n_steps}


def run_ref_gpu_arm(wl, device, B: int, mode: str, steps: int = 3) -> dict:
    """The reference's PyTorch GPU path on this box: oracle under bf16 autocast with F.scaled_dot_product_attention,
    torch.optim.AdamW (fused=False, the reference default), eager; `compile` additionally wraps each transformer block in
    torch.compile (inductor), the per-block equivalent of the reference's `torch.compile(denoiser)` (base_trainer.py:81-86)."""
    import torch

    from oracle import dit_oracle as O

    out: dict = {}
    g = torch.Generator().manual_seed(7)
    ocfg = oracle_cfg_for(wl, device)
    ocfg["device_select"] = True
    O.set_round(None)
    O.set_fused_sdpa(True)
    saved = (O.dit_block, O.mmdit_block, O.single_stream_block)
    try:
        for variant in (["eager"] if mode == "eager" else ["eager", "compile"]):
            if variant == "compile":
                def sub(fn):
                    comp = torch.compile(fn, dynamic=False)

                    def wrapped(sd, p, *a, **k):  # same dict keys for every block -> ONE compiled graph
                        return comp({"b" + key[len(p):]: v for key, v in sd.items() if key.startswith(p + ".")}, "b", *a, **k)
                    return wrapped
                O.dit_block, O.mmdit_block, O.single_stream_block = sub(saved[0]), sub(saved[1]), sub(saved[2])
            b_use = B
            while True:
                try:
                    sd, rsd, opt = oracle_setup(wl, device)
                    b = wl.batch(b_use, g)
                    for _ in range(2):
                        oracle_train_step(wl, ocfg, sd, rsd, opt, b, wl.p_cfg, device, True)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        loss = oracle_train_step(wl, ocfg, sd, rsd, opt, b, wl.p_cfg, device, True)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    out[f"ref_gpu_{variant}_img_per_s"] = round(b_use / (ms / 1e3), 2)
                    out[f"ref_gpu_{variant}_batch"] = b_use
                    out[f"ref_gpu_{variant}_loss"] = round(float(loss), 4)
                    break
                except torch.OutOfMemoryError:
                    b_use //= 2
                    if b_use < 1:
                        out[f"ref_gpu_{variant}_img_per_s"] = None
                        break
                finally:
                    sd = rsd = opt = None
                    torch.cuda.empty_cache()
            O.dit_block, O.mmdit_block, O.single_stream_block = saved
    except Exception as e:  # a failing comparison arm must not take the measurement down
        out["ref_gpu_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    finally:
        O.dit_block, O.mmdit_block, O.single_stream_block = saved
        O.set_fused_sdpa(False)
    out["ref_gpu_what"] = ("reference algorithm (oracle port) on this GPU: bf16 autocast, F.scaled_dot_product_attention, torch.optim.AdamW; "
                           "full train step, CUDA events")
    return out


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import diffulab_b200 as dl
    from diffulab_b200 import _lib, ops
    from diffulab_b200.synthetic import Workload, build_workload
    from diffulab_b200.training import EMA, DevicePrefetcher, FusedAdamW, GradReducer, LossReader, make_reduce_group, training_step

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

    overrides = ([f"dataloader.batch_size={args.batch}"] if args.batch else []) + (args.override or [])
    wl = build_workload(args.config, overrides, device=device, seed=1234 + rank, live_gates=not args.zero_init)
    cfg, model, repa = wl.cfg, wl.model.train(), wl.repa
    B = int(cfg["dataloader"]["batch_size"])  # per-GPU batch (weak scaling: global = B * world)
    extra_mods = [repa] if repa is not None else []
    if world > 1:  # identical initial weights on every rank (DDP broadcasts rank 0's; same effect)
        for t in list(model.parameters()) + [q for m in extra_mods for q in m.parameters()]:
            dist.broadcast(t.data, 0)
    d = cfg["diffuser"]
    diffuser = dl.Diffuser(model, sampling_method=d["sampling_method"], model_type=d["model_type"], n_steps=d["n_steps"],
                           extra_args=d.get("extra_args", {}), extra_losses=extra_mods)
    from diffulab_b200.config import instantiate

    opt = instantiate(cfg["optimizer"], params=list(model.parameters()) + [q for m in extra_mods for q in m.parameters()])
    assert isinstance(opt, FusedAdamW)
    ema = None
    if not args.no_ema:  # configs/trainer/default.yaml: use_ema true, ema_rate 0.999, update every 10 steps (base_trainer.py:152-153)
        tr = cfg["trainer"]
        ema = EMA(opt, beta=float(tr.get("ema_rate", 0.999)), update_after_step=int(tr.get("ema_update_after_step", 0)),
                  update_every=int(tr.get("ema_update_every", 10)))
    reducer = None
    if world > 1:
        def make_reducer(mode: str):
            return GradReducer(stores=opt.stores, bucket_mb=args.bucket_mb or GradReducer.DEFAULT_BUCKET_MB, tail_bucket_mb=args.tail_bucket_mb or None,
                               process_group=make_reduce_group(args.comm_ctas) if mode == "nccl" and args.comm_ctas > 0 else None,
                               reserve_sms=args.reserve_sms, mode=mode, comm_ctas=max(1, args.comm_ctas), use_graphs=not args.no_dp_graphs)

        mode, why = args.dp_mode, None
        if mode != "nccl":  # symmetric (peer-mapped) memory must come up on EVERY rank, else all ranks use NCCL
            ok = torch.ones(1, device=device)
            try:
                reducer = make_reducer(mode)
            except Exception as e:  # noqa: BLE001
                ok.zero_()
                why = repr(e)[:200]
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                print(f"[bench] rank {rank}: dp mode {mode!r} unavailable ({why or 'failed on another rank'}); using NCCL all-reduce", file=sys.stderr, flush=True)
                mode, reducer = "nccl", None
        if reducer is None:
            reducer = make_reducer("nccl")
    n_params = sum(p.numel() for p in model.parameters())

    pool = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host = [Workload.pin(wl.batch(B, g)) for _ in range(pool)]
    dev = [Workload.to_step(b, device) for b in host]

    def fresh(step_batch: dict) -> dict:  # training_step mutates model_inputs (p, x): hand it shallow copies
        return {"model_inputs": dict(step_batch["model_inputs"]), "extra": step_batch["extra"]}

    def step_resident(i: int):
        return training_step(diffuser, opt, fresh(dev[i % pool]), wl.p_cfg, reducer, ema=ema)

    def run_e2e(n: int) -> float:
        """n steps through the public API with HOST batches: every step's inputs are copied from pinned host memory (the copy of
        batch i + 1 runs on a side stream while step i computes) and every step's losses are read back to the host (asynchronously:
        the value of step i is collected while step i + 1 runs, the last one after the loop)."""
        reader, total = LossReader(), 0.0
        restructure = lambda b: Workload.to_step(b, "cpu")  # noqa: E731  (dict reshuffle only: the tensors are already on the host)
        for batch in DevicePrefetcher((host[i % pool] for i in range(n)), device, restructure=restructure):
            got = reader.push(training_step(diffuser, opt, batch, wl.p_cfg, reducer, ema=ema))
            total += sum(got.values()) if got else 0.0
        total += sum(reader.flush().values())
        return total

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

    # ---- timed region 1: inputs resident in HBM (nothing but the steps in here) -----------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        last = step_resident(i)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    value = B * world * args.steps / (ms / 1e3)
    final_losses = {k: float(v.item()) for k, v in last.items()}

    # ---- timed region 2: end to end through the public API with host batches -------------------------------
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loss_sum = run_e2e(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert e2e_loss_sum == e2e_loss_sum and e2e_loss_sum > 0, "end-to-end leg produced no finite losses"
    e2e_value = B * world * args.steps / e2e_s
    h2d = Workload.nbytes(host[0])

    # ---- data parallel: replicas must hold identical parameters after the steps; per-bucket timeline of one step ------
    dp = None
    if world > 1:
        sums = torch.stack([st.flat_p.double().sum() for st in opt.stores] + [st.flat_p.double().abs().sum() for st in opt.stores])
        gathered = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)
        identical = all(bool(torch.equal(gathered[0], gi)) for gi in gathered)
        assert identical, f"data-parallel replicas diverged: per-rank parameter checksums {[gi.tolist() for gi in gathered]}"
        reducer.timeline = []
        step_resident(0)
        torch.cuda.synchronize()
        tl = reducer.read_timeline()
        reducer.timeline = None
        dp = {"mode": reducer.mode, "replicas_identical_after_steps": identical, "buckets": len(reducer.buckets), "bucket_mb": args.bucket_mb or GradReducer.DEFAULT_BUCKET_MB,
              "tail_bucket_mb": args.tail_bucket_mb or None, "comm_ctas": args.comm_ctas, "reserve_sms": args.reserve_sms, "bucket_graphs": len(getattr(reducer, "_graphs", {})),
              "grad_bytes_per_step": int(sum(b.numel() for b in reducer.buckets) * 4), "backward_after_first_bucket_ms": tl["end_backward_ms"],
              "exposed_tail_ms": tl["exposed_tail_ms"]}
        if rank == 0:
            os.makedirs("gpurun_out", exist_ok=True)
            with open(f"gpurun_out/dp_timeline_{world}gpu.json", "w") as f:
                json.dump({**dp, "timeline": tl}, f, indent=1)

    # ---- separate profiled pass: per-kernel CUDA-event times -> roofline of the dominant kernel -------------
    psteps = max(1, min(args.profile_steps, args.steps))
    ops.profile_start()
    for i in range(psteps):
        step_resident(i)
    prof = ops.profile_stop()
    peaks, peak_src = load_peaks()
    g_items = {k: v for k, v in prof.items() if k.startswith("gemm_")}
    g_ms = sum(v["ms"] for v in g_items.values())
    g_fl = sum(v["work"] for v in g_items.values())
    g_calls = sum(v["calls"] for v in g_items.values())
    all_ms = sum(v["ms"] for v in prof.values())
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    peak = float(peaks["bf16_tflops_sustained"])
    traffic = None
    if os.path.exists(NCU_GEMM_TRAFFIC_FILE):
        with open(NCU_GEMM_TRAFFIC_FILE) as f:
            traffic = json.load(f)
    train_gf = wl.flops_per_image(True) / 1e9
    roofline = {"bound": "tensor", "kernel": "gemm2_tcgen05_kernel / gemm_tcgen05_kernel (fwd + dgrad + wgrad launches)", "achieved": round(achieved, 1),
                "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "traffic": traffic.get("mean_bytes_per_launch") if traffic else None,
                "traffic_by_shape": traffic.get("by_shape") if traffic else None,
                "traffic_source": "dram__bytes_read+write per launch per GEMM shape vs algorithmic bytes, profiles/ncu_gemm_traffic_r2.json (ncu --set full)",
                "peak_source": f"bf16_tflops_sustained, {peak_src}",
                "launches_per_step": int(g_calls / psteps), "gemm_ms_per_step": round(g_ms / psteps, 3),
                "gemm_share_of_kernel_time": round(g_ms / all_ms, 4) if all_ms else None,
                "profiled_steps": psteps,
                "step_model_flops_frac": round(value / world * train_gf * 1e9 / (peak * 1e12), 4)}

    def fam(k: str) -> str:
        return k.split(" ")[0]

    families: dict = {}
    for k, v in prof.items():
        f = families.setdefault(fam(k), {"ms_per_step": 0.0, "calls_per_step": 0.0})
        f["ms_per_step"] += v["ms"] / psteps
        f["calls_per_step"] += v["calls"] / psteps
    roofline["ms_per_step_by_family"] = {k: round(v["ms_per_step"], 3) for k, v in sorted(families.items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]}
    breakdown = {k: {"calls_per_step": v["calls"] / psteps, "ms_per_step": round(v["ms"] / psteps, 3),
                     **({"tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1)} if v["work"] > 0 and v["ms"] > 0 else {})}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- loss of the CUDA path vs the CPU oracle on the SAME (current) weights and inputs ---------------------
    loss_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import dit_oracle as O

        gg = torch.Generator().manual_seed(4242)
        b = wl.batch(2, gg)
        t = torch.rand(2, generator=gg) * 0.9 + 0.05
        eps = torch.randn(2, *wl.shape, generator=gg)
        step = Workload.to_step(b, device)
        flow = diffuser.diffusion
        sdc = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        ocfg = oracle_cfg_for(wl)
        kind = cfg["model"]["_target_"].rsplit(".", 1)[-1]
        model.eval()  # (only decides the SPRINT token drop: checked in eval mode = all tokens kept on both sides)
        with torch.no_grad():
            got_eval = flow.compute_loss(model, dict(step["model_inputs"], p=0.0), t.to(device), noise=eps.to(device))["loss"].item()
            x_t = O.flow_add_noise(b["x"], t, eps)
            common = dict(y=b.get("y"), context=b.get("context"), p=0.0, draws={"context": torch.ones(2)})
            O.set_round(None)
            if kind == "SprintDiT":
                pred = O.sprint_forward(sdc, ocfg, x_t, t, training=False, **common)
            elif kind == "DDT":
                pred = O.ddt_forward(sdc, ocfg, x_t, t, **common)
            else:
                pred = O.mmdit_forward(sdc, ocfg, x_t, t, **common)
            ref = float(O.flow_loss(pred, b["x"], eps))
        model.train()
        rel = abs(got_eval - ref) / abs(ref)
        loss_check = {"gpu": round(got_eval, 6), "cpu_oracle_fp32": round(ref, 6), "rel_diff": round(rel, 6), "tolerance": 1e-2, "batch": 2,
                      "weights": "after the timed steps"}
        assert rel <= 1e-2, f"loss of the CUDA path differs from the CPU oracle on the same weights: {loss_check}"

    # ---- 50-step Euler sampling sweep (sharded batch, no collective; loop replayed as one CUDA graph) --------
    sample = None
    if not args.no_sample:
        model.eval()
        name = args.config if args.config in METRICS else "custom"
        sweep = ([1, 8, 64, 256] if name == "sprint" else [args.sample_batch]) if not args.sample_batches else [int(v) for v in args.sample_batches.split(",")]
        shift = cfg["trainer"].get("val_step_shift")
        diffuser.set_steps(50, shift=shift) if shift else diffuser.set_steps(50)
        sample = {"steps": 50, "shift": shift, "cuda_graph": not args.no_graph}
        diffuser.diffusion.cuda_graph = not args.no_graph
        gs = torch.Generator().manual_seed(77 + rank)
        for sb in sweep:
            hb = wl.batch(sb, gs)
            inp = Workload.to_step(hb, device)["model_inputs"]
            for g_scale in (0.0, 4.0):
                if g_scale > 0 and not cfg["model"].get("classifier_free", False):
                    continue
                def once():
                    mi = dict(inp)
                    mi["x"] = torch.randn(sb, *wl.shape, device=device)
                    return diffuser.generate(mi, use_tqdm=False, guidance_scale=g_scale)["x"]
                once()  # warm-up (captures the graph)
                barrier()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                once()
                s1.record()
                barrier()
                sms = max_over_ranks(s0.elapsed_time(s1))
                sample[f"euler50_b{sb}_cfg{g_scale:g}_img_per_s"] = round(sb * world / (sms / 1e3), 2)
        diffuser.diffusion.cuda_graph = False
        model.train()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sdm = {k: v for k, v in model.state_dict().items()}
        cpu = run_cpu_arm(args, wl, as_reference=False, state_dict=sdm, repa_sd=repa.state_dict() if repa is not None else None)
        if args.ref_gpu != "off":
            del opt, diffuser
            torch.cuda.empty_cache()
            cpu.update(run_ref_gpu_arm(wl, device, min(B, args.ref_gpu_batch), args.ref_gpu))

    if rank == 0:
        name = args.config if args.config in METRICS else "custom"
        e2e = {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 * len(final_losses),
               "ms_per_step": round(1e3 * e2e_s / args.steps, 3)}
        if sample:
            e2e.update({f"sample_{k}": v for k, v in sample.items()})
        line = {
            "metric": METRICS.get(name, f"{name} train throughput"), "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": {**config_block(name, wl, B, world, n_params), **({"dp": dp} if dp else {})},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "losses": final_losses,
            "loss_check": loss_check, "sample": sample, "kernel_breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from diffulab_b200.synthetic import build_workload

    overrides = ([f"dataloader.batch_size={args.batch}"] if args.batch else []) + (args.override or [])
    wl = build_workload(args.config, overrides, device=None, seed=1234, live_gates=not args.zero_init)
    res = run_cpu_arm(args, wl, as_reference=True)
    name = args.config if args.config in METRICS else "custom"
    B = int(wl.cfg["dataloader"]["batch_size"])
    n_params = sum(p.numel() for p in wl.model.parameters())
    line = {"impl": "reference", "metric": METRICS.get(name, f"{name} train throughput"), "value": round(res["value"], 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(res["ms_per_step"], 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(name, wl, B, args.gpus, n_params),  # the SAME workload object as our arm; the host sample is below
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": round(res["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="imagenet_repa", help="imagenet_repa (metric config) | cifar10 | txt_to_img | sprint | path to a YAML")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's dataloader.batch_size)")
    ap.add_argument("--bucket-mb", type=float, default=0.0, help="gradient bucket size (default: GradReducer.DEFAULT_BUCKET_MB)")
    ap.add_argument("--tail-bucket-mb", type=float, default=32.0, help="size cap of the buckets backward produces last (0 = same as --bucket-mb)")
    ap.add_argument("--dp-mode", default="ce", choices=["nccl", "ce", "nvls"], help="gradient reduction: NCCL all-reduce | copy-engine pulls over peer "
                    "memory + reduce kernel | in-switch multimem reduction kernel")
    ap.add_argument("--no-dp-graphs", action="store_true", help="launch the per-bucket peer-memory reduction eagerly instead of replaying one CUDA graph per bucket")
    ap.add_argument("--comm-ctas", type=int, default=0, help="CTAs per gradient reduction (nccl: dedicated communicator, 0 = NCCL's default; nvls: kernel grid)")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs the persistent kernels leave to the collective while buckets are in flight")
    ap.add_argument("--sample-batch", type=int, default=64)
    ap.add_argument("--sample-batches", default="", help="comma-separated batch sweep for the sampling benchmark")
    ap.add_argument("--no-sample", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the sampling loop eagerly instead of as one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-gpu", default="eager", choices=["off", "eager", "compile"], help="reference-GPU comparison arm (N=1 only)")
    ap.add_argument("--ref-gpu-batch", type=int, default=64)
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--no-ema", action="store_true", help="do not attach the fused EMA (the reference trainer's default has use_ema: true)")
    ap.add_argument("--zero-init", action="store_true", help="keep the adaLN-Zero initialisation (gates exactly 0) instead of live gates")
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
