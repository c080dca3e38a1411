#!/usr/bin/env python
"""One DiT-XL/2 (+REPA) train step at the bench geometry (per-GPU batch 128, N 256, d 1152) with a reduced DEPTH, for ncu:
every kernel of the real step appears with its real shape, the capture stays short. The step to profile is wrapped in the NVTX
step is bracketed by cudaProfilerStart/Stop (two warm-up steps run before it; NVTX ranges are per thread and would miss the
backward kernels, which the autograd engine launches from its own thread):

    ncu --profile-from-start off --metrics <...> --csv --log-file gpurun_out/step_light.csv python scripts/profile_step.py
    ncu --profile-from-start off --set full --import-source on -k regex:'attn|ln_mod|qknorm|gate_res' ...

--config picks another BASELINE config (cifar10 | txt_to_img | sprint); --depth overrides the block count where the config has one.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="imagenet_repa")
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    import torch

    import diffulab_b200 as dl
    from diffulab_b200.config import instantiate
    from diffulab_b200.synthetic import Workload, build_workload
    from diffulab_b200.training import FusedAdamW, training_step

    ov = [f"dataloader.batch_size={args.batch}"]
    if args.config == "imagenet_repa":
        ov += [f"model.depth={args.depth}", f"repa.alignment_layer={min(args.depth, 8)}"]
    wl = build_workload(args.config, ov, device="cuda")
    model, repa = wl.model.train(), wl.repa
    d = wl.cfg["diffuser"]
    extra = [repa] if repa is not None else []
    diffuser = dl.Diffuser(model, sampling_method=d["sampling_method"], model_type=d["model_type"], n_steps=d["n_steps"],
                           extra_args=d.get("extra_args", {}), extra_losses=extra)
    opt = instantiate(wl.cfg["optimizer"], params=list(model.parameters()) + [q for m in extra for q in m.parameters()])
    assert isinstance(opt, FusedAdamW)
    g = torch.Generator().manual_seed(0)
    batches = [Workload.to_step(wl.batch(args.batch, g), "cuda") for _ in range(2)]

    def step(i):
        b = batches[i % 2]
        return training_step(diffuser, opt, {"model_inputs": dict(b["model_inputs"]), "extra": b["extra"]}, wl.p_cfg)

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(args.steps):
        out = step(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print({k: float(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
