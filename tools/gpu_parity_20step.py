#!/usr/bin/env python
"""North-star parity measurement (BASELINE.json: "denoised latents within 1e-3 max-abs of reference on fixed seed";
SURVEY.md 7.2: report (ours - ref32) next to (ref16 - ref32)).

Runs the N-step parallel-denoise loop at a given shape on one GPU, same seed and same (fp16-rounded) weights, through
  ours   : musev_b200 engine (ParallelDenoiser -> C ABI -> CUDA kernels), fp32 latents
  ref32  : the oracle in fp32 on the GPU (TF32 off) -- ground truth
  ref16  : the oracle as eager PyTorch fp16 on the GPU, fp16 latents -- the way the reference itself runs
           (scripts/inference/text2video.py:590 torch_dtype = float16)
  emu_all / emu_branch (optional): the fp32 oracle with fp16 rounding inserted where the engine stores fp16
           ("all") or everywhere except the residual stream ("branch" = what an fp32-residual-stream engine would keep)
and prints one JSON line with max-abs / rms distances of the final latents and of the first step's eps.

The oracle is test infrastructure (oracle/); this tool and tests/ are the only users.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(preset="musev", steps=20, T=16, h=64, w=64, boc=(320, 640, 1280, 1280), window=16, overlap=4, guidance=3.5,
        emulate=(), seed=1234, dev="cuda"):
    from musev_b200.pipeline import ParallelDenoiser
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    from musev_b200.schema import preset_config
    from musev_b200.synth import make_inputs, make_state_dict
    from musev_b200.unet import UNet3DConditionModel
    from oracle.pipeline_oracle import SD15_DDIM, DDIMOracle, denoise_loop
    from oracle.unet3d_oracle import UNet3DOracle

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = preset_config(preset, block_out_channels=tuple(boc))
    sd16 = make_state_dict(cfg, seed=0, dtype=torch.float16)
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(1, 4, T, h, w, generator=g)
    cond = torch.randn(1, 4, 1, h, w, generator=g) * 0.18215
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    extra = make_inputs(cfg, batch=2, frames=1, h=h, w=w, seed=seed)
    kw = {k: extra[k] for k in ("down_block_refer_embs", "mid_block_refer_emb", "vision_clip_emb") if k in extra}
    if kw:
        kw["ip_adapter_scale"] = 1.0

    def to_dev(v, dt):
        if torch.is_tensor(v):
            return v.to(dev, dt) if v.is_floating_point() else v
        if isinstance(v, (list, tuple)):
            return [to_dev(x, dt) for x in v]
        return v

    res = {"preset": preset, "steps": steps, "T": T, "h": h, "w": w, "block_out_channels": list(boc), "window": window,
           "overlap": overlap, "guidance_scale": guidance, "seed": seed}

    # ---- ours
    model = UNet3DConditionModel(cfg, device=dev, dtype=torch.float32)
    model.load_state_dict(sd16)
    den = ParallelDenoiser(model, DDIMScheduler(**SD15_DDIM_CONFIG))
    trace_ours = []
    t0 = time.time()
    out = den(latents.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=steps, guidance_scale=guidance,
              context_frames=window, context_overlap=overlap, unet_kwargs={k: to_dev(v, torch.float32) for k, v in kw.items()},
              callback=lambda i, t, l: trace_ours.append(l.float().cpu())).latents.float().cpu()
    torch.cuda.synchronize()
    res["ours_seconds"] = time.time() - t0
    del model, den
    torch.cuda.empty_cache()

    def oracle_loop(dtype, emu=None):
        o = UNet3DOracle(cfg, sd16, device=dev, dtype=dtype, emulate=emu)
        trace = []

        def unet(s, t, e, **k):
            r = o(s, t, e, **k).to("cpu", dtype)
            return r
        lat0 = latents.to(dtype)
        final, eps_trace = denoise_loop(unet, DDIMOracle(**SD15_DDIM), lat0, cond.to(dtype), prompt.to(dtype), steps, guidance,
                                        context_frames=window, context_overlap=overlap,
                                        unet_kwargs={k: to_dev(v, dtype) for k, v in kw.items()}, return_eps=True)
        del o
        torch.cuda.empty_cache()
        return final.float(), [e.float() for e in eps_trace]

    t0 = time.time()
    ref32, eps32 = oracle_loop(torch.float32)
    res["ref32_seconds"] = time.time() - t0
    ref16, eps16 = oracle_loop(torch.float16)

    def dist(a, b):
        d = (a - b).abs()
        return {"max_abs": d.max().item(), "rms": d.pow(2).mean().sqrt().item()}

    res["latents_std"] = ref32.std().item()
    res["ours_minus_ref32"] = dist(out, ref32)
    res["ref16_minus_ref32"] = dist(ref16, ref32)
    res["ours_minus_ref16"] = dist(out, ref16)
    # first-step eps (one forward + CFG): the per-forward distance without loop amplification
    a_t = None
    for name in emulate:
        e_lat, e_eps = oracle_loop(torch.float32, emu=name)
        res[f"emu_{name}_minus_ref32"] = dist(e_lat, ref32)
        res[f"emu_{name}_eps0_minus_ref32"] = dist(e_eps[0], eps32[0])
        res[f"ours_minus_emu_{name}"] = dist(out, e_lat)
    res["ref16_eps0_minus_ref32"] = dist(eps16[0], eps32[0])
    # per-step growth of the latent distance (ours vs ref32 needs the ref32 latent trace: recompute from eps trace)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="musev")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--hw", type=int, nargs=2, default=[64, 64])
    ap.add_argument("--narrow", action="store_true")
    ap.add_argument("--emulate", nargs="*", default=[])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    boc = (64, 128, 128, 128) if a.narrow else (320, 640, 1280, 1280)
    r = run(a.preset, a.steps, a.frames, a.hw[0], a.hw[1], boc, emulate=a.emulate)
    line = json.dumps(r)
    print("PARITY20 " + line, flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as fh:
            fh.write(line + "\n")


if __name__ == "__main__":
    main()
