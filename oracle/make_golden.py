"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python -m oracle.make_golden [--full]
The fixtures hold seeds + reference outputs; weights and inputs are regenerated from the seeds by
musev_b200.synth (bit-identical CPU RNG), so the files stay small. /root/reference is never read at test time.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from musev_b200.schema import (ControlNetConfig, ReferenceNetConfig, VAEConfig, controlnet_param_shapes,  # noqa: E402
                               preset_config, referencenet_param_shapes, unet_param_shapes, vae_decoder_param_shapes)
from musev_b200.synth import make_controlnet_inputs, make_inputs, make_referencenet_inputs, make_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.pipeline_oracle import SD15_DDIM  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NARROW = (64, 128, 128, 128)
FULL = (320, 640, 1280, 1280)


def build_reference(preset: str, boc, sd):
    U, _ = ref_shim.load()
    kw = dict(ref_shim.SD15_KW)
    kw.update(ref_shim.PRESET_KW[preset])
    kw["block_out_channels"] = boc
    cfg = preset_config(preset, block_out_channels=boc)
    with torch.device("meta"):
        m = U(**kw)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine = {k: tuple(v) for k, v in unet_param_shapes(cfg).items()}
    assert ref_shapes == mine, "schema mismatch vs reference state_dict"
    m = m.to_empty(device="cpu")
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval(), cfg


def run_reference_unet(m, inp, t, frame_rate, ip_scale):
    with torch.no_grad():
        return m(inp["sample"], torch.tensor(t), inp["encoder_hidden_states"], sample_index=inp["sample_index"],
                 vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"],
                 sample_frame_rate=frame_rate, do_classifier_free_guidance=True,
                 down_block_refer_embs=inp.get("down_block_refer_embs"),
                 mid_block_refer_emb=inp.get("mid_block_refer_emb"), vision_clip_emb=inp.get("vision_clip_emb"),
                 ip_adapter_scale=ip_scale)[0]


def golden_unet(preset, boc, tag, batch, frames, h, w, t, wseed=0, iseed=1234, frame_rate=8, ip_scale=0.7):
    cfg = preset_config(preset, block_out_channels=boc)
    t0 = time.time()
    sd = make_state_dict(cfg, seed=wseed)
    m, cfg = build_reference(preset, boc, sd)
    inp = make_inputs(cfg, batch=batch, frames=frames, h=h, w=w, n_vis_cond=1, seed=iseed)
    out = run_reference_unet(m, inp, t, frame_rate, ip_scale)
    meta = dict(preset=preset, block_out_channels=list(boc), batch=batch, frames=frames, h=h, w=w, timestep=t,
                weight_seed=wseed, input_seed=iseed, sample_frame_rate=frame_rate, ip_adapter_scale=ip_scale,
                n_vis_cond=1, source="reference musev.models.unet_3d_condition.UNet3DConditionModel, CPU fp32")
    path = os.path.join(GOLDEN, f"unet_{preset}_{tag}.pt")
    torch.save({"meta": meta, "out": out.clone()}, path)
    print(f"{path}: out std {out.std().item():.4f} ({time.time() - t0:.1f}s)", flush=True)
    return m, cfg, sd


def golden_loop(preset, boc, tag, m, cfg, T=20, h=8, w=8, steps=2, wseed=0, iseed=77):
    """Restates the window loop (pipeline_controlnet.py:1846-2117) around the IMPORTED UNet and the IMPORTED
    musev DDIMScheduler + prepare_global_context."""
    _, DDIM = ref_shim.load()
    spec = importlib.util.spec_from_file_location(
        "mmcm.utils.itertools_util", os.path.join(ref_shim.REFERENCE_ROOT, "MMCM/mmcm/utils/itertools_util.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["mmcm.utils.itertools_util"] = mod
    from musev.pipelines.context import prepare_global_context

    sched = DDIM(**SD15_DDIM)
    sched.set_timesteps(steps)
    g = torch.Generator().manual_seed(iseed)
    latents = torch.randn(1, 4, T, h, w, generator=g)
    cond = torch.randn(1, 4, 1, h, w, generator=g) * 0.5
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    extra = make_inputs(cfg, batch=2, frames=1, h=h, w=w, seed=iseed)
    kw = {k: extra[k] for k in ("down_block_refer_embs", "mid_block_refer_emb", "vision_clip_emb") if k in extra}
    ctx = prepare_global_context("uniform_v2", steps, T, 8, 1, 2, 1)
    gs = 3.5
    vis_idx = torch.arange(1)
    with torch.no_grad():
        for t in sched.timesteps:
            noise_pred = torch.zeros(2, 4, T, h, w)
            counter = torch.zeros(1, 1, T, 1, 1)
            for context in ctx:
                c = context[0]
                lat = torch.cat([latents[:, :, c]] * 2)
                sub = torch.arange(len(c)) + 1
                full = torch.zeros(2, 4, 1 + len(c), h, w)
                full[:, :, vis_idx] = torch.cat([cond] * 2)
                full[:, :, sub] = lat
                eps = m(full, t, prompt, sample_index=sub, vision_conditon_frames_sample_index=vis_idx,
                        sample_frame_rate=8, do_classifier_free_guidance=True, ip_adapter_scale=1.0, **kw)[0]
                noise_pred[:, :, c] += eps[:, :, sub]
                counter[:, :, c] += 1
            noise_pred = noise_pred / counter
            u, tx = noise_pred.chunk(2)
            noise_pred = u + gs * (tx - u)
            latents_next = sched.step(noise_pred, t, latents, eta=0.0).prev_sample
            latents = latents_next
    meta = dict(preset=preset, block_out_channels=list(boc), T=T, h=h, w=w, steps=steps, weight_seed=wseed,
                input_seed=iseed, context_frames=8, context_overlap=2, guidance_scale=gs, contexts=[c[0] for c in ctx],
                source="imported reference UNet + musev DDIMScheduler + prepare_global_context; loop restated")
    path = os.path.join(GOLDEN, f"loop_{preset}_{tag}.pt")
    torch.save({"meta": meta, "latents": latents.clone()}, path)
    print(path, "final latents std", latents.std().item(), flush=True)


def golden_contexts():
    spec = importlib.util.spec_from_file_location(
        "mmcm.utils.itertools_util", os.path.join(ref_shim.REFERENCE_ROOT, "MMCM/mmcm/utils/itertools_util.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["mmcm.utils.itertools_util"] = mod
    ref_shim.load()
    from musev.pipelines.context import prepare_global_context

    cases = []
    for sched in ("uniform", "uniform_v2"):
        for (T, win, ov) in [(16, 16, 4), (48, 16, 4), (48, 16, 8), (128, 16, 4), (512, 16, 8), (12, 12, 4),
                             (20, 8, 2), (7, 12, 4), (33, 12, 4)]:
            ctx = prepare_global_context(sched, 20, T, win, 1, ov, 1)
            cases.append(dict(schedule=sched, T=T, window=win, overlap=ov,
                              contexts=[[int(i) for i in c[0]] for c in ctx]))
    with open(os.path.join(GOLDEN, "contexts.json"), "w") as fh:
        json.dump(cases, fh)
    print("contexts.json", len(cases), "cases")


def golden_ddim():
    _, DDIM = ref_shim.load()
    sched = DDIM(**SD15_DDIM)
    sched.set_timesteps(20)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 3, 8, 8, generator=g)
    eps = torch.randn(1, 4, 3, 8, 8, generator=g)
    outs = {}
    for t in (951, 501, 1):
        outs[str(t)] = sched.step(eps, t, x, eta=0.0).prev_sample.clone()
    torch.save({"timesteps": sched.timesteps.clone(), "x": x, "eps": eps, "prev": outs,
                "source": "musev.schedulers.DDIMScheduler (imported reference), SD-1.5 config, 20 steps"},
               os.path.join(GOLDEN, "ddim_sd15.pt"))
    print("ddim_sd15.pt timesteps", sched.timesteps.tolist())


def golden_vae(boc, tag, frames, h, w, wseed=11, iseed=1357):
    """The unmodified diffusers `AutoencoderKL.decode` (vendored fork) + the pipeline's decode_latents post-processing
    (pipeline_stable_diffusion_img2img.py:490-492) on seeded decoder weights; 2048 seeded sample positions of the image."""
    ref_shim.load()
    from diffusers.models.autoencoder_kl import AutoencoderKL
    cfg = VAEConfig(block_out_channels=tuple(boc))
    kw = dict(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
              block_out_channels=tuple(boc), layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=32,
              sample_size=512, scaling_factor=0.18215)
    t0 = time.time()
    m = AutoencoderKL(**kw).eval()
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    mine = {k: tuple(v) for k, v in vae_decoder_param_shapes(cfg).items()}
    assert ref_shapes == mine, "VAE decoder schema mismatch vs reference state_dict"
    sd = make_state_dict(cfg, seed=wseed)
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith(("encoder.", "quant_conv.")) for k in res.missing_keys)
    g = torch.Generator().manual_seed(iseed)
    latents = torch.randn(1, 4, frames, h, w, generator=g) * 0.18215 * 1.2
    with torch.no_grad():
        z = latents.permute(0, 2, 1, 3, 4).reshape(frames, 4, h, w) / 0.18215
        raw = m.decode(z, return_dict=False)[0]
        img = (raw / 2 + 0.5).clamp(0, 1)
    flat_raw, flat_img = raw.reshape(-1), img.reshape(-1)
    idx = torch.randint(0, flat_raw.numel(), (2048,), generator=torch.Generator().manual_seed(3000))
    meta = dict(block_out_channels=list(boc), frames=frames, h=h, w=w, weight_seed=wseed, input_seed=iseed,
                shape=list(raw.shape), sample_seed=3000, n_samples=2048,
                source="reference diffusers.models.autoencoder_kl.AutoencoderKL.decode (vendored fork), CPU fp32")
    path = os.path.join(GOLDEN, f"vae_{tag}.pt")
    torch.save({"meta": meta, "raw": flat_raw[idx].clone(), "img": flat_img[idx].clone(),
                "stats": [float(raw.mean()), float(raw.abs().mean()), float(img.mean())]}, path)
    print(f"{path}: raw abs-mean {float(raw.abs().mean()):.4f} img mean {float(img.mean()):.4f} ({time.time() - t0:.1f}s)", flush=True)


def golden_samplers():
    """The imported musev Euler / LCM schedulers (musev/schedulers/scheduling_{euler_discrete,lcm}.py) on the SD-1.5
    scheduler config: timesteps, sigmas, init_noise_sigma and a full deterministic loop (model = sample * t / (t + 1))."""
    ref_shim.load()
    from musev.schedulers import EulerDiscreteScheduler, LCMScheduler
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(1, 4, 3, 8, 8, generator=g)
    out = {"x": x0}
    e = EulerDiscreteScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                               steps_offset=1, timestep_spacing="leading")
    e.set_timesteps(20)
    x = x0 * e.init_noise_sigma
    trace = []
    for t in e.timesteps:
        xs = e.scale_model_input(x, t)
        x = e.step(xs * t / (t + 1), t, x, generator=torch.Generator().manual_seed(1)).prev_sample
        trace.append(x.clone())
    out["euler"] = dict(timesteps=e.timesteps.clone(), sigmas=e.sigmas.clone(), init_noise_sigma=float(e.init_noise_sigma),
                        trace=trace)
    l = LCMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    l.set_timesteps(4)
    gen = torch.Generator().manual_seed(7)
    x = x0.clone()
    trace, den = [], []
    for t in l.timesteps:
        r = l.step(x * t / (t + 1), t, x, generator=gen)
        x = r.prev_sample
        trace.append(x.clone())
        den.append(r.denoised.clone())
    out["lcm"] = dict(timesteps=l.timesteps.clone(), trace=trace, denoised=den, noise_seed=7)
    out["source"] = "musev.schedulers.EulerDiscreteScheduler / LCMScheduler (imported reference), SD-1.5 betas"
    torch.save(out, os.path.join(GOLDEN, "samplers_sd15.pt"))
    print("samplers_sd15.pt euler timesteps", e.timesteps.tolist()[:4], "... lcm", l.timesteps.tolist())


def golden_controlnet(boc, tag, frames, h, w, t, scale, guess, wseed=3, iseed=4321):
    """The per-window-step ControlNet: the unmodified diffusers `ControlNetModel` of the reference
    tree, called the way `get_controlnet_emb` calls it (pipeline_controlnet.py:1238-1262). The 13 residual maps are
    large, so the fixture keeps 512 seeded sample positions + mean / abs-mean of each."""
    ref_shim.load()
    from diffusers.models.controlnet import ControlNetModel
    cfg = ControlNetConfig(block_out_channels=tuple(boc))
    kw = dict(in_channels=4, conditioning_channels=3, block_out_channels=tuple(boc), layers_per_block=2,
              cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32,
              down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"))
    t0 = time.time()
    with torch.device("meta"):
        m = ControlNetModel(**kw)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine = {k: tuple(v) for k, v in controlnet_param_shapes(cfg).items()}
    assert ref_shapes == mine, "ControlNet schema mismatch vs reference state_dict"
    sd = make_state_dict(cfg, seed=wseed)
    m = m.to_empty(device="cpu")
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.eval()
    inp = make_controlnet_inputs(cfg, frames=frames, h=h, w=w, seed=iseed)
    with torch.no_grad():
        down, mid = m(inp["sample"], torch.tensor(t), inp["encoder_hidden_states"], controlnet_cond=inp["controlnet_cond"],
                      conditioning_scale=scale, guess_mode=guess, return_dict=False)
        # second call exactly as the pipeline issues it: the condition embedding computed once, passed as latents
        cond_lat = m.controlnet_cond_embedding(inp["controlnet_cond"])
        down2, mid2 = m(inp["sample"], torch.tensor(t), inp["encoder_hidden_states"], controlnet_cond=None,
                        controlnet_cond_latents=cond_lat, conditioning_scale=scale, guess_mode=guess, return_dict=False)
    assert all(torch.equal(a, b) for a, b in zip(down, down2)) and torch.equal(mid, mid2)
    maps = list(down) + [mid]
    samples, stats = [], []
    for k, mp in enumerate(maps):
        flat = mp.reshape(-1)
        g = torch.Generator().manual_seed(1000 + k)
        idx = torch.randint(0, flat.numel(), (512,), generator=g)
        samples.append(flat[idx].clone())
        stats.append([float(flat.mean()), float(flat.abs().mean())])
    meta = dict(block_out_channels=list(boc), frames=frames, h=h, w=w, timestep=t, conditioning_scale=scale,
                guess_mode=guess, weight_seed=wseed, input_seed=iseed, shapes=[list(mp.shape) for mp in maps],
                sample_seed_base=1000, n_samples=512,
                source="reference diffusers.models.controlnet.ControlNetModel (vendored fork), CPU fp32")
    path = os.path.join(GOLDEN, f"controlnet_{tag}.pt")
    torch.save({"meta": meta, "samples": samples, "stats": stats}, path)
    print(f"{path}: mid abs-mean {stats[-1][1]:.4f} ({time.time() - t0:.1f}s)", flush=True)


def golden_referencenet(boc, tag, batch, n_ref, h, w, wseed=5, iseed=2468):
    """The one-shot ReferenceNet: the unmodified `musev.models.referencenet.ReferenceNet2D`, built and
    called the way the reference does (referencenet_loader.py:109-118; pipeline_controlnet.py:918-929: timestep 0,
    encoder_hidden_states = IP-Adapter image tokens, return_ndim 5). 512 seeded sample positions per map."""
    ref_shim.load()
    from musev.models.referencenet import ReferenceNet2D
    cfg = ReferenceNetConfig(block_out_channels=tuple(boc))
    kw = dict(sample_size=64, in_channels=4, out_channels=4, block_out_channels=tuple(boc), layers_per_block=2,
              cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32,
              down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
              up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
              need_self_attn_block_embs=False, need_block_embs=True)
    t0 = time.time()
    with torch.device("meta"):
        m = ReferenceNet2D(**kw)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine = {k: tuple(v) for k, v in referencenet_param_shapes(cfg).items()}
    assert ref_shapes == mine, "ReferenceNet schema mismatch vs reference state_dict"
    sd = make_state_dict(cfg, seed=wseed)
    m = m.to_empty(device="cpu")
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.eval()
    inp = make_referencenet_inputs(cfg, batch=batch, n_ref=n_ref, h=h, w=w, seed=iseed)
    with torch.no_grad():
        down, mid, self_attn = m(inp["sample"], torch.zeros((), dtype=torch.long), inp["encoder_hidden_states"],
                                 num_frames=n_ref, return_ndim=5)
    assert self_attn is None
    maps = list(down) + [mid]
    samples, stats = [], []
    for k, mp in enumerate(maps):
        flat = mp.reshape(-1)
        g = torch.Generator().manual_seed(2000 + k)
        idx = torch.randint(0, flat.numel(), (512,), generator=g)
        samples.append(flat[idx].clone())
        stats.append([float(flat.mean()), float(flat.abs().mean())])
    meta = dict(block_out_channels=list(boc), batch=batch, n_ref=n_ref, h=h, w=w, weight_seed=wseed, input_seed=iseed,
                shapes=[list(mp.shape) for mp in maps], sample_seed_base=2000, n_samples=512,
                source="reference musev.models.referencenet.ReferenceNet2D, CPU fp32")
    path = os.path.join(GOLDEN, f"referencenet_{tag}.pt")
    torch.save({"meta": meta, "samples": samples, "stats": stats}, path)
    print(f"{path}: mid abs-mean {stats[-1][1]:.4f} ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also produce the full-width (1.4 B parameter) fixtures")
    ap.add_argument("--only", default="", help="'controlnet' / 'referencenet': regenerate only those fixtures")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if args.only == "vae":
        golden_vae((64, 64, 128, 128), "narrow", frames=2, h=8, w=8)
        if args.full:
            golden_vae((128, 256, 512, 512), "full", frames=1, h=8, w=8)
        sys.exit(0)
    if args.only == "samplers":
        golden_samplers()
        sys.exit(0)
    if args.only == "referencenet":
        golden_referencenet(NARROW, "narrow", batch=2, n_ref=1, h=16, w=16)
        golden_referencenet(NARROW, "narrow_t2", batch=1, n_ref=2, h=8, w=8)
        if args.full:
            golden_referencenet(FULL, "full", batch=2, n_ref=1, h=8, w=8)
        sys.exit(0)
    if args.only == "controlnet":
        golden_controlnet(NARROW, "narrow", frames=3, h=16, w=16, t=601, scale=0.8, guess=False)
        golden_controlnet(NARROW, "narrow_guess", frames=2, h=8, w=8, t=301, scale=1.0, guess=True)
        if args.full:
            golden_controlnet(FULL, "full", frames=2, h=8, w=8, t=601, scale=1.0, guess=False)
        sys.exit(0)
    golden_contexts()
    golden_ddim()
    golden_samplers()
    for preset in ("musev", "musev_referencenet"):
        m, cfg, sd = golden_unet(preset, NARROW, "narrow", batch=2, frames=4, h=16, w=16, t=601)
        golden_loop(preset, NARROW, "narrow", m, cfg)
        del m, sd
    golden_controlnet(NARROW, "narrow", frames=3, h=16, w=16, t=601, scale=0.8, guess=False)
    golden_controlnet(NARROW, "narrow_guess", frames=2, h=8, w=8, t=301, scale=1.0, guess=True)
    golden_referencenet(NARROW, "narrow", batch=2, n_ref=1, h=16, w=16)
    golden_referencenet(NARROW, "narrow_t2", batch=1, n_ref=2, h=8, w=8)
    golden_vae((64, 64, 128, 128), "narrow", frames=2, h=8, w=8)
    if args.full:
        golden_vae((128, 256, 512, 512), "full", frames=1, h=8, w=8)
        golden_referencenet(FULL, "full", batch=2, n_ref=1, h=8, w=8)
        for preset in ("musev", "musev_referencenet"):
            m, cfg, sd = golden_unet(preset, FULL, "full", batch=2, frames=2, h=8, w=8, t=601)
            del m, sd
        golden_controlnet(FULL, "full", frames=2, h=8, w=8, t=601, scale=1.0, guess=False)
