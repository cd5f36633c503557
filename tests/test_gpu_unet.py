"""GPU parity tests of the whole denoiser forward and of the parallel-denoise loop, through the public classes
(which call the C ABI), against (a) the oracle on the same fp16-rounded weights and (b) outputs of the UNMODIFIED
reference stored in tests/golden/ (fp32 weights).

Tolerance. Storage is fp16 with fp32 accumulation (the reference itself runs fp16, scripts/inference/text2video.py:590).
One forward on eps of std ~0.6 lands at 3e-3..8e-3 max-abs from the fp32 oracle; the fp16 rounding of the WEIGHTS alone
moves the fp32 oracle by ~2e-3..4e-3 from the fp32 reference. FWD_TOL = 1e-2 max-abs per forward is just above the largest
measured case; every measured distance is appended to gpurun_out/parity_measured.jsonl (`_record`). The 20-step latents
comparison on the north star's own terms is tests/test_gpu_parity20.py.
"""
import os

import pytest
import torch

from conftest import GOLDEN
from musev_b200.schema import preset_config
from musev_b200.synth import make_inputs, make_state_dict

pytestmark = pytest.mark.gpu
dev = "cuda"
FWD_TOL = 1e-2


def _record(name, value):
    """Measured distances go to gpurun_out/parity_measured.jsonl so that the written bounds can be audited against them."""
    import json
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/parity_measured.jsonl", "a") as fh:
            fh.write(json.dumps({"test": name, "value": value}) + "\n")
    except OSError:
        pass


def _setup(preset, boc, built_lib, io_dtype=torch.float32):
    from musev_b200.unet import UNet3DConditionModel
    from oracle.unet3d_oracle import UNet3DOracle
    cfg = preset_config(preset, block_out_channels=boc)
    sd16 = {k: v.half() for k, v in make_state_dict(cfg, seed=0).items()}
    model = UNet3DConditionModel(cfg, device=dev, dtype=io_dtype)
    model.load_state_dict({k: v.to(dev) for k, v in sd16.items()})
    oracle = UNet3DOracle(cfg, {k: v.float() for k, v in sd16.items()}, device=dev)
    return cfg, model, oracle


def _call_kwargs(inp, frame_rate=8, ip_scale=0.7):
    return dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"],
                sample_frame_rate=frame_rate, down_block_refer_embs=inp.get("down_block_refer_embs"),
                mid_block_refer_emb=inp.get("mid_block_refer_emb"), vision_clip_emb=inp.get("vision_clip_emb"),
                ip_adapter_scale=ip_scale)


def _to(v, d, dt):
    if torch.is_tensor(v) and v.is_floating_point():
        return v.to(d, dt)
    if isinstance(v, list):
        return [_to(x, d, dt) for x in v]
    return v


@pytest.mark.parametrize("preset", ["musev", "musev_referencenet"])
@pytest.mark.parametrize("tag,boc", [("narrow", (64, 128, 128, 128)), ("full", (320, 640, 1280, 1280))])
def test_forward_vs_oracle_and_reference_golden(built_lib, preset, tag, boc):
    g = torch.load(os.path.join(GOLDEN, f"unet_{preset}_{tag}.pt"))
    m = g["meta"]
    cfg, model, oracle = _setup(preset, boc, built_lib)
    inp = make_inputs(cfg, batch=m["batch"], frames=m["frames"], h=m["h"], w=m["w"], n_vis_cond=1, seed=m["input_seed"])
    kw = _call_kwargs(inp, m["sample_frame_rate"], m["ip_adapter_scale"])
    ref = oracle(inp["sample"], m["timestep"], inp["encoder_hidden_states"], **kw)
    out = model(_to(inp["sample"], dev, torch.float32), torch.tensor(m["timestep"]), _to(inp["encoder_hidden_states"], dev, torch.float32),
                do_classifier_free_guidance=True, **{k: _to(v, dev, torch.float32) for k, v in kw.items()}).sample
    assert out.shape == ref.shape and out.dtype == torch.float32 and not torch.isnan(out).any()
    _record(f"fwd_{preset}_{tag}_vs_oracle", (out - ref).abs().max().item())
    _record(f"fwd_{preset}_{tag}_vs_reference_golden", (out.cpu() - g["out"]).abs().max().item())
    assert (out - ref).abs().max().item() < FWD_TOL
    assert (out.cpu() - g["out"]).abs().max().item() < FWD_TOL            # reference itself (fp32 weights)
    # every layer output, not only the final one (catches errors that later layers would wash out)
    oracle.taps = {}
    oracle(inp["sample"], m["timestep"], inp["encoder_hidden_states"], **kw)
    taps = model.debug_taps()
    checked = 0
    for name, got in taps.items():
        if name in oracle.taps:
            r = oracle.taps[name].permute(0, 2, 3, 1).reshape(-1, got.shape[1])
            assert (got - r).abs().max().item() < 0.03 * max(1.0, r.abs().max().item()), name
            checked += 1
    assert checked >= 20


def test_fp16_io_tuple_return_and_skip_temporal(built_lib):
    cfg, model, oracle = _setup("musev", (64, 128, 128, 128), built_lib, io_dtype=torch.float16)
    inp = make_inputs(cfg, batch=2, frames=3, h=8, w=8, n_vis_cond=1, seed=5)
    kw = _call_kwargs(inp)
    out = model(_to(inp["sample"], dev, torch.float16), 301, _to(inp["encoder_hidden_states"], dev, torch.float16),
                return_dict=False, **{k: _to(v, dev, torch.float16) for k, v in kw.items()})
    assert isinstance(out, tuple) and out[0].dtype == torch.float16
    ref = oracle(inp["sample"].half().float(), 301, inp["encoder_hidden_states"].half().float(), **kw)
    assert (out[0].float() - ref).abs().max().item() < FWD_TOL
    # skip_temporal_layers=True is the first-frame (text-to-image) mode of the predictor
    o2 = model(_to(inp["sample"], dev, torch.float16), 301, _to(inp["encoder_hidden_states"], dev, torch.float16),
               skip_temporal_layers=True, **{k: _to(v, dev, torch.float16) for k, v in kw.items()}).sample
    r2 = oracle(inp["sample"].half().float(), 301, inp["encoder_hidden_states"].half().float(), skip_temporal_layers=True, **kw)
    assert (o2.float() - r2).abs().max().item() < FWD_TOL
    assert model.skip_temporal_layers is False                     # restored after the call (unet_3d_condition.py:1275)


def test_controlnet_residuals_and_separate_vis_cond(built_lib):
    cfg, model, oracle = _setup("musev", (64, 128, 128, 128), built_lib)
    inp = make_inputs(cfg, batch=2, frames=3, h=8, w=8, n_vis_cond=1, seed=6)
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 8, 8)] * 3 + [(64, 4, 4)] + [(128, 4, 4)] * 2 + [(128, 2, 2)] + [(128, 2, 2)] * 2 + [(128, 1, 1)] * 3
    NF = 2 * 4
    down = [torch.randn(NF, c, h, w, generator=g) * 0.1 for (c, h, w) in shapes]
    mid = torch.randn(NF, 128, 1, 1, generator=g) * 0.1
    # pass the vis-cond frame separately (unet_3d_condition.py:875-882) instead of pre-concatenated
    sample, vc = inp["sample"][:, :, 1:], inp["sample"][:, :, :1]
    kw = dict(sample_index=inp["sample_index"], vision_condition_frames_sample=vc,
              vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"], sample_frame_rate=8,
              down_block_additional_residuals=down, mid_block_additional_residual=mid)
    ref = oracle(sample, 301, inp["encoder_hidden_states"], **kw)
    out = model(sample.to(dev), 301, inp["encoder_hidden_states"].to(dev), **{k: _to(v, dev, torch.float32) for k, v in kw.items()}).sample
    assert (out - ref).abs().max().item() < FWD_TOL


def test_error_behaviour(built_lib):
    from musev_b200._capi import MvbError
    from musev_b200.unet import UNet3DConditionModel
    cfg = preset_config("musev", block_out_channels=(64, 128, 128, 128))
    model = UNet3DConditionModel(cfg, device=dev)
    x = torch.zeros(2, 4, 3, 8, 8, device=dev)
    enc = torch.zeros(2, 77, 768, device=dev)
    with pytest.raises(RuntimeError, match="weights not loaded"):
        model(x, 1, enc)
    sd = make_state_dict(cfg, seed=0)
    bad = dict(sd)
    bad.pop("conv_in.weight")
    with pytest.raises(RuntimeError, match="missing"):
        model.load_state_dict(bad)
    model.load_state_dict(sd)
    with pytest.raises(ValueError, match="only support ndim"):
        model(x, 1, enc[0])
    with pytest.raises(MvbError, match="divisible"):
        model(torch.zeros(2, 4, 3, 6, 6, device=dev), 1, enc)
    with pytest.raises(NotImplementedError):
        model(x, 1, enc, class_labels=torch.zeros(2))


@pytest.mark.parametrize("preset", ["musev", "musev_referencenet"])
def test_parallel_denoise_loop_vs_reference_golden(built_lib, preset):
    """2 DDIM steps x 3 overlapping windows through ParallelDenoiser vs the loop run with the imported reference."""
    from musev_b200.pipeline import ParallelDenoiser
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    g = torch.load(os.path.join(GOLDEN, f"loop_{preset}_narrow.pt"))
    m = g["meta"]
    cfg, model, oracle = _setup(preset, tuple(m["block_out_channels"]), built_lib)
    gen = torch.Generator().manual_seed(m["input_seed"])
    latents = torch.randn(1, 4, m["T"], m["h"], m["w"], generator=gen)
    cond = torch.randn(1, 4, 1, m["h"], m["w"], generator=gen) * 0.5
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=gen)
    extra = make_inputs(cfg, batch=2, frames=1, h=m["h"], w=m["w"], seed=m["input_seed"])
    kw = {k: _to(extra[k], dev, torch.float32) for k in ("down_block_refer_embs", "mid_block_refer_emb", "vision_clip_emb") if k in extra}
    kw["ip_adapter_scale"] = 1.0
    den = ParallelDenoiser(model, DDIMScheduler(**SD15_DDIM_CONFIG))
    res = den(latents.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=m["steps"], guidance_scale=m["guidance_scale"],
              context_frames=m["context_frames"], context_overlap=m["context_overlap"], motion_speed=8, unet_kwargs=kw)
    assert res.windows == m["contexts"]
    err = (res.latents.cpu() - g["latents"]).abs().max().item()
    _record(f"loop_{preset}_narrow_2step_vs_reference_golden", err)
    # CFG multiplies the per-forward eps error by up to (2 g - 1) = 6 and two steps accumulate; measured 1.5e-2 .. 2.6e-2
    assert err < 6e-2, err
    with pytest.raises(NotImplementedError):
        den(latents.to(dev), cond.to(dev), prompt.to(dev), guidance_scale=1.0)


def test_full_size_properties(built_lib):
    """BASELINE config-2 shape (B=2 CFG, 16+1 frames, 64x64 latents, full width): properties that do not need the
    oracle -- determinism, batch independence (swapping the CFG halves swaps the outputs), finite output."""
    from musev_b200.unet import UNet3DConditionModel
    cfg = preset_config("musev")
    model = UNet3DConditionModel(cfg, device=dev, dtype=torch.float16)
    model.load_state_dict(make_state_dict(cfg, seed=0, dtype=torch.float16))
    inp = make_inputs(cfg, batch=2, frames=16, h=64, w=64, n_vis_cond=1)
    x, enc = inp["sample"].to(dev).half(), inp["encoder_hidden_states"].to(dev).half()
    kw = dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"], sample_frame_rate=8)
    a = model(x, 601, enc, **kw).sample
    b = model(x, 601, enc, **kw).sample
    assert torch.equal(a, b) and torch.isfinite(a).all() and 0.2 < a.float().std().item() < 2.0
    c = model(x.flip(0), 601, enc.flip(0), **kw).sample
    assert (c.flip(0).float() - a.float()).abs().max().item() < 2e-3


def test_full_size_vs_eager_fp16_and_timing(built_lib):
    """Config-2 shape, full width: the engine against the oracle run the way the reference runs in production -- plain
    PyTorch fp16 on the same GPU (cuDNN / cuBLAS / SDPA flash kernels). Checks parity at the full size (both sides carry
    fp16 rounding, so the bound is looser than FWD_TOL) and records the eager time next to the engine time: the
    practical "beat this" number of SURVEY.md section 8(d). The timing line goes to stdout and gpurun_out/."""
    import json, os
    from musev_b200.unet import UNet3DConditionModel
    from oracle.unet3d_oracle import UNet3DOracle
    cfg = preset_config("musev")
    sd16 = make_state_dict(cfg, seed=0, dtype=torch.float16)
    model = UNet3DConditionModel(cfg, device=dev, dtype=torch.float16)
    model.load_state_dict(sd16)
    oracle = UNet3DOracle(cfg, sd16, device=dev, dtype=torch.float16)
    del sd16
    inp = make_inputs(cfg, batch=2, frames=16, h=64, w=64, n_vis_cond=1)
    x, enc = inp["sample"].to(dev).half(), inp["encoder_hidden_states"].to(dev).half()
    kw = dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"], sample_frame_rate=8)

    def timed(fn, iters=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) / iters

    with torch.no_grad():
        got, ms_engine = timed(lambda: model(x, 601, enc, **kw).sample)
        ref, ms_eager = timed(lambda: oracle(x, 601, enc, **kw))
    err = (got.float() - ref.float()).abs().max().item()
    line = {"shape": "B=2 T=16+1 64x64 musev", "engine_ms": ms_engine, "eager_fp16_torch_ms": ms_eager,
            "speedup": ms_eager / ms_engine, "max_abs_diff": err}
    print("EAGER_BASELINE " + json.dumps(line), flush=True)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(line, open("gpurun_out/eager_baseline.json", "w"))
    except OSError:
        pass
    assert torch.isfinite(got).all() and err < 4e-2
    assert ms_engine < ms_eager


def test_parallel_denoise_loop_euler_and_eta(built_lib):
    """The loop with the predictor's default sampler (EulerDiscreteScheduler: scaled model input + affine fused step,
    pipeline_controlnet_predictor.py:258-261) against the oracle loop around the Euler oracle; and DDIM with eta > 0 runs the
    fused noisy step (scheduling_ddim.py:266-295) deterministically for a seeded generator."""
    from musev_b200.pipeline import ParallelDenoiser
    from musev_b200.samplers import EulerDiscreteScheduler
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    from oracle.pipeline_oracle import denoise_loop
    from oracle.sampler_oracle import EulerOracle
    cfg, model, oracle = _setup("musev", (64, 128, 128, 128), built_lib)
    gen = torch.Generator().manual_seed(17)
    T, h, w = 12, 16, 16
    kw = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
              timestep_spacing="leading", steps_offset=1)
    sched = EulerDiscreteScheduler(**kw)
    latents = torch.randn(1, 4, T, h, w, generator=gen) * float(sched.init_noise_sigma)
    cond = torch.randn(1, 4, 1, h, w, generator=gen) * 0.5
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=gen)
    out = ParallelDenoiser(model, sched)(latents.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=3, guidance_scale=3.5,
                                         context_frames=8, context_overlap=4).latents.cpu()
    ref = denoise_loop(lambda s, t, e, **k: oracle(s, t, e, **k).cpu(), EulerOracle(**kw), latents, cond, prompt, 3, 3.5,
                       context_frames=8, context_overlap=4)
    err = (out - ref).abs().max().item() / float(sched.init_noise_sigma)
    _record("loop_euler_narrow_3step_rel", err)
    assert err < 2e-2, err
    # eta > 0
    den = ParallelDenoiser(model, DDIMScheduler(**SD15_DDIM_CONFIG))
    lat1 = latents / float(sched.init_noise_sigma)
    a = den(lat1.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=2, guidance_scale=3.5, context_frames=8,
            context_overlap=4, eta=0.5, generator=torch.Generator(device=dev).manual_seed(3), noise_type="video_fusion").latents
    b = den(lat1.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=2, guidance_scale=3.5, context_frames=8,
            context_overlap=4, eta=0.5, generator=torch.Generator(device=dev).manual_seed(3), noise_type="video_fusion").latents
    c = den(lat1.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=2, guidance_scale=3.5, context_frames=8,
            context_overlap=4).latents
    assert torch.equal(a, b) and torch.isfinite(a).all() and (a - c).abs().max().item() > 1e-2


@pytest.mark.parametrize("preset,frames,h,w", [("musev_referencenet", 12, 64, 64), ("musev", 8, 64, 96), ("musev_referencenet", 16, 64, 64)])
def test_full_width_other_window_shapes_vs_fp32_oracle(built_lib, preset, frames, h, w):
    """Full-width forward at the other shapes the BASELINE configs produce: the short last windows of configs 3 / 4
    (T = 12+1, 8+1), 512x768 (64x96 latents, config 5) and the heavier `musev_referencenet` preset at the
    config-2 shape -- against the fp32 oracle on the same fp16-rounded weights (same bound as the small-shape forward test)."""
    from musev_b200.unet import UNet3DConditionModel
    from oracle.unet3d_oracle import UNet3DOracle
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = preset_config(preset)
    sd16 = make_state_dict(cfg, seed=0, dtype=torch.float16)
    model = UNet3DConditionModel(cfg, device=dev, dtype=torch.float32)
    model.load_state_dict(sd16)
    oracle = UNet3DOracle(cfg, sd16, device=dev, dtype=torch.float32)
    del sd16
    inp = make_inputs(cfg, batch=2, frames=frames, h=h, w=w, n_vis_cond=1, seed=99)
    kw = _call_kwargs(inp, 8, 1.0)
    with torch.no_grad():
        ref = oracle(inp["sample"], 451, inp["encoder_hidden_states"], **kw)
        out = model(inp["sample"].to(dev), 451, inp["encoder_hidden_states"].to(dev),
                    **{k: _to(v, dev, torch.float32) for k, v in kw.items()}).sample
        # yardstick at this size: the same oracle run the way the reference runs (eager PyTorch fp16)
        oracle.sd = {k: v.half() for k, v in oracle.sd.items()}
        oracle.dtype = torch.float16
        ref16 = oracle(inp["sample"], 451, inp["encoder_hidden_states"], **kw).float()
    err = (out - ref).abs().max().item()
    err16 = (ref16 - ref).abs().max().item()
    rms, rms16 = (out - ref).pow(2).mean().sqrt().item(), (ref16 - ref).pow(2).mean().sqrt().item()
    _record(f"fwd_full_{preset}_T{frames + 1}_{h}x{w}_vs_oracle", err)
    _record(f"fwd_full_{preset}_T{frames + 1}_{h}x{w}_ref16_vs_oracle", err16)
    # the max over 10^6..10^7 outputs sits higher than in the small-shape test (measured 1.3e-2 for musev_referencenet at
    # 64x64); what matters is that the engine is no further from fp32 than the reference's own fp16 arithmetic
    assert torch.isfinite(out).all() and err < 2.5e-2 and rms < 1.2 * rms16 and err < 1.5 * err16, (err, err16, rms, rms16)


def test_tensor_map_cache_hits_on_repeated_forward(built_lib):
    """Encoded TMA descriptors are memoized: the second forward of the same shapes through the same workspace re-uses them
    (hits, a handful of misses for freshly allocated torch outputs at most) and reproduces the first bit for bit."""
    import os
    from musev_b200 import _capi
    from musev_b200.unet import UNet3DConditionModel
    if os.environ.get("MVB_TMAP_CACHE") == "0":
        pytest.skip("cache disabled by MVB_TMAP_CACHE=0")
    cfg = preset_config("musev", block_out_channels=(64, 128, 128, 128))
    model = UNet3DConditionModel(cfg, device=dev, dtype=torch.float16)
    model.load_state_dict(make_state_dict(cfg, seed=0, dtype=torch.float16))
    inp = make_inputs(cfg, batch=2, frames=8, h=16, w=16, n_vis_cond=1, seed=3)
    kw = _call_kwargs(inp, 8, 1.0)
    x, enc = inp["sample"].to(dev).half(), inp["encoder_hidden_states"].to(dev).half()
    kw = {k: _to(v, dev, torch.float16) for k, v in kw.items()}
    a = model(x, 301, enc, **kw).sample.clone()
    ok = False
    for _ in range(3):      # the process-wide table is cleared wholesale when it fills up: that may fall into one attempt
        h0, m0 = _capi.tensor_map_cache_stats()
        b = model(x, 301, enc, **kw).sample.clone()
        h1, m1 = _capi.tensor_map_cache_stats()
        torch.cuda.synchronize()
        assert torch.equal(a, b)
        if h1 - h0 > 500 and (m1 - m0) * 20 < (h1 - h0):
            ok = True
            break
    assert ok, (h0, m0, h1, m1)
