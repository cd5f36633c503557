"""CPU tests: the oracle (oracle/) against the reference's own known-answer tests and against fixtures produced by
running the UNMODIFIED reference (oracle/make_golden.py -> tests/golden/)."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from musev_b200.schema import UNetConfig, preset_config
from musev_b200.synth import make_inputs, make_state_dict
from oracle.pipeline_oracle import SD15_DDIM, DDIMOracle, denoise_loop, prepare_global_context
from oracle.unet3d_oracle import UNet3DOracle


# ---- DDIM: upstream KATs, diffusers/tests/schedulers/test_scheduler_ddim.py:46-54,102-176
def _dummy_sample_deter():
    n = 4 * 3 * 8 * 8
    return (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2)


def _dummy_noise_deter():
    n = 4 * 3 * 8 * 8
    return (torch.arange(n).flip(-1).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2)


def _full_loop(**cfg):
    kw = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)
    kw.update(cfg)
    s = DDIMOracle(**kw)
    s.set_timesteps(10)
    x = _dummy_sample_deter()
    for t in s.timesteps:
        x, _ = s.step(x * t / (t + 1), t, x, 0.0)
    return x


@pytest.mark.parametrize("cfg,exp_sum,exp_mean", [
    ({}, 172.0067, 0.223967),
    ({"prediction_type": "v_prediction"}, 52.5302, 0.0684),
    ({"set_alpha_to_one": True, "beta_start": 0.01}, 149.8295, 0.1951),
    ({"set_alpha_to_one": False, "beta_start": 0.01}, 149.0784, 0.1941),
])
def test_ddim_full_loop_kat(cfg, exp_sum, exp_mean):
    x = _full_loop(**cfg)
    assert abs(x.abs().sum().item() - exp_sum) < 1e-2
    assert abs(x.abs().mean().item() - exp_mean) < 1e-3


def test_ddim_with_noise_kat():
    s = DDIMOracle(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)
    s.set_timesteps(10)
    ts = s.timesteps[8:]
    x = s.add_noise(_dummy_sample_deter(), _dummy_noise_deter(), ts[:1])
    for t in ts:
        x, _ = s.step(x * t / (t + 1), t, x, 0.0)
    assert abs(x.abs().sum().item() - 354.5418) < 1e-2
    assert abs(x.abs().mean().item() - 0.4616) < 1e-3


def test_ddim_variance_and_offset_kat():
    s = DDIMOracle(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)
    for (t, p, v) in [(0, 0, 0.0), (420, 400, 0.14771), (980, 960, 0.32460), (487, 486, 0.00979), (999, 998, 0.02)]:
        assert abs(float(s._get_variance(t, p)) - v) < 1e-5
    s2 = DDIMOracle(steps_offset=1)
    s2.set_timesteps(5)
    assert s2.timesteps.tolist() == [801, 601, 401, 201, 1]


def test_ddim_matches_reference_scheduler_fixture():
    g = torch.load(os.path.join(GOLDEN, "ddim_sd15.pt"))
    s = DDIMOracle(**SD15_DDIM)
    s.set_timesteps(20)
    assert s.timesteps.tolist() == g["timesteps"].tolist() == [951 - 50 * i for i in range(20)]
    for t, ref in g["prev"].items():
        prev, _ = s.step(g["eps"], int(t), g["x"], 0.0)
        assert torch.allclose(prev, ref, atol=1e-6, rtol=1e-6)


# ---- ResnetBlock2D: upstream KAT, diffusers/tests/models/test_layers_utils.py:224-236
def test_resnet_block_kat():
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 64, 64)
    temb = torch.randn(1, 128)
    conv1 = torch.nn.Conv2d(32, 32, 3, padding=1)      # same construction order (= RNG stream) as ResnetBlock2D.__init__
    lin = torch.nn.Linear(128, 32)
    conv2 = torch.nn.Conv2d(32, 32, 3, padding=1)
    sd = {"r.norm1.weight": torch.ones(32), "r.norm1.bias": torch.zeros(32), "r.norm2.weight": torch.ones(32),
          "r.norm2.bias": torch.zeros(32), "r.conv1.weight": conv1.weight.data, "r.conv1.bias": conv1.bias.data,
          "r.conv2.weight": conv2.weight.data, "r.conv2.bias": conv2.bias.data,
          "r.time_emb_proj.weight": lin.weight.data, "r.time_emb_proj.bias": lin.bias.data}
    o = UNet3DOracle(UNetConfig(norm_eps=1e-6), sd)
    with torch.no_grad():
        out = o.resnet(sample, temb, "r")
    exp = torch.tensor([-1.9010, -0.2974, -0.8245, -1.3533, 0.8742, -0.9645, -2.0584, 1.3387, -0.4746])
    assert torch.allclose(out[0, -1, -3:, -3:].flatten(), exp, atol=1e-3)


# ---- window schedules vs the reference's prepare_global_context
def test_contexts_match_reference():
    from musev_b200 import context as host_ctx
    cases = json.load(open(os.path.join(GOLDEN, "contexts.json")))
    assert len(cases) >= 10
    for c in cases:
        args = (c["schedule"], 20, c["T"], c["window"], 1, c["overlap"], 1)
        assert [w[0] for w in prepare_global_context(*args)] == c["contexts"], c
        assert [w[0] for w in host_ctx.prepare_global_context(*args)] == c["contexts"], c


# ---- UNet3D forward vs the imported reference (narrow width; both presets)
def _oracle_forward(preset, meta, sd=None, dtype=torch.float32):
    cfg = preset_config(preset, block_out_channels=tuple(meta["block_out_channels"]))
    sd = sd or make_state_dict(cfg, seed=meta["weight_seed"])
    inp = make_inputs(cfg, batch=meta["batch"], frames=meta["frames"], h=meta["h"], w=meta["w"],
                      n_vis_cond=meta["n_vis_cond"], seed=meta["input_seed"])
    o = UNet3DOracle(cfg, sd)
    return o(inp["sample"], meta["timestep"], inp["encoder_hidden_states"], sample_index=inp["sample_index"],
             vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"],
             sample_frame_rate=meta["sample_frame_rate"], down_block_refer_embs=inp.get("down_block_refer_embs"),
             mid_block_refer_emb=inp.get("mid_block_refer_emb"), vision_clip_emb=inp.get("vision_clip_emb"),
             ip_adapter_scale=meta["ip_adapter_scale"])


@pytest.mark.parametrize("preset", ["musev", "musev_referencenet"])
def test_unet_oracle_matches_reference(preset):
    g = torch.load(os.path.join(GOLDEN, f"unet_{preset}_narrow.pt"))
    out = _oracle_forward(preset, g["meta"])
    err = (out - g["out"]).abs().max().item()
    assert g["out"].std().item() > 0.3
    assert err < 5e-5, err   # fp32 vs fp32, same weights: only summation-order noise


@pytest.mark.parametrize("preset", ["musev", "musev_referencenet"])
def test_denoise_loop_oracle_matches_reference(preset):
    g = torch.load(os.path.join(GOLDEN, f"loop_{preset}_narrow.pt"))
    m = g["meta"]
    cfg = preset_config(preset, block_out_channels=tuple(m["block_out_channels"]))
    o = UNet3DOracle(cfg, make_state_dict(cfg, seed=m["weight_seed"]))
    gen = torch.Generator().manual_seed(m["input_seed"])
    latents = torch.randn(1, 4, m["T"], m["h"], m["w"], generator=gen)
    cond = torch.randn(1, 4, 1, m["h"], m["w"], generator=gen) * 0.5
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=gen)
    extra = make_inputs(cfg, batch=2, frames=1, h=m["h"], w=m["w"], seed=m["input_seed"])
    kw = {k: extra[k] for k in ("down_block_refer_embs", "mid_block_refer_emb", "vision_clip_emb") if k in extra}
    out = denoise_loop(lambda *a, **k: o(*a, **k), DDIMOracle(**SD15_DDIM), latents, cond, prompt, m["steps"],
                       m["guidance_scale"], context_frames=m["context_frames"], context_overlap=m["context_overlap"],
                       motion_speed=8, unet_kwargs=dict(kw, ip_adapter_scale=1.0))
    assert len(m["contexts"]) >= 3                        # overlapping windows were exercised
    err = (out - g["latents"]).abs().max().item()
    assert err < 2e-4, err


# ---- ControlNet encoder: oracle vs the unmodified diffusers ControlNetModel of the reference tree
@pytest.mark.parametrize("tag", ["narrow", "narrow_guess", "full"])
def test_controlnet_oracle_matches_reference(tag):
    from musev_b200.schema import ControlNetConfig, controlnet_param_shapes
    from musev_b200.synth import make_controlnet_inputs
    from oracle.controlnet_oracle import ControlNetOracle
    g = torch.load(os.path.join(GOLDEN, f"controlnet_{tag}.pt"))
    m = g["meta"]
    cfg = ControlNetConfig(block_out_channels=tuple(m["block_out_channels"]))
    sd = make_state_dict(cfg, seed=m["weight_seed"])
    assert set(sd) == set(controlnet_param_shapes(cfg))
    oracle = ControlNetOracle(cfg, sd)
    inp = make_controlnet_inputs(cfg, frames=m["frames"], h=m["h"], w=m["w"], seed=m["input_seed"])
    down, mid = oracle(inp["sample"], m["timestep"], inp["encoder_hidden_states"], controlnet_cond=inp["controlnet_cond"],
                       conditioning_scale=m["conditioning_scale"], guess_mode=m["guess_mode"])
    maps = list(down) + [mid]
    assert [list(t.shape) for t in maps] == m["shapes"]
    for k, mp in enumerate(maps):
        flat = mp.reshape(-1)
        idx = torch.randint(0, flat.numel(), (m["n_samples"],), generator=torch.Generator().manual_seed(m["sample_seed_base"] + k))
        ref = g["samples"][k]
        assert (flat[idx] - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item()), f"map {k}"
        assert abs(float(flat.mean()) - g["stats"][k][0]) < 1e-5 and abs(float(flat.abs().mean()) - g["stats"][k][1]) < 1e-5
    # the pipeline passes the condition embedding pre-computed (`controlnet_cond_latents`, pipeline_controlnet.py:1258)
    lat = oracle.cond_embedding(inp["controlnet_cond"])
    down2, mid2 = oracle(inp["sample"], m["timestep"], inp["encoder_hidden_states"], controlnet_cond_latents=lat,
                         conditioning_scale=m["conditioning_scale"], guess_mode=m["guess_mode"])
    assert torch.equal(mid, mid2) and all(torch.equal(a, b) for a, b in zip(down, down2))


@pytest.mark.parametrize("tag", ["narrow", "narrow_t2"])
def test_referencenet_oracle_matches_reference_golden(tag):
    """oracle/referencenet_oracle.py against samples of the unmodified musev ReferenceNet2D (oracle/make_golden.py)."""
    import os
    from conftest import GOLDEN
    from musev_b200.schema import ReferenceNetConfig
    from musev_b200.synth import make_referencenet_inputs, make_state_dict
    from oracle.referencenet_oracle import ReferenceNetOracle
    g = torch.load(os.path.join(GOLDEN, f"referencenet_{tag}.pt"))
    m = g["meta"]
    cfg = ReferenceNetConfig(block_out_channels=tuple(m["block_out_channels"]))
    o = ReferenceNetOracle(cfg, make_state_dict(cfg, seed=m["weight_seed"]))
    inp = make_referencenet_inputs(cfg, batch=m["batch"], n_ref=m["n_ref"], h=m["h"], w=m["w"], seed=m["input_seed"])
    down, mid = o(inp["sample"], 0, inp["encoder_hidden_states"], num_frames=m["n_ref"], return_ndim=5)
    maps = list(down) + [mid]
    assert [list(x.shape) for x in maps] == m["shapes"]
    for k, mp in enumerate(maps):
        flat = mp.reshape(-1)
        idx = torch.randint(0, flat.numel(), (m["n_samples"],), generator=torch.Generator().manual_seed(m["sample_seed_base"] + k))
        assert (flat[idx] - g["samples"][k]).abs().max().item() < 2e-5 * max(1.0, g["samples"][k].abs().max().item()), k


# ------------------------------------------------------------------ other samplers
@pytest.mark.parametrize("pred,exp_sum,exp_mean", [("epsilon", 10.0807, 0.0131), ("v_prediction", 0.0002, 2.2676e-06)])
def test_euler_full_loop_kat(pred, exp_sum, exp_mean):
    """diffusers/tests/schedulers/test_scheduler_euler.py:41-110 replayed on the oracle."""
    from oracle.sampler_oracle import EulerOracle
    s = EulerOracle(num_train_timesteps=1100, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", prediction_type=pred)
    s.set_timesteps(10)
    x = _dummy_sample_deter() * s.init_noise_sigma
    for t in s.timesteps:
        x = s.scale_model_input(x, t)          # the upstream test feeds the scaled sample back into `step`
        x, _ = s.step(x * t / (t + 1), t, x)
    assert abs(x.abs().sum().item() - exp_sum) < 1e-2
    assert abs(x.abs().mean().item() - exp_mean) < 1e-3


@pytest.mark.parametrize("steps,exp_sum,exp_mean", [(1, 18.7097, 0.0244), (10, 197.7616, 0.2575)])
def test_lcm_full_loop_kat(steps, exp_sum, exp_mean):
    """diffusers/tests/schedulers/test_scheduler_lcm.py:209-244 replayed on the oracle (global generator seed 0)."""
    from oracle.sampler_oracle import LCMOracle
    s = LCMOracle()
    s.set_timesteps(steps)
    g = torch.manual_seed(0)
    x = _dummy_sample_deter()
    for t in s.timesteps:
        x, _, _ = s.step(x * t / (t + 1), t, x, g)
    assert abs(x.abs().sum().item() - exp_sum) < 1e-3
    assert abs(x.abs().mean().item() - exp_mean) < 1e-3


def test_sampler_oracles_match_imported_musev_schedulers():
    """tests/golden/samplers_sd15.pt: loops run by the imported musev.schedulers classes (oracle/make_golden.py)."""
    from oracle.sampler_oracle import EulerOracle, LCMOracle
    g = torch.load(os.path.join(GOLDEN, "samplers_sd15.pt"))
    e = EulerOracle(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                    timestep_spacing="leading", steps_offset=1)
    e.set_timesteps(20)
    assert torch.equal(e.timesteps, g["euler"]["timesteps"]) and torch.allclose(e.sigmas, g["euler"]["sigmas"], atol=1e-6)
    assert abs(float(e.init_noise_sigma) - g["euler"]["init_noise_sigma"]) < 1e-5
    x = g["x"] * e.init_noise_sigma
    for i, t in enumerate(e.timesteps):
        xs = e.scale_model_input(x, t)
        x, _ = e.step(xs * t / (t + 1), t, x)
        assert (x - g["euler"]["trace"][i]).abs().max().item() < 1e-4
    l = LCMOracle()
    l.set_timesteps(4)
    assert l.timesteps.tolist() == g["lcm"]["timesteps"].tolist()
    gen = torch.Generator().manual_seed(g["lcm"]["noise_seed"])
    x = g["x"].clone()
    for i, t in enumerate(l.timesteps):
        x, den, _ = l.step(x * t / (t + 1), t, x, gen)
        assert (x - g["lcm"]["trace"][i]).abs().max().item() < 1e-5 and (den - g["lcm"]["denoised"][i]).abs().max().item() < 1e-5


@pytest.mark.parametrize("tag", ["narrow", "full"])
def test_vae_oracle_matches_reference_golden(tag):
    """oracle/vae_oracle.py against samples of the unmodified diffusers AutoencoderKL.decode (oracle/make_golden.py)."""
    from musev_b200.schema import VAEConfig
    from oracle.vae_oracle import VAEDecoderOracle
    g = torch.load(os.path.join(GOLDEN, f"vae_{tag}.pt"))
    m = g["meta"]
    cfg = VAEConfig(block_out_channels=tuple(m["block_out_channels"]))
    o = VAEDecoderOracle(cfg, make_state_dict(cfg, seed=m["weight_seed"]))
    lat = torch.randn(1, 4, m["frames"], m["h"], m["w"], generator=torch.Generator().manual_seed(m["input_seed"])) * 0.18215 * 1.2
    z = lat.permute(0, 2, 1, 3, 4).reshape(m["frames"], 4, m["h"], m["w"]) / cfg.scaling_factor
    raw = o.decode(z)
    assert list(raw.shape) == m["shape"]
    idx = torch.randint(0, raw.numel(), (m["n_samples"],), generator=torch.Generator().manual_seed(m["sample_seed"]))
    assert (raw.reshape(-1)[idx] - g["raw"]).abs().max().item() < 2e-5
    img = o.decode_latents(lat)                                   # [1, 3, f, H, W]
    flat = img.permute(0, 2, 1, 3, 4).reshape(-1)
    assert (flat[idx] - g["img"]).abs().max().item() < 2e-5
