"""ControlNet encoder on the engine against the oracle and the reference golden samples."""
import os

import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.mark.parametrize("tag", ["narrow", "narrow_guess", "full"])
def test_controlnet_vs_oracle_and_reference_golden(built_lib, tag):
    from musev_b200.controlnet import ControlNetModel
    from musev_b200.schema import ControlNetConfig
    from musev_b200.synth import make_controlnet_inputs, make_state_dict
    from oracle.controlnet_oracle import ControlNetOracle
    g = torch.load(os.path.join(GOLDEN, f"controlnet_{tag}.pt"))
    m = g["meta"]
    cfg = ControlNetConfig(block_out_channels=tuple(m["block_out_channels"]))
    sd16 = {k: v.half() for k, v in make_state_dict(cfg, seed=m["weight_seed"]).items()}
    model = ControlNetModel(cfg, device=dev, dtype=torch.float32)
    model.load_state_dict(sd16)
    oracle = ControlNetOracle(cfg, {k: v.float() for k, v in sd16.items()}, device=dev)
    inp = make_controlnet_inputs(cfg, frames=m["frames"], h=m["h"], w=m["w"], seed=m["input_seed"])
    kw = dict(conditioning_scale=m["conditioning_scale"], guess_mode=m["guess_mode"])
    lat = oracle.cond_embedding(inp["controlnet_cond"].to(dev))
    down, mid = model(inp["sample"].to(dev), m["timestep"], inp["encoder_hidden_states"].to(dev),
                      controlnet_cond_latents=lat, return_dict=False, **kw)
    rdown, rmid = oracle(inp["sample"], m["timestep"], inp["encoder_hidden_states"], controlnet_cond_latents=lat, **kw)
    for k, (a, b) in enumerate(zip(list(down) + [mid], list(rdown) + [rmid])):
        err = (a.float() - b.float()).abs().max().item()
        assert err < 2e-2 * max(1.0, b.abs().max().item()), f"map {k}: {err}"
    # reference golden samples (fp32 weights): same bound as the UNet forward test
    for k, mp in enumerate(list(down) + [mid]):
        flat = mp.float().reshape(-1).cpu()
        idx = torch.randint(0, flat.numel(), (m["n_samples"],), generator=torch.Generator().manual_seed(m["sample_seed_base"] + k))
        assert (flat[idx] - g["samples"][k]).abs().max().item() < 3e-2 * max(1.0, g["samples"][k].abs().max().item())
    # the embedding path through the mirror (torch convs) gives the same maps
    down2, mid2 = model(inp["sample"].to(dev), m["timestep"], inp["encoder_hidden_states"].to(dev),
                        controlnet_cond=inp["controlnet_cond"].to(dev), return_dict=False, **kw)
    assert (mid2.float() - rmid.float()).abs().max().item() < 3e-2 * max(1.0, rmid.abs().max().item())


def test_controlnet_in_the_denoise_loop(built_lib):
    """Config-4 style loop at narrow width: every window-step runs the ControlNet encoder on the engine and feeds its 12 + 1
    residual maps to the UNet (pipeline_controlnet.py:1992-2067). Checked against the oracle loop with the ControlNet oracle."""
    from musev_b200.controlnet import ControlNetModel
    from musev_b200.pipeline import ParallelDenoiser, make_controlnet_fn
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    from musev_b200.schema import ControlNetConfig, preset_config
    from musev_b200.synth import make_state_dict
    from musev_b200.unet import UNet3DConditionModel
    from oracle.controlnet_oracle import ControlNetOracle
    from oracle.pipeline_oracle import SD15_DDIM, DDIMOracle, denoise_loop
    from oracle.unet3d_oracle import UNet3DOracle
    boc = (64, 128, 128, 128)
    cfg = preset_config("musev", block_out_channels=boc)
    ccfg = ControlNetConfig(block_out_channels=boc)
    sd = {k: v.half() for k, v in make_state_dict(cfg, seed=0).items()}
    csd = {k: v.half() for k, v in make_state_dict(ccfg, seed=3).items()}
    unet = UNet3DConditionModel(cfg, device=dev, dtype=torch.float32)
    unet.load_state_dict(sd)
    cnet = ControlNetModel(ccfg, device=dev, dtype=torch.float32)
    cnet.load_state_dict(csd)
    g = torch.Generator().manual_seed(21)
    T, h, w, n_vc = 12, 16, 16, 1
    latents = torch.randn(1, 4, T, h, w, generator=g)
    cond = torch.randn(1, 4, n_vc, h, w, generator=g) * 0.5
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    cn_lat = torch.randn(2, boc[0], n_vc + T, h, w, generator=g) * 0.3          # condition embedding of every frame
    den = ParallelDenoiser(unet, DDIMScheduler(**SD15_DDIM_CONFIG))
    fn = make_controlnet_fn(cnet, cn_lat.to(dev), prompt.to(dev), n_vc, controlnet_conditioning_scale=0.9)
    out = den(latents.to(dev), cond.to(dev), prompt.to(dev), num_inference_steps=2, guidance_scale=3.5, context_frames=8,
              context_overlap=4, controlnet_fn=fn).latents.cpu()
    uo = UNet3DOracle(cfg, {k: v.float() for k, v in sd.items()}, device=dev)
    co = ControlNetOracle(ccfg, {k: v.float() for k, v in csd.items()}, device=dev)
    ref = denoise_loop(lambda s, t, e, **k: uo(s, t, e, **k).cpu(), DDIMOracle(**SD15_DDIM), latents, cond, prompt, 2, 3.5,
                       context_frames=8, context_overlap=4,
                       controlnet=lambda x, t, e, **k: co(x, t, e, **k), controlnet_latents=cn_lat,
                       controlnet_conditioning_scale=0.9)
    no_cn = denoise_loop(lambda s, t, e, **k: uo(s, t, e, **k).cpu(), DDIMOracle(**SD15_DDIM), latents, cond, prompt, 2, 3.5,
                         context_frames=8, context_overlap=4)
    err = (out - ref).abs().max().item()
    assert (ref - no_cn).abs().max().item() > 10 * err, "the ControlNet residuals must matter in this test"
    assert err < 5e-2, err
