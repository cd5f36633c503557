"""VAE decode on the engine against the oracle and the reference golden samples
(tests/golden/vae_*.pt, produced by the unmodified diffusers AutoencoderKL.decode)."""
import os

import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.mark.parametrize("tag", ["narrow", "full"])
def test_vae_decode_vs_oracle_and_reference_golden(built_lib, tag):
    from musev_b200.schema import VAEConfig
    from musev_b200.synth import make_state_dict
    from musev_b200.vae import AutoencoderKLDecoder
    from oracle.vae_oracle import VAEDecoderOracle
    g = torch.load(os.path.join(GOLDEN, f"vae_{tag}.pt"))
    m = g["meta"]
    cfg = VAEConfig(block_out_channels=tuple(m["block_out_channels"]))
    sd16 = {k: v.half() for k, v in make_state_dict(cfg, seed=m["weight_seed"]).items()}
    vae = AutoencoderKLDecoder(cfg, device=dev, dtype=torch.float32, frames_per_call=1)
    vae.load_state_dict(sd16)
    oracle = VAEDecoderOracle(cfg, {k: v.float() for k, v in sd16.items()}, device=dev)
    lat = torch.randn(1, 4, m["frames"], m["h"], m["w"], generator=torch.Generator().manual_seed(m["input_seed"])) * 0.18215 * 1.2
    z = lat.permute(0, 2, 1, 3, 4).reshape(m["frames"], 4, m["h"], m["w"]) / cfg.scaling_factor
    raw = vae.decode(z.to(dev)).sample
    ref = oracle.decode(z)
    assert list(raw.shape) == m["shape"] and torch.isfinite(raw).all()
    scale = max(1.0, ref.abs().max().item())
    assert (raw - ref).abs().max().item() < 1.5e-2 * scale
    idx = torch.randint(0, raw.numel(), (m["n_samples"],), generator=torch.Generator().manual_seed(m["sample_seed"]))
    assert (raw.reshape(-1)[idx].cpu() - g["raw"]).abs().max().item() < 2e-2 * scale            # the reference (fp32 weights)
    video = vae.decode_latents(lat.to(dev))                       # [1, 3, f, H, W] in [0, 1]
    assert video.shape == (1, 3, m["frames"], 8 * m["h"], 8 * m["w"]) and video.min() >= 0 and video.max() <= 1
    flat = video.permute(0, 2, 1, 3, 4).reshape(-1)
    assert (flat[idx].cpu() - g["img"]).abs().max().item() < 1e-2 * scale
    assert (video - oracle.decode_latents(lat)).abs().max().item() < 1e-2 * scale


def test_vae_decode_512(built_lib):
    """SD-1.5 decoder at the headline size (64x64 latents -> 512x512), 2 frames, against the oracle run as eager fp32 on the
    GPU; also exercises chunked decoding (frames_per_call 1 vs 2 give the same frames)."""
    from musev_b200.schema import VAEConfig
    from musev_b200.synth import make_state_dict
    from musev_b200.vae import AutoencoderKLDecoder
    from oracle.vae_oracle import VAEDecoderOracle
    cfg = VAEConfig()
    sd16 = {k: v.half() for k, v in make_state_dict(cfg, seed=11).items()}
    vae = AutoencoderKLDecoder(cfg, device=dev, dtype=torch.float32, frames_per_call=2)
    vae.load_state_dict(sd16)
    lat = torch.randn(1, 4, 2, 64, 64, generator=torch.Generator().manual_seed(4)) * 0.18215
    video = vae.decode_latents(lat.to(dev))
    ref = VAEDecoderOracle(cfg, {k: v.float() for k, v in sd16.items()}, device=dev).decode_latents(lat)
    assert video.shape == (1, 3, 2, 512, 512)
    assert (video - ref).abs().max().item() < 1.5e-2
    vae.frames_per_call = 1
    assert torch.equal(vae.decode_latents(lat.to(dev)), video)
