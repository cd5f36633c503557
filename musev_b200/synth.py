"""Seeded synthetic weights / inputs (no checkpoints exist offline).

Every tensor is drawn from its own generator seeded by (seed, crc32(name)), on CPU in fp32, so the build
container and the GPU box produce bit-identical values irrespective of creation order. The zero-initialised
tensors of the reference (temporal `proj_out`, `conv4`, ReferEmbFuse `to_out`) are drawn non-zero
and `temporal_weight` gets O(1) magnitudes with mixed signs, otherwise the temporal / reference paths would
contribute ~1e-5 and parity tests could not see bugs in them.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch

from .schema import (ControlNetConfig, ImageProjConfig, ReferenceNetConfig, UNetConfig, VAEConfig, controlnet_param_shapes,
                     image_proj_param_shapes, refer_emb_shapes, referencenet_param_shapes, unet_param_shapes,
                     vae_decoder_param_shapes)

_BRANCH_OUT = ("conv2.weight", "proj_out.weight", "to_out.0.weight", "ff.net.2.weight", "conv4.3.weight")
# the ControlNet's zero-initialised convolutions (controlnet.py:97-99,425-444) are drawn non-zero for the same reason
_ZERO_INIT = ("controlnet_cond_embedding.conv_out.weight", "controlnet_mid_block.weight")


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def make_state_dict(cfg, seed: int = 0, dtype: torch.dtype = torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights for a `UNetConfig` (denoiser) or a `ControlNetConfig` (ControlNet encoder)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    if isinstance(cfg, ReferenceNetConfig):
        shapes = referencenet_param_shapes(cfg)
    elif isinstance(cfg, ControlNetConfig):
        shapes = controlnet_param_shapes(cfg)
    elif isinstance(cfg, ImageProjConfig):
        shapes = image_proj_param_shapes(cfg)
    elif isinstance(cfg, VAEConfig):
        shapes = vae_decoder_param_shapes(cfg)
    else:
        shapes = unet_param_shapes(cfg)
    for name, shape in shapes.items():
        g = _gen(seed, name)
        if name.endswith("temporal_weight"):
            t = torch.empty(shape).uniform_(0.4, 0.9, generator=g)
            if zlib.crc32(name.encode()) & 1:
                t = -t  # the reference applies abs() (musev/models/resnet.py:128, temporal_transformer.py:299)
        elif len(shape) == 1:
            is_norm_w = name.endswith(".weight")
            t = torch.randn(shape, generator=g) * (0.1 if is_norm_w else 0.02)
            if is_norm_w:
                t = t + 1.0
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 0.5 if (name.endswith(_BRANCH_OUT) or name.endswith(_ZERO_INIT) or name.startswith("controlnet_down_blocks")) else 1.0
            t = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        sd[name] = t.to(dtype)
    return sd


def make_inputs(cfg: UNetConfig, batch: int, frames: int, h: int, w: int, n_vis_cond: int = 1, seed: int = 1234,
                n_ref: int = 1) -> Dict[str, object]:
    """Synthetic call arguments of `UNet3DConditionModel.forward` for one window.

    `sample` already contains the vision-condition frame(s) at the front, as in the pipeline
    (musev/pipelines/pipeline_controlnet.py:1921-1946)."""
    def r(name, *shape, scale=1.0):
        return torch.randn(*shape, generator=_gen(seed, name)) * scale

    T = frames + n_vis_cond
    out: Dict[str, object] = {
        "sample": r("sample", batch, cfg.in_channels, T, h, w),
        "encoder_hidden_states": r("encoder_hidden_states", batch, 77, cfg.cross_attention_dim),
        "sample_index": torch.arange(n_vis_cond, T),
        "vision_conditon_frames_sample_index": torch.arange(n_vis_cond) if n_vis_cond > 0 else None,
    }
    if cfg.ip_adapter_cross_attn:
        out["vision_clip_emb"] = r("vision_clip_emb", batch, 4, cfg.cross_attention_dim)
    if cfg.need_refer_emb:
        shapes, mid = refer_emb_shapes(cfg, h, w)
        out["down_block_refer_embs"] = [r(f"refer{i}", batch, c, n_ref, hh, ww) for i, (c, hh, ww) in enumerate(shapes)]
        out["mid_block_refer_emb"] = r("refer_mid", batch, mid[0], n_ref, mid[1], mid[2])
    return out


def make_controlnet_inputs(cfg: ControlNetConfig, frames: int, h: int, w: int, seed: int = 4321) -> Dict[str, object]:
    """Synthetic call arguments of `ControlNetModel.forward` as `get_controlnet_emb` issues it
    (musev/pipelines/pipeline_controlnet.py:1238-1262): `sample` is `(b t) c h w`, the prompt embedding is repeated
    per frame, the condition image is 8x the latent size."""
    def r(name, *shape, scale=1.0):
        return torch.randn(*shape, generator=_gen(seed, name)) * scale

    return {
        "sample": r("cn_sample", frames, cfg.in_channels, h, w),
        "encoder_hidden_states": r("cn_text", frames, 77, cfg.cross_attention_dim),
        "controlnet_cond": torch.rand(frames, cfg.conditioning_channels, 8 * h, 8 * w, generator=_gen(seed, "cn_cond")),
    }


def make_referencenet_inputs(cfg: ReferenceNetConfig, batch: int, n_ref: int, h: int, w: int, n_tokens: int = 4,
                             seed: int = 2468) -> Dict[str, object]:
    """Synthetic call arguments of `ReferenceNet2D.forward` as `get_referencenet_emb` issues it
    (musev/pipelines/pipeline_controlnet.py:918-929): `sample` = reference-image VAE latents (b t) c h w, timestep 0,
    `encoder_hidden_states` = the IP-Adapter image tokens [(b t), 4 n_img, 768]."""
    def r(name, *shape, scale=1.0):
        return torch.randn(*shape, generator=_gen(seed, name)) * scale

    return {
        "sample": r("rn_sample", batch * n_ref, cfg.in_channels, h, w, scale=0.7),
        "encoder_hidden_states": r("rn_tokens", batch * n_ref, n_tokens, cfg.cross_attention_dim),
        "num_frames": n_ref,
    }
