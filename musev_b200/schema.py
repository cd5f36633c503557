"""Configuration and state-dict schema of the denoiser (names/shapes of `UNet3DConditionModel.state_dict()`) and of
the per-window-step ControlNet encoder that feeds it.

Mirrors what the reference builds in musev/models/unet_3d_condition.py:213-610 for the two released presets
(musev/models/unet_loader.py:232-268); oracle/make_golden.py checks the generated schema against the
reference's own `state_dict()` key by key.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import asdict, dataclass, field
from typing import Dict, Tuple


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: int = 8          # number of heads in the reference (a diffusers naming accident)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    sample_size: int = 64
    # musev switches
    need_transformer_in: bool = True
    use_anivv1_cfg: bool = False
    resnet_2d_skip_time_act: bool = False
    keep_vision_condtion: bool = False
    need_refer_emb: bool = False
    ip_adapter_cross_attn: bool = False
    need_t2i_ip_adapter: bool = True      # reference-only self-attention toward the vision-condition frame
    preset: str = "musev"

    @property
    def heads(self) -> int:
        return self.attention_head_dim

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * 4

    def to_dict(self):
        return asdict(self)


def preset_config(name: str, **overrides) -> UNetConfig:
    """The two released configurations (musev/models/unet_loader.py:232-268)."""
    if name == "musev":
        cfg = UNetConfig(preset="musev")
    elif name in ("musev_referencenet", "musev_referencenet_pose"):
        cfg = UNetConfig(
            preset="musev_referencenet", need_transformer_in=False, use_anivv1_cfg=True,
            resnet_2d_skip_time_act=True, keep_vision_condtion=True, need_refer_emb=True,
            ip_adapter_cross_attn=True)
    else:
        raise ValueError(
            f"unsupport model_name={name}, only support musev, musev_referencenet, musev_referencenet_pose")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise ValueError(f"unknown config field {k}")
        setattr(cfg, k, tuple(v) if k == "block_out_channels" else v)
    return cfg


def _attn(prefix: str, C: int, kv_dim: int, out: Dict, ip: bool = False):
    out[f"{prefix}.to_q.weight"] = (C, C)
    out[f"{prefix}.to_k.weight"] = (C, kv_dim)
    out[f"{prefix}.to_v.weight"] = (C, kv_dim)
    out[f"{prefix}.to_out.0.weight"] = (C, C)
    out[f"{prefix}.to_out.0.bias"] = (C,)
    if ip:
        out[f"{prefix}.to_k_ip.weight"] = (C, kv_dim)
        out[f"{prefix}.to_v_ip.weight"] = (C, kv_dim)


def _tblock(prefix: str, C: int, cross_dim, out: Dict, ip: bool):
    for n in ("norm1", "norm2", "norm3"):
        out[f"{prefix}.{n}.weight"] = (C,)
        out[f"{prefix}.{n}.bias"] = (C,)
    _attn(f"{prefix}.attn1", C, C, out)
    _attn(f"{prefix}.attn2", C, cross_dim if cross_dim else C, out, ip=ip)
    out[f"{prefix}.ff.net.0.proj.weight"] = (8 * C, C)
    out[f"{prefix}.ff.net.0.proj.bias"] = (8 * C,)
    out[f"{prefix}.ff.net.2.weight"] = (C, 4 * C)
    out[f"{prefix}.ff.net.2.bias"] = (C,)


def _resnet(prefix: str, cin: int, C: int, temb: int, out: Dict):
    out[f"{prefix}.norm1.weight"] = (cin,)
    out[f"{prefix}.norm1.bias"] = (cin,)
    out[f"{prefix}.conv1.weight"] = (C, cin, 3, 3)
    out[f"{prefix}.conv1.bias"] = (C,)
    out[f"{prefix}.time_emb_proj.weight"] = (C, temb)
    out[f"{prefix}.time_emb_proj.bias"] = (C,)
    out[f"{prefix}.norm2.weight"] = (C,)
    out[f"{prefix}.norm2.bias"] = (C,)
    out[f"{prefix}.conv2.weight"] = (C, C, 3, 3)
    out[f"{prefix}.conv2.bias"] = (C,)
    if cin != C:
        out[f"{prefix}.conv_shortcut.weight"] = (C, cin, 1, 1)
        out[f"{prefix}.conv_shortcut.bias"] = (C,)


def _temp_conv(prefix: str, C: int, out: Dict):
    for i, conv_idx in ((1, 2), (2, 3), (3, 3), (4, 3)):
        out[f"{prefix}.conv{i}.0.weight"] = (C,)
        out[f"{prefix}.conv{i}.0.bias"] = (C,)
        out[f"{prefix}.conv{i}.{conv_idx}.weight"] = (C, C, 3, 1, 1)
        out[f"{prefix}.conv{i}.{conv_idx}.bias"] = (C,)
    out[f"{prefix}.temporal_weight"] = (1,)


def _spatial_tfm(prefix: str, C: int, cfg: UNetConfig, out: Dict):
    out[f"{prefix}.norm.weight"] = (C,)
    out[f"{prefix}.norm.bias"] = (C,)
    out[f"{prefix}.proj_in.weight"] = (C, C, 1, 1)
    out[f"{prefix}.proj_in.bias"] = (C,)
    _tblock(f"{prefix}.transformer_blocks.0", C, cfg.cross_attention_dim, out, ip=cfg.ip_adapter_cross_attn)
    out[f"{prefix}.proj_out.weight"] = (C, C, 1, 1)
    out[f"{prefix}.proj_out.bias"] = (C,)


def _temporal_tfm(prefix: str, C: int, cfg: UNetConfig, out: Dict):
    out[f"{prefix}.temporal_weight"] = (1,)
    out[f"{prefix}.norm.weight"] = (C,)
    out[f"{prefix}.norm.bias"] = (C,)
    out[f"{prefix}.proj_in.weight"] = (C, C)
    out[f"{prefix}.proj_in.bias"] = (C,)
    out[f"{prefix}.frame_emb_proj.weight"] = (C, cfg.temb_dim)
    out[f"{prefix}.frame_emb_proj.bias"] = (C,)
    _tblock(f"{prefix}.transformer_blocks.0", C, None, out, ip=False)
    out[f"{prefix}.proj_out.weight"] = (C, C)
    out[f"{prefix}.proj_out.bias"] = (C,)


def _refer_attn(prefix: str, C: int, out: Dict):
    _attn(prefix, C, C, out)


def unet_param_shapes(cfg: UNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape for every tensor of the reference `state_dict()` (order is not significant)."""
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = cfg.block_out_channels
    c0, temb = boc[0], cfg.temb_dim
    out["conv_in.weight"] = (c0, cfg.in_channels, 3, 3)
    out["conv_in.bias"] = (c0,)
    for emb in ("time_embedding", "frame_embedding"):
        out[f"{emb}.linear_1.weight"] = (temb, c0)
        out[f"{emb}.linear_1.bias"] = (temb,)
        out[f"{emb}.linear_2.weight"] = (temb, temb)
        out[f"{emb}.linear_2.bias"] = (temb,)
    if cfg.need_transformer_in:
        _temporal_tfm("transformer_in", c0, cfg, out)
    if cfg.need_refer_emb:
        _refer_attn("first_refer_emb_attns", c0, out)
        _refer_attn("mid_block_refer_emb_attns", boc[-1], out)
    nb = len(boc)
    # down
    ch = c0
    for i in range(nb):
        cin, ch = ch, boc[i]
        final = i == nb - 1
        has_attn = not final
        for j in range(cfg.layers_per_block):
            _resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else ch, ch, temb, out)
            _temp_conv(f"down_blocks.{i}.temp_convs.{j}", ch, out)
            if has_attn:
                _spatial_tfm(f"down_blocks.{i}.attentions.{j}", ch, cfg, out)
                _temporal_tfm(f"down_blocks.{i}.temp_attentions.{j}", ch, cfg, out)
        if not final:
            out[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (ch, ch, 3, 3)
            out[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (ch,)
        if cfg.need_refer_emb:
            for k in range(cfg.layers_per_block + (0 if final else 1)):
                _refer_attn(f"down_blocks.{i}.refer_emb_attns.{k}", ch, out)
    # mid
    cm = boc[-1]
    _resnet("mid_block.resnets.0", cm, cm, temb, out)
    _temp_conv("mid_block.temp_convs.0", cm, out)
    _spatial_tfm("mid_block.attentions.0", cm, cfg, out)
    _temporal_tfm("mid_block.temp_attentions.0", cm, cfg, out)
    _resnet("mid_block.resnets.1", cm, cm, temb, out)
    _temp_conv("mid_block.temp_convs.1", cm, out)
    # up
    rev = list(reversed(boc))
    ch = rev[0]
    for i in range(nb):
        prev, ch = ch, rev[i]
        cin_block = rev[min(i + 1, nb - 1)]
        has_attn = i > 0
        final = i == nb - 1
        for j in range(cfg.layers_per_block + 1):
            skip = cin_block if j == cfg.layers_per_block else ch
            rin = prev if j == 0 else ch
            _resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, ch, temb, out)
            _temp_conv(f"up_blocks.{i}.temp_convs.{j}", ch, out)
            if has_attn:
                _spatial_tfm(f"up_blocks.{i}.attentions.{j}", ch, cfg, out)
                _temporal_tfm(f"up_blocks.{i}.temp_attentions.{j}", ch, cfg, out)
        if not final:
            out[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (ch, ch, 3, 3)
            out[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (ch,)
    out["conv_norm_out.weight"] = (c0,)
    out["conv_norm_out.bias"] = (c0,)
    out["conv_out.weight"] = (cfg.out_channels, c0, 3, 3)
    out["conv_out.bias"] = (cfg.out_channels,)
    return out


def refer_emb_shapes(cfg: UNetConfig, h: int, w: int):
    """Shapes [C, h, w] of the 12 down-block ReferenceNet maps + the mid map the UNet consumes
    (musev/models/referencenet.py:1116-1127; consumed at unet_3d_condition.py:1052-1095,1176-1187)."""
    boc = cfg.block_out_channels
    shapes = [(boc[0], h, w)]
    hh, ww = h, w
    for i, c in enumerate(boc):
        final = i == len(boc) - 1
        for _ in range(cfg.layers_per_block):
            shapes.append((c, hh, ww))
        if not final:
            hh, ww = hh // 2, ww // 2
            shapes.append((c, hh, ww))
    mid = (boc[-1], hh, ww)
    return shapes, mid


# ------------------------------------------------------------------------------------------------ ControlNet (8f-1)
@dataclass
class ControlNetConfig:
    """diffusers `ControlNetModel.__init__` defaults for SD-1.5 ControlNets (models/controlnet.py:181-262)."""
    in_channels: int = 4
    conditioning_channels: int = 3
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: int = 8          # number of heads, as in UNetConfig
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    conditioning_embedding_out_channels: Tuple[int, ...] = (16, 32, 96, 256)
    resnet_2d_skip_time_act: bool = False   # lets the oracle reuse the UNet's ResnetBlock2D restatement

    @property
    def heads(self) -> int:
        return self.attention_head_dim

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * 4


def _vanilla_tfm(prefix: str, C: int, cross_dim: int, out: Dict):
    out[f"{prefix}.norm.weight"] = (C,)
    out[f"{prefix}.norm.bias"] = (C,)
    out[f"{prefix}.proj_in.weight"] = (C, C, 1, 1)
    out[f"{prefix}.proj_in.bias"] = (C,)
    _tblock(f"{prefix}.transformer_blocks.0", C, cross_dim, out, ip=False)
    out[f"{prefix}.proj_out.weight"] = (C, C, 1, 1)
    out[f"{prefix}.proj_out.bias"] = (C,)


def controlnet_param_shapes(cfg: ControlNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of the reference `ControlNetModel.state_dict()` (diffusers models/controlnet.py:181-447):
    SD-1.5 encoder half (3 x CrossAttnDownBlock2D + DownBlock2D + UNetMidBlock2DCrossAttn), the conditioning
    embedding and the 12 + 1 zero convolutions."""
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = cfg.block_out_channels
    c0, temb = boc[0], cfg.temb_dim
    out["conv_in.weight"] = (c0, cfg.in_channels, 3, 3)
    out["conv_in.bias"] = (c0,)
    out["time_embedding.linear_1.weight"] = (temb, c0)
    out["time_embedding.linear_1.bias"] = (temb,)
    out["time_embedding.linear_2.weight"] = (temb, temb)
    out["time_embedding.linear_2.bias"] = (temb,)
    ce = cfg.conditioning_embedding_out_channels
    out["controlnet_cond_embedding.conv_in.weight"] = (ce[0], cfg.conditioning_channels, 3, 3)
    out["controlnet_cond_embedding.conv_in.bias"] = (ce[0],)
    for i in range(len(ce) - 1):
        out[f"controlnet_cond_embedding.blocks.{2 * i}.weight"] = (ce[i], ce[i], 3, 3)
        out[f"controlnet_cond_embedding.blocks.{2 * i}.bias"] = (ce[i],)
        out[f"controlnet_cond_embedding.blocks.{2 * i + 1}.weight"] = (ce[i + 1], ce[i], 3, 3)
        out[f"controlnet_cond_embedding.blocks.{2 * i + 1}.bias"] = (ce[i + 1],)
    out["controlnet_cond_embedding.conv_out.weight"] = (c0, ce[-1], 3, 3)
    out["controlnet_cond_embedding.conv_out.bias"] = (c0,)
    nb = len(boc)
    taps = [c0]                     # channels of the 12 residual taps (conv_in output first)
    ch = c0
    for i in range(nb):
        cin, ch = ch, boc[i]
        final = i == nb - 1
        for j in range(cfg.layers_per_block):
            _resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else ch, ch, temb, out)
            if not final:
                _vanilla_tfm(f"down_blocks.{i}.attentions.{j}", ch, cfg.cross_attention_dim, out)
            taps.append(ch)
        if not final:
            out[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (ch, ch, 3, 3)
            out[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (ch,)
            taps.append(ch)
    cm = boc[-1]
    _resnet("mid_block.resnets.0", cm, cm, temb, out)
    _vanilla_tfm("mid_block.attentions.0", cm, cfg.cross_attention_dim, out)
    _resnet("mid_block.resnets.1", cm, cm, temb, out)
    for k, c in enumerate(taps):
        out[f"controlnet_down_blocks.{k}.weight"] = (c, c, 1, 1)
        out[f"controlnet_down_blocks.{k}.bias"] = (c,)
    out["controlnet_mid_block.weight"] = (cm, cm, 1, 1)
    out["controlnet_mid_block.bias"] = (cm,)
    return out


# ------------------------------------------------------------------------------------------------ ReferenceNet (8f-2)
@dataclass
class ReferenceNetConfig(ControlNetConfig):
    """`ReferenceNet2D` as `load_referencenet_by_name("musev_referencenet")` builds it (musev/models/referencenet_loader.py
    :109-118: need_block_embs=True, need_self_attn_block_embs=False) from an SD-1.5 `unet/config.json`: the encoder half +
    mid block of the 2-D UNet; conv_norm_out / conv_out / up_blocks are set to None (referencenet.py:624-636)."""


def referencenet_param_shapes(cfg: ReferenceNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of the reference `ReferenceNet2D.state_dict()` (296 tensors at SD-1.5 size): conv_in, time_embedding,
    down_blocks (3 x CrossAttnDownBlock2D + DownBlock2D), mid_block (musev/models/referencenet.py:213-636)."""
    full = controlnet_param_shapes(cfg)
    return OrderedDict((k, v) for k, v in full.items()
                       if not (k.startswith("controlnet_cond_embedding.") or k.startswith("controlnet_down_blocks.")
                               or k.startswith("controlnet_mid_block.")))


@dataclass
class ImageProjConfig:
    """`ImageProjModel` of the IP-Adapter package (ip_adapter/ip_adapter.py, tencent-ailab/IP-Adapter@main -- a pip
    dependency of the reference, requirements.txt:2, not vendored) with the arguments the reference passes
    (musev/models/ip_adapter_loader.py:89-93)."""
    cross_attention_dim: int = 768
    clip_embeddings_dim: int = 1024
    clip_extra_context_tokens: int = 4


def image_proj_param_shapes(cfg: ImageProjConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    out["proj.weight"] = (cfg.clip_extra_context_tokens * cfg.cross_attention_dim, cfg.clip_embeddings_dim)
    out["proj.bias"] = (cfg.clip_extra_context_tokens * cfg.cross_attention_dim,)
    out["norm.weight"] = (cfg.cross_attention_dim,)
    out["norm.bias"] = (cfg.cross_attention_dim,)
    return out


# ------------------------------------------------------------------------------------------------ VAE decoder (8f-3)
@dataclass
class VAEConfig:
    """SD-1.5 `vae/config.json` as `AutoencoderKL.__init__` takes it (diffusers models/autoencoder_kl.py:65-118); only the
    decoder half (+ post_quant_conv) is built."""
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215


def vae_decoder_param_shapes(cfg: VAEConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of the `post_quant_conv.*` and `decoder.*` entries of the reference `AutoencoderKL.state_dict()`
    (diffusers models/vae.py:201-263; 138 tensors at SD-1.5 size)."""
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = cfg.block_out_channels
    zc, cm = cfg.latent_channels, boc[-1]

    def resnet(p, cin, c):
        out[f"{p}.norm1.weight"] = (cin,)
        out[f"{p}.norm1.bias"] = (cin,)
        out[f"{p}.conv1.weight"] = (c, cin, 3, 3)
        out[f"{p}.conv1.bias"] = (c,)
        out[f"{p}.norm2.weight"] = (c,)
        out[f"{p}.norm2.bias"] = (c,)
        out[f"{p}.conv2.weight"] = (c, c, 3, 3)
        out[f"{p}.conv2.bias"] = (c,)
        if cin != c:
            out[f"{p}.conv_shortcut.weight"] = (c, cin, 1, 1)
            out[f"{p}.conv_shortcut.bias"] = (c,)

    out["post_quant_conv.weight"] = (zc, zc, 1, 1)
    out["post_quant_conv.bias"] = (zc,)
    out["decoder.conv_in.weight"] = (cm, zc, 3, 3)
    out["decoder.conv_in.bias"] = (cm,)
    resnet("decoder.mid_block.resnets.0", cm, cm)
    a = "decoder.mid_block.attentions.0"
    out[f"{a}.group_norm.weight"] = (cm,)
    out[f"{a}.group_norm.bias"] = (cm,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out[f"{a}.{n}.weight"] = (cm, cm)
        out[f"{a}.{n}.bias"] = (cm,)
    resnet("decoder.mid_block.resnets.1", cm, cm)
    ch = cm
    nb = len(boc)
    for i in range(nb):
        prev, ch = ch, boc[nb - 1 - i]
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else ch, ch)
        if i != nb - 1:
            out[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (ch, ch, 3, 3)
            out[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (ch,)
    out["decoder.conv_norm_out.weight"] = (boc[0],)
    out["decoder.conv_norm_out.bias"] = (boc[0],)
    out["decoder.conv_out.weight"] = (cfg.out_channels, boc[0], 3, 3)
    out["decoder.conv_out.bias"] = (cfg.out_channels,)
    return out
