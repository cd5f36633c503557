"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the VAE decode that follows the denoise loop.

`AutoencoderKL.decode` = post_quant_conv + `Decoder.forward` (diffusers models/autoencoder_kl.py:275-302, models/vae.py:265-316):
conv_in -> UNetMidBlock2D (unet_2d_blocks.py: ResnetBlock2D without time embedding, eps 1e-6; `Attention` with ONE head of
dim C, GroupNorm eps 1e-6, biased q/k/v/out, residual connection, attention_processor.py:1166-1250) -> 4 x UpDecoderBlock2D
(3 resnets + nearest-2x upsample + 3x3 conv) -> GroupNorm(eps 1e-6) + SiLU + conv_out; and the pipeline-level
`decode_latents` (pipeline_stable_diffusion_img2img.py:486-495 via musev/pipelines/pipeline_controlnet.py:233-238).
Pinned against the unmodified diffusers `AutoencoderKL` by oracle/make_golden.py -> tests/golden/vae_*.pt.
Not imported by the product path.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


class VAEDecoderOracle:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], device="cpu", dtype=torch.float32):
        self.cfg = cfg
        self.sd = {k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}
        self.device, self.dtype = device, dtype

    def _gn(self, x, p):
        return F.group_norm(x, self.cfg.norm_num_groups, self.sd[p + ".weight"], self.sd[p + ".bias"], 1e-6)

    def resnet(self, x, p):
        """ResnetBlock2D.forward without temb (diffusers models/resnet.py:696-770; resnet_eps=1e-6, output_scale_factor=1)."""
        h = F.conv2d(F.silu(self._gn(x, p + ".norm1")), self.sd[p + ".conv1.weight"], self.sd[p + ".conv1.bias"], padding=1)
        h = F.conv2d(F.silu(self._gn(h, p + ".norm2")), self.sd[p + ".conv2.weight"], self.sd[p + ".conv2.bias"], padding=1)
        if (p + ".conv_shortcut.weight") in self.sd:
            x = F.conv2d(x, self.sd[p + ".conv_shortcut.weight"], self.sd[p + ".conv_shortcut.bias"])
        return x + h

    def attention(self, x, p):
        n, c, hh, ww = x.shape
        t = self._gn(x.view(n, c, hh * ww), p + ".group_norm").transpose(1, 2)          # [n, hw, c]
        q = F.linear(t, self.sd[p + ".to_q.weight"], self.sd[p + ".to_q.bias"])
        k = F.linear(t, self.sd[p + ".to_k.weight"], self.sd[p + ".to_k.bias"])
        v = F.linear(t, self.sd[p + ".to_v.weight"], self.sd[p + ".to_v.bias"])
        a = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]        # one head, scale c ** -0.5
        a = F.linear(a, self.sd[p + ".to_out.0.weight"], self.sd[p + ".to_out.0.bias"])
        return a.transpose(1, 2).reshape(n, c, hh, ww) + x

    @torch.no_grad()
    def decode(self, z):
        cfg = self.cfg
        z = z.to(self.device, self.dtype)
        z = F.conv2d(z, self.sd["post_quant_conv.weight"], self.sd["post_quant_conv.bias"])           # autoencoder_kl.py:283
        x = F.conv2d(z, self.sd["decoder.conv_in.weight"], self.sd["decoder.conv_in.bias"], padding=1)
        x = self.resnet(x, "decoder.mid_block.resnets.0")
        x = self.attention(x, "decoder.mid_block.attentions.0")
        x = self.resnet(x, "decoder.mid_block.resnets.1")
        nb = len(cfg.block_out_channels)
        for i in range(nb):
            for j in range(cfg.layers_per_block + 1):
                x = self.resnet(x, f"decoder.up_blocks.{i}.resnets.{j}")
            if i != nb - 1:
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                x = F.conv2d(x, self.sd[p + ".weight"], self.sd[p + ".bias"], padding=1)
        x = F.silu(self._gn(x, "decoder.conv_norm_out"))
        return F.conv2d(x, self.sd["decoder.conv_out.weight"], self.sd["decoder.conv_out.bias"], padding=1)

    @torch.no_grad()
    def decode_latents(self, latents):
        b, c, f, h, w = latents.shape
        z = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w) / self.cfg.scaling_factor
        img = (self.decode(z) / 2 + 0.5).clamp(0, 1)
        return img.view(b, f, *img.shape[1:]).permute(0, 2, 1, 3, 4).contiguous()
