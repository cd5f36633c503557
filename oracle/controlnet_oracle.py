"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the ControlNet encoder MuseV runs per window-step.

Follows diffusers `ControlNetModel.forward` (diffusers/src/diffusers/models/controlnet.py:645-852) as it is called by
`MusevControlNetPipeline.get_controlnet_emb` (musev/pipelines/pipeline_controlnet.py:1202-1291): frames flattened to
the batch axis, prompt embedding repeated per frame, optional pre-computed condition embedding
(`controlnet_cond_latents`). The blocks are the vanilla diffusers ones (unet_2d_blocks.py:1115-1180 CrossAttnDownBlock2D,
:630-760 UNetMidBlock2DCrossAttn; transformer_2d.py:257-389; attention.py:71-340 with all three LayerNorm eps = 1e-5, unlike
the UNet's Q1), ResnetBlock2D / Downsample2D as in the UNet oracle. Pinned against the unmodified reference class by
oracle/make_golden.py -> tests/golden/controlnet_*.pt. Not imported by the product path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from oracle.unet3d_oracle import UNet3DOracle, timestep_embedding


class ControlNetOracle(UNet3DOracle):
    """Shares the small helpers (_linear/_gn/_ln/_sdpa/resnet/feed_forward) with the UNet oracle."""

    def vanilla_block(self, x, enc, p):
        """diffusers BasicTransformerBlock.forward (models/attention.py:240-340): LN -> self-attn -> LN -> text
        cross-attn -> LN -> GEGLU feed-forward, residual after each; LayerNorm eps 1e-5 throughout."""
        n = self._ln(x, p + ".norm1", 1e-5)
        a = self._sdpa(self._linear(n, p + ".attn1.to_q", False), self._linear(n, p + ".attn1.to_k", False),
                       self._linear(n, p + ".attn1.to_v", False))
        x = self._linear(a, p + ".attn1.to_out.0") + x
        n = self._ln(x, p + ".norm2", 1e-5)
        a = self._sdpa(self._linear(n, p + ".attn2.to_q", False), self._linear(enc, p + ".attn2.to_k", False),
                       self._linear(enc, p + ".attn2.to_v", False))
        x = self._linear(a, p + ".attn2.to_out.0") + x
        n = self._ln(x, p + ".norm3", 1e-5)
        return self.feed_forward(n, p) + x

    def vanilla_transformer(self, x, enc, p):
        """diffusers Transformer2DModel.forward, continuous conv-projection path (models/transformer_2d.py:257-389)."""
        bt, c, hh, ww = x.shape
        h = self._gn(x, p + ".norm", 1e-6)
        h = F.conv2d(h, self._w(p + ".proj_in.weight"), self._w(p + ".proj_in.bias"))
        h = h.permute(0, 2, 3, 1).reshape(bt, hh * ww, c)
        h = self.vanilla_block(h, enc, p + ".transformer_blocks.0")
        h = h.reshape(bt, hh, ww, c).permute(0, 3, 1, 2)
        h = F.conv2d(h, self._w(p + ".proj_out.weight"), self._w(p + ".proj_out.bias"))
        return h + x

    def cond_embedding(self, cond):
        """ControlNetConditioningEmbedding.forward (models/controlnet.py:101-112): conv 3x3 + SiLU, then
        (conv 3x3, conv 3x3 stride 2) x 3 each followed by SiLU, then the zero-initialised conv_out."""
        p = "controlnet_cond_embedding"
        e = F.silu(F.conv2d(cond, self._w(p + ".conv_in.weight"), self._w(p + ".conv_in.bias"), padding=1))
        n_blocks = 2 * (len(self.cfg.conditioning_embedding_out_channels) - 1)
        for i in range(n_blocks):
            e = F.silu(F.conv2d(e, self._w(f"{p}.blocks.{i}.weight"), self._w(f"{p}.blocks.{i}.bias"), padding=1,
                                stride=2 if i % 2 else 1))
        return F.conv2d(e, self._w(p + ".conv_out.weight"), self._w(p + ".conv_out.bias"), padding=1)

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond=None, conditioning_scale=1.0,
                guess_mode=False, controlnet_cond_latents=None) -> Tuple[List[torch.Tensor], torch.Tensor]:
        """ControlNetModel.forward, models/controlnet.py:645-852 (no class / addition embeddings for SD-1.5 nets,
        `global_pool_conditions` False). Returns (12 down residuals, mid residual), already scaled."""
        cfg = self.cfg
        dev, dt = self.device, self.dtype
        x = sample.to(dev, dt)
        enc = encoder_hidden_states.to(dev, dt)
        t = torch.as_tensor(timestep, device=dev).reshape(-1).expand(x.shape[0])
        emb = self._mlp_emb(timestep_embedding(t, cfg.block_out_channels[0]).to(dt), "time_embedding")   # :733-741
        x = F.conv2d(x, self._w("conv_in.weight"), self._w("conv_in.bias"), padding=1)                     # :780
        if controlnet_cond_latents is None:
            cond = self.cond_embedding(controlnet_cond.to(dev, dt))                                         # :781-782
        else:
            cond = controlnet_cond_latents.to(dev, dt)
        x = x + cond                                                                                        # :785
        taps = [x]
        nb = len(cfg.block_out_channels)
        for i in range(nb):                                                                                 # :788-801
            final = i == nb - 1
            for j in range(cfg.layers_per_block):
                x = self.resnet(x, emb, f"down_blocks.{i}.resnets.{j}")
                if not final:
                    x = self.vanilla_transformer(x, enc, f"down_blocks.{i}.attentions.{j}")
                taps.append(x)
            if not final:
                x = F.conv2d(x, self._w(f"down_blocks.{i}.downsamplers.0.conv.weight"),
                             self._w(f"down_blocks.{i}.downsamplers.0.conv.bias"), stride=2, padding=1)
                taps.append(x)
        x = self.resnet(x, emb, "mid_block.resnets.0")                                                     # :804-811
        x = self.vanilla_transformer(x, enc, "mid_block.attentions.0")
        x = self.resnet(x, emb, "mid_block.resnets.1")
        down = [F.conv2d(tp, self._w(f"controlnet_down_blocks.{k}.weight"), self._w(f"controlnet_down_blocks.{k}.bias"))
                for k, tp in enumerate(taps)]                                                               # :815-821
        mid = F.conv2d(x, self._w("controlnet_mid_block.weight"), self._w("controlnet_mid_block.bias"))   # :823
        if guess_mode:                                                                                      # :826-830
            scales = torch.logspace(-1, 0, len(down) + 1, device=dev).to(dt) * conditioning_scale
            down = [d * s for d, s in zip(down, scales)]
            mid = mid * scales[-1]
        else:                                                                                               # :831-833
            down = [d * conditioning_scale for d in down]
            mid = mid * conditioning_scale
        return down, mid

    __call__ = forward
