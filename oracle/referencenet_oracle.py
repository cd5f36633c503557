"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two one-shot side paths of SURVEY.md 8(a15) / 8(f)-2.

  * ReferenceNetOracle   `ReferenceNet2D.forward` (musev/models/referencenet.py:640-1127) as `get_referencenet_emb` calls it
                         (musev/pipelines/pipeline_controlnet.py:867-964): SD-1.5 encoder half + mid block built from
                         musev/models/unet_2d_blocks.py (CrossAttnDownBlock2D :1107-1240, DownBlock2D :1243-1340,
                         UNetMidBlock2DCrossAttn :627-770) whose attentions are musev `Transformer2DModel` /
                         `BasicTransformerBlock` -- LayerNorm eps 0 / 1e-5 / 0 -- with the default
                         `AttnProcessor2_0` (plain self attention, text cross attention). Returns the 12 down-block maps and
                         the mid map reshaped `(b t) c h w -> b c t h w` (:1041-1049,1063-1127). Pinned against the
                         unmodified reference class by oracle/make_golden.py -> tests/golden/referencenet_*.pt.
  * image_proj_oracle    `ImageProjModel.forward` of the IP-Adapter package (ip_adapter/ip_adapter.py,
                         tencent-ailab/IP-Adapter@main; requirements.txt:2). The package is NOT vendored under
                         /root/reference, so this restates its published algorithm -- Linear(clip_dim -> tokens * cross_dim),
                         reshape to [-1, tokens, cross_dim], LayerNorm(cross_dim, eps 1e-5) -- and is anchored on the
                         reference's call sites (musev/models/ip_adapter_loader.py:89-93;
                         musev/pipelines/pipeline_controlnet.py:725-770 incl. the uncond branch = proj(zeros)).
                         PARITY UNPINNED for this function (no reference code or golden vector to execute).
Not imported by the product path.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from oracle.unet3d_oracle import UNet3DOracle, timestep_embedding


class ReferenceNetOracle(UNet3DOracle):
    """Shares _linear/_gn/_ln/_sdpa/resnet/feed_forward with the UNet oracle."""

    def musev_block_2d(self, x, enc, p):
        """musev BasicTransformerBlock.forward (musev/models/attention.py:172-431) in a 2-D net: no reference-only K/V
        extension (the processor is the default AttnProcessor2_0, diffusers attention_processor.py:1166-1250), no IP
        branch; eps 0 / 1e-5 / 0."""
        n = self._ln(x, p + ".norm1", 0.0)
        a = self._sdpa(self._linear(n, p + ".attn1.to_q", False), self._linear(n, p + ".attn1.to_k", False),
                       self._linear(n, p + ".attn1.to_v", False))
        x = self._linear(a, p + ".attn1.to_out.0") + x
        n = self._ln(x, p + ".norm2", 1e-5)
        a = self._sdpa(self._linear(n, p + ".attn2.to_q", False), self._linear(enc, p + ".attn2.to_k", False),
                       self._linear(enc, p + ".attn2.to_v", False))
        x = self._linear(a, p + ".attn2.to_out.0") + x
        n = self._ln(x, p + ".norm3", 0.0)
        return self.feed_forward(n, p) + x

    def transformer_2d(self, x, enc, p):
        """musev Transformer2DModel.forward continuous path (musev/models/transformer_2d.py:257-276,313-389)."""
        bt, c, hh, ww = x.shape
        h = self._gn(x, p + ".norm", 1e-6)
        h = F.conv2d(h, self._w(p + ".proj_in.weight"), self._w(p + ".proj_in.bias"))
        h = h.permute(0, 2, 3, 1).reshape(bt, hh * ww, c)
        h = self.musev_block_2d(h, enc, p + ".transformer_blocks.0")
        h = h.reshape(bt, hh, ww, c).permute(0, 3, 1, 2)
        h = F.conv2d(h, self._w(p + ".proj_out.weight"), self._w(p + ".proj_out.bias"))
        return h + x

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, num_frames=1, return_ndim=5) -> Tuple[List[torch.Tensor], torch.Tensor]:
        cfg = self.cfg
        dev, dt = self.device, self.dtype
        x = sample.to(dev, dt)
        enc = encoder_hidden_states.to(dev, dt)
        t = torch.as_tensor(timestep, device=dev).reshape(-1).expand(x.shape[0])                           # :880-901
        emb = self._mlp_emb(timestep_embedding(t, cfg.block_out_channels[0]).to(dt), "time_embedding")    # :903-910
        x = F.conv2d(x, self._w("conv_in.weight"), self._w("conv_in.bias"), padding=1)                     # :1004
        taps = [x]
        nb = len(cfg.block_out_channels)
        for i in range(nb):                                                                                 # :1021-1061
            final = i == nb - 1
            for j in range(cfg.layers_per_block):
                x = self.resnet(x, emb, f"down_blocks.{i}.resnets.{j}")
                if not final:
                    x = self.transformer_2d(x, enc, f"down_blocks.{i}.attentions.{j}")
                taps.append(x)
            if not final:
                x = F.conv2d(x, self._w(f"down_blocks.{i}.downsamplers.0.conv.weight"),
                             self._w(f"down_blocks.{i}.downsamplers.0.conv.bias"), stride=2, padding=1)
                taps.append(x)
        x = self.resnet(x, emb, "mid_block.resnets.0")                                                     # :1078-1092
        x = self.transformer_2d(x, enc, "mid_block.attentions.0")
        x = self.resnet(x, emb, "mid_block.resnets.1")

        def reshape(e):                                                                                     # :1041-1049
            if return_ndim == 4:
                return e
            bt, c, hh, ww = e.shape
            return e.view(bt // num_frames, num_frames, c, hh, ww).permute(0, 2, 1, 3, 4).contiguous()
        return [reshape(e) for e in taps], reshape(x)

    __call__ = forward


def image_proj_oracle(sd: Dict[str, torch.Tensor], image_embeds: torch.Tensor, tokens: int = 4, cross_dim: int = 768) -> torch.Tensor:
    """ImageProjModel.forward (published IP-Adapter algorithm, see the module docstring): [N, clip_dim] or [N, 1, clip_dim]
    -> [N, tokens, cross_dim]."""
    e = image_embeds.float()
    y = F.linear(e, sd["proj.weight"].float(), sd["proj.bias"].float()).reshape(-1, tokens, cross_dim)
    return F.layer_norm(y, (cross_dim,), sd["norm.weight"].float(), sd["norm.bias"].float(), 1e-5)
