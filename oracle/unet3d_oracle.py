"""TEST INFRASTRUCTURE ONLY -- plain-torch fp32 restatement of the reference denoiser
(`musev.models.unet_3d_condition.UNet3DConditionModel.forward`) driven by a reference-format state_dict.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this file,
and only as the checker / reported baseline. The product (musev_b200/) never does.

PINNING: this restatement is checked against the *imported, unmodified* reference on identical weights and inputs
(tests/test_oracle_pinned.py uses fixtures produced by oracle/make_golden.py, which runs
/root/reference/musev/... through oracle/ref_shim.py). The reference itself ships no test for this model
(SURVEY.md section 4), so parity is pinned by executing the reference code, not by upstream golden vectors;
the diffusers building blocks it uses (ResnetBlock2D, DDIM) do have upstream KATs, replayed in
tests/test_oracle_pinned.py as well.

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers/src/diffusers/models/embeddings.py:26-66 with flip_sin_to_cos=True, downscale_freq_shift=0
    (Timesteps(320, True, 0), musev/models/unet_3d_condition.py:343,356)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class UNet3DOracle:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], device="cpu", dtype=torch.float32, emulate=None):
        """emulate: None (plain fp32 restatement, the pinned oracle) | "all" | "branch". The two emulation modes round
        intermediate tensors to fp16 at the points where the CUDA engine stores them in fp16 (GEMM / norm / attention
        outputs); "branch" leaves the residual stream (the running hidden state every layer adds into) unrounded. They
        exist to attribute the engine's distance from the fp32 oracle to its storage precision (tools/gpu_parity_20step.py)
        and are never used as the parity reference."""
        self.emulate = emulate
        self.cfg = cfg
        self.sd = {k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}
        self.device = device
        self.dtype = dtype
        self.taps = None  # optional dict name -> tensor of intermediate activations ((b t) c h w), for bisecting

    # ------------------------------------------------------------------ small helpers
    def _w(self, name):
        return self.sd[name]

    def _q(self, x):
        """branch tensor as the engine stores it (fp16) -- identity unless emulating"""
        return x.half().to(x.dtype) if self.emulate else x

    def _qs(self, x):
        """residual-stream tensor: rounded only in emulate="all" """
        return x.half().to(x.dtype) if self.emulate == "all" else x

    def _tap(self, name, x):
        if self.taps is not None:
            self.taps[name] = x.detach().clone()

    def _linear(self, x, prefix, bias=True):
        return F.linear(x, self._w(prefix + ".weight"), self.sd.get(prefix + ".bias") if bias else None)

    def _gn(self, x, prefix, eps):
        return self._q(F.group_norm(x, self.cfg.norm_num_groups, self._w(prefix + ".weight"), self._w(prefix + ".bias"), eps))

    def _ln(self, x, prefix, eps):
        return self._q(F.layer_norm(x, (x.shape[-1],), self._w(prefix + ".weight"), self._w(prefix + ".bias"), eps))

    def _mlp_emb(self, x, prefix):
        """TimestepEmbedding: diffusers embeddings.py:190-253 (linear_1, SiLU, linear_2)."""
        return self._linear(F.silu(self._linear(x, prefix + ".linear_1")), prefix + ".linear_2")

    # ------------------------------------------------------------------ attention core
    def _heads(self, x):
        b, n, c = x.shape
        h = self.cfg.heads
        return x.view(b, n, h, c // h).permute(0, 2, 1, 3)

    def _sdpa(self, q, k, v):
        """softmax(q k^T / sqrt(d)) v -- what xformers.ops.memory_efficient_attention / SDPA compute
        (musev/models/attention_processor.py:258,292,519,724; diffusers attention_processor.py:1166-1250);
        scale = dim_head ** -0.5 (diffusers attention_processor.py:127)."""
        q, k, v = self._heads(q), self._heads(k), self._heads(v)
        o = F.scaled_dot_product_attention(self._q(q), self._q(k), self._q(v))
        b, h, n, d = o.shape
        return self._q(o.permute(0, 2, 1, 3).reshape(b, n, h * d))

    # ------------------------------------------------------------------ blocks
    def resnet(self, x, temb, p):
        """ResnetBlock2D.forward, diffusers models/resnet.py:696-770 (time_embedding_norm='default',
        output_scale_factor=1, eps=norm_eps)."""
        cfg = self.cfg
        h = self._q(F.silu(self._gn(x, p + ".norm1", cfg.norm_eps)))
        h = F.conv2d(h, self._w(p + ".conv1.weight"), self._w(p + ".conv1.bias"), padding=1)
        t = temb if cfg.resnet_2d_skip_time_act else F.silu(temb)
        h = self._q(h + self._linear(self._q(t), p + ".time_emb_proj")[:, :, None, None])
        h = self._q(F.silu(self._gn(h, p + ".norm2", cfg.norm_eps)))
        h = F.conv2d(h, self._w(p + ".conv2.weight"), self._w(p + ".conv2.bias"), padding=1)
        if (p + ".conv_shortcut.weight") in self.sd:
            x = self._q(F.conv2d(self._q(x), self._w(p + ".conv_shortcut.weight"), self._w(p + ".conv_shortcut.bias")))
        return self._qs(x + h)

    def temp_conv(self, x, T, p):
        """TemporalConvLayer.forward, musev/models/resnet.py:95-135: 4 x [GroupNorm over (c/g, t, h, w) -> SiLU ->
        Conv3d (3,1,1)], then identity + |temporal_weight| * x. `femb` is unused (Q6)."""
        if self.skip_temporal:
            return x
        bt, c, hh, ww = x.shape
        v = x.view(bt // T, T, c, hh, ww).permute(0, 2, 1, 3, 4)
        identity = v
        for i, ci in ((1, 2), (2, 3), (3, 3), (4, 3)):
            v = self._q(F.silu(self._gn(v, f"{p}.conv{i}.0", 1e-5)))
            v = F.conv3d(v, self._w(f"{p}.conv{i}.{ci}.weight"), self._w(f"{p}.conv{i}.{ci}.bias"), padding=(1, 0, 0))
            if i < 4:
                v = self._q(v)
        v = self._qs(identity + torch.abs(self._w(p + ".temporal_weight")) * v)
        return v.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)

    def feed_forward(self, x, p):
        """FeedForward with GEGLU, diffusers models/attention.py:342-395, activations.py:89-102 (erf GELU)."""
        h = self._linear(x, p + ".ff.net.0.proj")
        val, gate = h.chunk(2, dim=-1)
        return self._linear(self._q(val * F.gelu(gate)), p + ".ff.net.2")

    def spatial_block(self, x, enc, T, p, vis_idx, clip_emb, ip_scale):
        """musev BasicTransformerBlock.forward (musev/models/attention.py:172-431) for a spatial layer.
        LayerNorm eps: norm1 = norm3 = 0, norm2 = 1e-5 (Q1). attn1 = reference-only self attention
        (NonParamT2ISelfReferenceXFormersAttnProcessor, attention_processor.py:378-546): keys/values of each frame are
        its own tokens followed by the tokens of the vision-condition frame(s). attn2 = text cross-attention, plus
        ip_scale * image cross-attention when the block has to_k_ip (attention_processor.py:176-359).
        The CFG recompute at attention.py:319-334 is dead code (Q3) and is not restated."""
        n = self._ln(x, p + ".norm1", 0.0)
        q = self._linear(n, p + ".attn1.to_q", bias=False)
        enc = self._q(enc)
        kv_src = n
        if self.cfg.need_t2i_ip_adapter and vis_idx is not None and T > 1:
            bt, hw, c = n.shape
            nb = n.view(bt // T, T, hw, c)
            ip = nb[:, vis_idx].reshape(bt // T, 1, len(vis_idx) * hw, c).expand(-1, T, -1, -1)
            kv_src = torch.cat([nb, ip], dim=2).reshape(bt, -1, c)
        k = self._linear(kv_src, p + ".attn1.to_k", bias=False)
        v = self._linear(kv_src, p + ".attn1.to_v", bias=False)
        x = self._qs(self._linear(self._sdpa(q, k, v), p + ".attn1.to_out.0") + x)
        n = self._ln(x, p + ".norm2", 1e-5)
        q = self._linear(n, p + ".attn2.to_q", bias=False)
        k = self._linear(enc, p + ".attn2.to_k", bias=False)
        v = self._linear(enc, p + ".attn2.to_v", bias=False)
        a = self._sdpa(q, k, v)
        if (p + ".attn2.to_k_ip.weight") in self.sd and clip_emb is not None and ip_scale > 0:
            ik = self._linear(self._q(clip_emb), p + ".attn2.to_k_ip", bias=False)
            iv = self._linear(self._q(clip_emb), p + ".attn2.to_v_ip", bias=False)
            a = self._q(a + ip_scale * self._sdpa(q, ik, iv))
        x = self._qs(self._linear(a, p + ".attn2.to_out.0") + x)
        n = self._ln(x, p + ".norm3", 0.0)
        return self._qs(self.feed_forward(n, p) + x)

    def temporal_block(self, x, p):
        """musev BasicTransformerBlock with double_self_attention=True (temporal_transformer.py:145-163):
        attn1 and attn2 are both plain self-attention over the frame axis."""
        n = self._ln(x, p + ".norm1", 0.0)
        a = self._sdpa(self._linear(n, p + ".attn1.to_q", False), self._linear(n, p + ".attn1.to_k", False),
                       self._linear(n, p + ".attn1.to_v", False))
        x = self._qs(self._linear(a, p + ".attn1.to_out.0") + x)
        n = self._ln(x, p + ".norm2", 1e-5)
        a = self._sdpa(self._linear(n, p + ".attn2.to_q", False), self._linear(n, p + ".attn2.to_k", False),
                       self._linear(n, p + ".attn2.to_v", False))
        x = self._qs(self._linear(a, p + ".attn2.to_out.0") + x)
        n = self._ln(x, p + ".norm3", 0.0)
        return self._qs(self.feed_forward(n, p) + x)

    def spatial_transformer(self, x, enc, T, p, vis_idx, clip_emb, ip_scale):
        """musev Transformer2DModel.forward continuous path (musev/models/transformer_2d.py:257-276,313-389):
        GroupNorm(eps 1e-6) -> 1x1 conv -> tokens -> block -> 1x1 conv -> + residual."""
        bt, c, hh, ww = x.shape
        h = self._gn(x, p + ".norm", 1e-6)
        h = self._qs(F.conv2d(h, self._w(p + ".proj_in.weight"), self._w(p + ".proj_in.bias")))
        h = h.permute(0, 2, 3, 1).reshape(bt, hh * ww, c)
        h = self.spatial_block(h, enc, T, p + ".transformer_blocks.0", vis_idx, clip_emb, ip_scale)
        h = h.reshape(bt, hh, ww, c).permute(0, 3, 1, 2)
        h = F.conv2d(self._q(h), self._w(p + ".proj_out.weight"), self._w(p + ".proj_out.bias"))
        return self._qs(h + x)

    def temporal_transformer(self, x, femb, T, p):
        """TransformerTemporalModel.forward, musev/models/temporal_transformer.py:189-308: GroupNorm(eps 1e-6) over
        (c/g, t, h, w) -> (b h w) t c -> proj_in -> + frame_emb_proj(SiLU(femb)) -> block -> proj_out ->
        residual + |temporal_weight| * x."""
        if self.skip_temporal:
            return x
        bt, c, hh, ww = x.shape
        b = bt // T
        v = x.view(b, T, c, hh, ww).permute(0, 2, 1, 3, 4)
        residual = v
        v = self._gn(v, p + ".norm", 1e-6)
        v = v.permute(0, 3, 4, 2, 1).reshape(b * hh * ww, T, c)
        v = self._linear(v, p + ".proj_in")
        fe = self._linear(self._q(F.silu(femb)), p + ".frame_emb_proj")         # [b, T, c]
        v = self._qs(v + fe.repeat_interleave(hh * ww, dim=0))           # align_repeat_tensor_single_dim
        v = self.temporal_block(v, p + ".transformer_blocks.0")
        v = self._linear(self._q(v), p + ".proj_out")
        v = v.view(b, hh, ww, T, c).permute(0, 4, 3, 1, 2)
        v = self._qs(residual + torch.abs(self._w(p + ".temporal_weight")) * v)
        return v.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)

    def refer_fuse(self, x, ref, T, p):
        """ReferEmbFuseAttention.forward, musev/models/attention_processor.py:629-750: Q = frame tokens,
        K/V = [reference-feature tokens ; frame tokens], no input norm, q/k/v without bias, residual."""
        bt, c, hh, ww = x.shape
        b = bt // T
        tok = x.permute(0, 2, 3, 1).reshape(bt, hh * ww, c)
        r = ref.permute(0, 2, 3, 4, 1).reshape(b, -1, ref.shape[1])       # b (t2 h w) c
        r = r.repeat_interleave(T, dim=0)
        enc = self._q(torch.cat([r, tok], dim=1))
        a = self._sdpa(self._linear(self._q(tok), p + ".to_q", False), self._linear(enc, p + ".to_k", False),
                       self._linear(enc, p + ".to_v", False))
        a = self._linear(a, p + ".to_out.0")
        return self._qs(a.reshape(bt, hh, ww, c).permute(0, 3, 1, 2) + x)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, sample_index=None,
                vision_condition_frames_sample=None, vision_conditon_frames_sample_index=None, sample_frame_rate=10,
                down_block_refer_embs: Optional[Sequence[torch.Tensor]] = None, mid_block_refer_emb=None,
                vision_clip_emb=None, ip_adapter_scale=1.0, down_block_additional_residuals=None,
                mid_block_additional_residual=None, skip_temporal_layers=False, **unused):
        """UNet3DConditionModel.forward, musev/models/unet_3d_condition.py:773-1280."""
        cfg = self.cfg
        dev, dt = self.device, self.dtype
        self.skip_temporal = bool(skip_temporal_layers)
        sample = sample.to(dev, dt)
        enc = encoder_hidden_states.to(dev, dt)
        vis_idx = vision_conditon_frames_sample_index
        if vis_idx is not None:
            vis_idx = [int(i) for i in torch.as_tensor(vis_idx).tolist()]
        if vision_condition_frames_sample is not None:
            # batch_concat_two_tensor_with_index, musev/data/data_util.py:242-292 (unet_3d_condition.py:875-882)
            vc = vision_condition_frames_sample.to(dev, dt)
            sidx = [int(i) for i in torch.as_tensor(sample_index).tolist()]
            total = sample.shape[2] + vc.shape[2]
            merged = sample.new_zeros(sample.shape[0], sample.shape[1], total, *sample.shape[3:])
            merged[:, :, sidx] = sample
            merged[:, :, vis_idx] = vc
            sample = merged
        B, _, T, H, W = sample.shape
        # 1. time embedding (unet_3d_condition.py:887-906)
        t = torch.as_tensor(timestep, device=dev).reshape(-1).expand(B)
        emb = self._mlp_emb(timestep_embedding(t, cfg.block_out_channels[0]).to(dt), "time_embedding")
        if cfg.use_anivv1_cfg:
            emb = F.silu(emb)
        emb = emb.repeat_interleave(T, dim=0)
        if cfg.keep_vision_condtion and T > 1 and sample_index is not None and vis_idx is not None:
            emb = emb.view(B, T, -1).clone()
            emb[:, vis_idx, :] = 0
            emb = emb.view(B * T, -1)
        # 2. frame embedding (unet_3d_condition.py:909-937)
        fidx = torch.arange(T, dtype=torch.long, device=dev)
        if cfg.use_anivv1_cfg:
            fidx = (fidx * sample_frame_rate).to(torch.long)
        femb = timestep_embedding(fidx, cfg.block_out_channels[0]).to(dt)[None].expand(B, -1, -1)
        femb = self._mlp_emb(femb, "frame_embedding")
        if cfg.use_anivv1_cfg:
            femb = F.silu(femb)
        # 3. per-frame copies of the text / image conditioning (unet_3d_condition.py:938-957)
        enc = enc.repeat_interleave(T, dim=0)
        clip = None
        if cfg.ip_adapter_cross_attn and vision_clip_emb is not None:
            clip = vision_clip_emb.to(dev, dt).repeat_interleave(T, dim=0)
        # 4. conv_in (unet_3d_condition.py:1008-1009)
        x = sample.permute(0, 2, 1, 3, 4).reshape(B * T, -1, H, W)
        x = self._qs(F.conv2d(self._q(x), self._w("conv_in.weight"), self._w("conv_in.bias"), padding=1))
        self._tap("conv_in", x)
        if cfg.need_transformer_in:
            x = self.temporal_transformer(x, femb, T, "transformer_in")
            self._tap("transformer_in", x)
        use_ref = cfg.need_refer_emb and down_block_refer_embs is not None
        if use_ref:
            refs = [r.to(dev, dt) for r in down_block_refer_embs]
            x = self.refer_fuse(x, refs[0], T, "first_refer_emb_attns")
            self._tap("first_refer", x)
        # 5. down (unet_3d_condition.py:1075-1145; blocks: unet_3d_blocks.py:594-772, 884-985)
        skips = [x]
        nb = len(cfg.block_out_channels)
        for i in range(nb):
            final = i == nb - 1
            has_attn = not final
            if use_ref:
                num_block = cfg.layers_per_block + int(not final)      # Q19: slice start uses *this* block's count
                start = 1 + num_block * i
                block_refs = refs[start:start + num_block]
            for j in range(cfg.layers_per_block):
                p = f"down_blocks.{i}"
                x = self.resnet(x, emb, f"{p}.resnets.{j}")
                self._tap(f"{p}.resnets.{j}", x)
                x = self.temp_conv(x, T, f"{p}.temp_convs.{j}")
                self._tap(f"{p}.temp_convs.{j}", x)
                if has_attn:
                    x = self.spatial_transformer(x, enc, T, f"{p}.attentions.{j}", vis_idx, clip, ip_adapter_scale)
                    self._tap(f"{p}.attentions.{j}", x)
                    x = self.temporal_transformer(x, femb, T, f"{p}.temp_attentions.{j}")
                    self._tap(f"{p}.temp_attentions.{j}", x)
                # AdaIN toward the vision-condition frame is an identity (Q2) -- nothing to do
                if use_ref:
                    x = self.refer_fuse(x, block_refs[j], T, f"{p}.refer_emb_attns.{j}")
                    self._tap(f"{p}.refer_emb_attns.{j}", x)
                skips.append(x)
            if not final:
                p = f"down_blocks.{i}.downsamplers.0.conv"
                x = self._qs(F.conv2d(self._q(x), self._w(p + ".weight"), self._w(p + ".bias"), stride=2, padding=1))
                if use_ref:
                    x = self.refer_fuse(x, block_refs[cfg.layers_per_block], T,
                                        f"down_blocks.{i}.refer_emb_attns.{cfg.layers_per_block}")
                self._tap(f"down_blocks.{i}.down", x)
                skips.append(x)
        if down_block_additional_residuals is not None:
            skips = [s + r.to(dev, dt) for s, r in zip(skips, down_block_additional_residuals)]
        # 6. mid (unet_3d_blocks.py:364-433)
        x = self.resnet(x, emb, "mid_block.resnets.0")
        x = self.temp_conv(x, T, "mid_block.temp_convs.0")
        x = self.spatial_transformer(x, enc, T, "mid_block.attentions.0", vis_idx, clip, ip_adapter_scale)
        x = self.temporal_transformer(x, femb, T, "mid_block.temp_attentions.0")
        x = self.resnet(x, emb, "mid_block.resnets.1")
        x = self.temp_conv(x, T, "mid_block.temp_convs.1")
        self._tap("mid", x)
        if cfg.need_refer_emb and mid_block_refer_emb is not None:
            x = self.refer_fuse(x, mid_block_refer_emb.to(dev, dt), T, "mid_block_refer_emb_attns")
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual.to(dev, dt)
        # 7. up (unet_3d_condition.py:1199-1255; unet_3d_blocks.py:1106-1251, 1324-1413)
        for i in range(nb):
            has_attn = i > 0
            final = i == nb - 1
            for j in range(cfg.layers_per_block + 1):
                p = f"up_blocks.{i}"
                x = torch.cat([x, skips.pop()], dim=1)
                x = self.resnet(x, emb, f"{p}.resnets.{j}")
                x = self.temp_conv(x, T, f"{p}.temp_convs.{j}")
                if has_attn:
                    x = self.spatial_transformer(x, enc, T, f"{p}.attentions.{j}", vis_idx, clip, ip_adapter_scale)
                    x = self.temporal_transformer(x, femb, T, f"{p}.temp_attentions.{j}")
                self._tap(f"{p}.{j}", x)
            if not final:
                # Upsample2D: nearest x2 then 3x3 conv (diffusers models/resnet.py:167-210)
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")
                p = f"up_blocks.{i}.upsamplers.0.conv"
                x = self._qs(F.conv2d(self._q(x), self._w(p + ".weight"), self._w(p + ".bias"), padding=1))
                self._tap(f"up_blocks.{i}.up", x)
        # 8. out (unet_3d_condition.py:1258-1263)
        x = self._q(F.silu(self._gn(x, "conv_norm_out", cfg.norm_eps)))
        x = F.conv2d(x, self._w("conv_out.weight"), self._w("conv_out.bias"), padding=1)
        return x.view(B, T, -1, H, W).permute(0, 2, 1, 3, 4).contiguous()

    __call__ = forward
