"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's scheduler, window schedule and parallel-denoise loop.

  * DDIMOracle            musev/schedulers/scheduling_ddim.py:136-302 + diffusers schedulers/scheduling_ddim.py
                          :186-237 (betas), :299-342 (set_timesteps), :473-495 (add_noise)
                          pinned by the upstream known-answer tests diffusers/tests/schedulers/test_scheduler_ddim.py
                          :46-54,102-176 (replayed in tests/test_oracle_pinned.py)
  * window schedules      musev/pipelines/context.py:21-66,105-149 + MMCM/mmcm/utils/itertools_util.py:6-46
                          pinned against the imported reference by oracle/make_golden.py (tests/golden/contexts.json)
  * denoise_loop          musev/pipelines/pipeline_controlnet.py:1846-2117 (loop body only: window gather, vis-cond
                          concat, UNet, overlap mean, CFG, scheduler.step); no upstream test exists for it.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


class DDIMOracle:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 clip_sample_range=1.0, timestep_spacing="leading"):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        if self.timestep_spacing == "leading":
            ratio = self.num_train_timesteps // n
            ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(self.num_train_timesteps, 0, -self.num_train_timesteps / n)).astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts)

    def _get_variance(self, t, prev_t):
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)

    def add_noise(self, x, noise, timesteps):
        a = self.alphas_cumprod.to(x.dtype)[timesteps]
        sa = (a ** 0.5).flatten()
        sb = ((1 - a) ** 0.5).flatten()
        while sa.dim() < x.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * x + sb * noise

    def step(self, model_output, timestep, sample, eta=0.0, variance_noise=None):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "v_prediction":
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        std = eta * self._get_variance(t, prev_t) ** 0.5
        prev = a_p ** 0.5 * x0 + (1 - a_p - std ** 2) ** 0.5 * eps
        if eta > 0:
            if variance_noise is None:
                raise ValueError("eta > 0 needs variance_noise in the oracle")
            prev = prev + std * variance_noise
        return prev, x0


# SD-1.5 scheduler_config.json values the reference's DDIM path uses
SD15_DDIM = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1)


# ---------------------------------------------------------------------------------------- window schedules
def _ordered_halving(val: int) -> float:
    return int(f"{val:064b}"[::-1], 2) / (1 << 64)


def uniform(step, num_frames, context_size, context_stride=3, context_overlap=4, closed_loop=True):
    """musev/pipelines/context.py:21-48."""
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    for context_step in 1 << np.arange(context_stride):
        pad = int(round(num_frames * _ordered_halving(step)))
        for j in range(int(_ordered_halving(step) * context_step) + pad,
                       num_frames + pad + (0 if closed_loop else -context_overlap),
                       (context_size * context_step - context_overlap)):
            yield [e % num_frames for e in range(j, j + context_size * context_step, context_step)]


def uniform_v2(num_frames, context_size, context_overlap):
    """musev/pipelines/context.py:51-66 -> generate_sample_idxs (MMCM/mmcm/utils/itertools_util.py:6-46)."""
    out, start, step = [], 0, context_size - context_overlap
    while start < num_frames:
        out.append(list(range(start, min(start + context_size, num_frames))))
        start += step
    return out


def prepare_global_context(context_schedule, num_inference_steps, time_size, context_frames, context_stride,
                           context_overlap, context_batch_size):
    """musev/pipelines/context.py:120-149 (including drop_last_repeat_context :105-117)."""
    if context_schedule == "uniform":
        q = [list(map(int, c)) for c in uniform(0, time_size, context_frames, context_stride, context_overlap)]
    elif context_schedule == "uniform_v2":
        q = uniform_v2(time_size, context_frames, context_overlap)
    else:
        raise ValueError(f"Unknown context_overlap policy {context_schedule}")
    if len(q) >= 2 and q[-1][-1] == q[-2][-1]:
        q = q[:-1]
    n = math.ceil(len(q) / context_batch_size)
    return [q[i * context_batch_size:(i + 1) * context_batch_size] for i in range(n)]


# ---------------------------------------------------------------------------------------- the loop
@torch.no_grad()
def denoise_loop(unet: Callable, scheduler: DDIMOracle, latents: torch.Tensor, condition_latents: torch.Tensor,
                 prompt_embeds: torch.Tensor, num_inference_steps: int, guidance_scale: float,
                 context_frames: int = 12, context_overlap: int = 4, context_schedule: str = "uniform_v2",
                 context_stride: int = 1, motion_speed: float = 8.0, unet_kwargs: Optional[dict] = None,
                 return_eps: bool = False, controlnet: Optional[Callable] = None,
                 controlnet_latents: Optional[torch.Tensor] = None, controlnet_conditioning_scale: float = 1.0):
    """Loop body of MusevControlNetPipeline.__call__, musev/pipelines/pipeline_controlnet.py:1846-2117, for the
    CFG-on case (guidance_scale > 1; Q14). With `controlnet` (a callable with the diffusers ControlNetModel signature) and
    `controlnet_latents` [2B, C0, n_vc + T, h, w] the per-window ControlNet call of :1992-2038 / :1202-1291 is included
    (guess_mode False, controlnet_keep 1).

    latents [B,4,T,h,w]; condition_latents [B,4,n_vc,h,w]; prompt_embeds [2B,77,768] = cat([negative, positive]).
    `unet(sample, t, encoder_hidden_states, **kw)` returns eps [2B,4,n_vc+Tc,h,w].
    """
    unet_kwargs = dict(unet_kwargs or {})
    B, C, T, h, w = latents.shape
    n_vc = condition_latents.shape[2]
    vis_idx = torch.arange(n_vc)
    scheduler.set_timesteps(num_inference_steps)
    contexts = prepare_global_context(context_schedule, num_inference_steps, T, context_frames, context_stride,
                                      context_overlap, 1)
    eps_trace = []
    for t in scheduler.timesteps:
        noise_pred = torch.zeros(2 * B, C, T, h, w, dtype=latents.dtype)             # :1870-1876
        counter = torch.zeros(1, 1, T, 1, 1, dtype=latents.dtype)                    # :1877-1882
        for context in contexts:                                                     # :1900
            c = context[0]
            lat_c = latents[:, :, c]                                                 # :1902
            model_in = torch.cat([lat_c] * 2)                                        # :1908-1910
            if hasattr(scheduler, "scale_model_input"):
                model_in = scheduler.scale_model_input(model_in, t)                  # :1911 (identity for DDIM / LCM)
            sub_idx = torch.arange(len(c)) + n_vc                                    # :1914-1920
            cond = torch.cat([condition_latents] * 2)
            full = torch.zeros(2 * B, C, n_vc + len(c), h, w, dtype=latents.dtype)   # batch_concat_two_tensor_with_index
            full[:, :, vis_idx] = cond
            full[:, :, sub_idx] = model_in
            kw = dict(unet_kwargs)
            if controlnet is not None:
                ctx = list(range(n_vc)) + [ci + n_vc for ci in c]                    # :1997-2000
                lat_c = controlnet_latents[:, :, ctx]                                # :2008-2013
                nb, _, tc, _, _ = full.shape
                x2 = full.permute(0, 2, 1, 3, 4).reshape(nb * tc, C, h, w)           # :1236-1238
                lat2 = lat_c.permute(0, 2, 1, 3, 4).reshape(nb * tc, lat_c.shape[1], h, w)
                enc2 = prompt_embeds.repeat_interleave(tc, dim=0)                    # :1242-1246
                down, mid = controlnet(x2, t, enc2, controlnet_cond_latents=lat2,
                                       conditioning_scale=controlnet_conditioning_scale)   # :1253-1262
                kw["down_block_additional_residuals"] = down
                kw["mid_block_additional_residual"] = mid
            eps = unet(full, t, prompt_embeds, sample_index=sub_idx, vision_conditon_frames_sample_index=vis_idx,
                       sample_frame_rate=motion_speed, **kw)                         # :2045-2067
            eps = eps[:, :, sub_idx]                                                 # :2068-2071
            noise_pred[:, :, c] = noise_pred[:, :, c] + eps                          # :2076
            counter[:, :, c] = counter[:, :, c] + 1                                  # :2077
        noise_pred = noise_pred / counter                                            # :2079
        uncond, text = noise_pred.chunk(2)                                           # :2101-2105
        noise_pred = uncond + guidance_scale * (text - uncond)
        if return_eps:
            eps_trace.append(noise_pred.clone())
        latents = scheduler.step(noise_pred, t, latents)[0]                          # :2112-2117 (eta = 0)
    return (latents, eps_trace) if return_eps else latents
