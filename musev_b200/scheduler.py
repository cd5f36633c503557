"""Drop-in mirror of `musev.schedulers.DDIMScheduler` (musev/schedulers/scheduling_ddim.py:44-302, on top of
diffusers/src/diffusers/schedulers/scheduling_ddim.py:131-342,473-495).

`step` runs the fused device kernel `mvb_fuse_cfg_ddim` (cfg = 0 form); the parallel-denoise loop in
musev_b200.pipeline calls the same kernel with the overlap mean and classifier-free guidance folded in.
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import ops


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.prev_sample, self.pred_original_sample)[i]


_PRED = {"epsilon": 0, "v_prediction": 1, "sample": 2}


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False):
        if thresholding:
            raise NotImplementedError("dynamic thresholding is unsuitable for latent diffusion and is not supported")
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if rescale_betas_zero_snr:
            self.betas = _rescale_zero_terminal_snr(self.betas)
        if prediction_type not in _PRED:
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or `v_prediction`")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
            trained_betas=trained_betas, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
            steps_offset=steps_offset, prediction_type=prediction_type, thresholding=thresholding,
            dynamic_thresholding_ratio=dynamic_thresholding_ratio, clip_sample_range=clip_sample_range,
            sample_max_value=sample_max_value, timestep_spacing=timestep_spacing,
            rescale_betas_zero_snr=rescale_betas_zero_snr)

    def scale_model_input(self, sample: torch.Tensor, timestep: Optional[int] = None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device] = None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {c.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {c.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ratio = c.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ratio = c.num_train_timesteps / num_inference_steps
            ts = np.round(np.arange(c.num_train_timesteps, 0, -ratio)).astype(np.int64) - 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported. Please make sure to choose one of 'leading' or 'trailing'.")
        self.timesteps = torch.from_numpy(ts).to(device)

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        return (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)

    def step_scalars(self, timestep: int, eta: float = 0.0) -> Tuple[float, float, float]:
        """(alpha_prod_t, alpha_prod_t_prev, std_dev_t) of `step` (scheduling_ddim.py:198-248)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        std = float(eta * self._get_variance(t, prev_t) ** 0.5) if eta > 0 else 0.0
        return a_t, a_p, std

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = True, w_ind_noise: float = 0.5, noise_type: str = "random"):
        if not sample.is_cuda:
            raise RuntimeError("musev_b200.DDIMScheduler.step runs on the GPU only")
        a_t, a_p, std = self.step_scalars(timestep, eta)
        shape = sample.shape
        x = sample.contiguous()
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        x5 = x.view(shape[0], shape[1], 1, 1, -1) if x.dim() != 5 else x
        eps = model_output.contiguous().float().view(x5.shape)
        noise = None
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                                 " `variance_noise` stays `None`.")
            # musev/schedulers/scheduling_ddim.py:273-292: the reference commented out the `if variance_noise is None`
            # guard and ALWAYS redraws the noise from `generator`; a caller-supplied tensor is ignored. Same here, unless
            # `self.honor_variance_noise` is set (an extension for deterministic tests; off by default).
            if variance_noise is None or not getattr(self, "honor_variance_noise", False):
                variance_noise = _variance_noise(model_output, generator, noise_type, w_ind_noise)
            noise = variance_noise.to(sample.device).contiguous().float().view(x5.shape)
        x0 = torch.empty(x5.shape, dtype=torch.float32, device=x.device)
        c = self.config
        prev = ops.fuse_cfg_ddim(eps, None, x5, 1.0, a_t, a_p, _PRED[c.prediction_type],
                                 c.clip_sample_range if c.clip_sample else 0.0, cfg=False,
                                 use_clipped=use_clipped_model_output, std_dev=std, noise=noise, x0_out=x0)
        prev = prev.view(shape).to(sample.dtype)
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0.view(shape).to(sample.dtype))

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        # one-shot, before the loop (pipeline_controlnet.py:240-431): plain tensor arithmetic
        a = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)[timesteps.to(original_samples.device)]
        sa, sb = (a ** 0.5).flatten(), ((1 - a) ** 0.5).flatten()
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def __len__(self):
        return self.config.num_train_timesteps


def _variance_noise(model_output, generator, noise_type, w_ind_noise):
    shape, dev, dt = model_output.shape, model_output.device, model_output.dtype
    if noise_type == "random":
        return torch.randn(shape, generator=generator, device=dev, dtype=dt)
    if noise_type == "video_fusion":
        # musev/utils/noise_util.py:32-83: shared noise per video + independent noise per frame
        b, ch, t, h, w = shape
        common = torch.randn((b, ch, 1, h, w), generator=generator, device=dev, dtype=dt)
        ind = torch.randn(shape, generator=generator, device=dev, dtype=dt)
        return (1 - w_ind_noise) ** 0.5 * common + w_ind_noise ** 0.5 * ind
    raise ValueError(f"unknown noise_type {noise_type}")


def _rescale_zero_terminal_snr(betas):
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = (alphas_bar_sqrt - aT) * a0 / (a0 - aT)
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
    return 1 - alphas


# SD-1.5 scheduler_config.json as used by the reference's DDIM path
SD15_DDIM_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                        clip_sample=False, set_alpha_to_one=False, steps_offset=1)
