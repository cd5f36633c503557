"""Host mirrors of the other samplers of SURVEY.md 8(f)-4 on the fused step kernel.

  * `EulerDiscreteScheduler`  musev/schedulers/scheduling_euler_discrete.py:21-293 over diffusers
                              schedulers/scheduling_euler_discrete.py:135-463 -- the predictor's DEFAULT sampler
                              (musev/pipelines/pipeline_controlnet_predictor.py:258-261).
  * `LCMScheduler`            musev/schedulers/scheduling_lcm.py:44-312 over diffusers schedulers/scheduling_lcm.py:196-547.
Both keep the reference's constructor / `set_timesteps` / `timesteps` / `sigmas` / `init_noise_sigma` /
`scale_model_input` / `step` / `add_noise` surface (the pipeline probes `inspect.signature(step)` for `generator` /
`noise_type`, pipeline_controlnet.py:1690-1696). Every step of these samplers is affine in (sample, model_output, noise),
so `step` and the fused loop (`ParallelDenoiser`) run ONE kernel (`mvb_fuse_cfg_affine`) with host-computed scalars
(`affine_step`). Integer / float bookkeeping restated from the reference; the tensor arithmetic is on the GPU only.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import ops
from .scheduler import _rescale_zero_terminal_snr, _variance_noise


def _betas(beta_schedule, beta_start, beta_end, n, trained_betas, cls):
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(f"{beta_schedule} does is not implemented for {cls}")


@dataclass
class AffineStep:
    """x_prev = c_x x + c_e eps + c_n noise; aux (pred_original_sample / denoised) = a_x x + a_e eps."""
    c_x: float
    c_e: float
    c_n: float
    a_x: float
    a_e: float


@dataclass
class EulerDiscreteSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


@dataclass
class LCMSchedulerOutput:
    prev_sample: torch.Tensor
    denoised: Optional[torch.Tensor] = None


def _run_affine(a: AffineStep, model_output, sample, noise):
    if not sample.is_cuda:
        raise RuntimeError("musev_b200 samplers run on the GPU only")
    shape = sample.shape
    x = sample.contiguous()
    if x.dtype not in (torch.float16, torch.float32):
        x = x.float()
    x5 = x.view(shape[0], shape[1], 1, 1, -1) if x.dim() != 5 else x
    eps = model_output.contiguous().float().view(x5.shape)
    nz = None if noise is None else noise.to(sample.device).contiguous().float().view(x5.shape)
    aux = torch.empty(x5.shape, dtype=torch.float32, device=x.device)
    prev = ops.fuse_cfg_affine(eps, None, x5, 1.0, a.c_x, a.c_e, a.c_n, nz, a.a_x, a.a_e, aux, cfg=False)
    return prev.view(shape).to(sample.dtype), aux.view(shape).to(sample.dtype)


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, prediction_type: str = "epsilon",
                 interpolation_type: str = "linear", use_karras_sigmas: Optional[bool] = False,
                 timestep_spacing: str = "linspace", steps_offset: int = 0):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, trained_betas=trained_betas, prediction_type=prediction_type,
                                      interpolation_type=interpolation_type, use_karras_sigmas=use_karras_sigmas,
                                      timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps, trained_betas, self.__class__)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.sigmas = torch.from_numpy(np.concatenate([sig[::-1], [0.0]]).astype(np.float32))
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy())
        self.is_scale_input_called = False
        self.use_karras_sigmas = use_karras_sigmas
        self._step_index = None

    @property
    def init_noise_sigma(self):
        if self.config.timestep_spacing in ["linspace", "trailing"]:
            return self.sigmas.max()
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _init_step_index(self, timestep):
        t = float(timestep)
        cand = (self.timesteps == t).nonzero()
        if len(cand) == 0:
            raise ValueError(f"timestep {t} is not one of scheduler.timesteps")
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def model_input_scale(self, timestep) -> float:
        """1 / sqrt(sigma^2 + 1) of `scale_model_input` as a host scalar."""
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self.sigmas[self._step_index])
        self.is_scale_input_called = True
        return 1.0 / (sigma * sigma + 1.0) ** 0.5

    def scale_model_input(self, sample: torch.Tensor, timestep) -> torch.Tensor:
        return sample * self.model_input_scale(timestep)

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif c.timestep_spacing == "leading":
            ratio = c.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ratio = c.num_train_timesteps / num_inference_steps
            ts = (np.arange(c.num_train_timesteps, 0, -ratio)).round().copy().astype(np.float32)
            ts -= 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported. Please make sure to choose one of 'linspace', 'leading' or 'trailing'.")
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        log_sig = np.log(sig)
        if c.interpolation_type == "linear":
            sig = np.interp(ts, np.arange(0, len(sig)), sig)
        elif c.interpolation_type == "log_linear":
            sig = torch.linspace(np.log(sig[-1]), np.log(sig[0]), num_inference_steps + 1).exp().numpy()
        else:
            raise ValueError(f"{c.interpolation_type} is not implemented. Please specify interpolation_type to either 'linear' or 'log_linear'")
        if self.use_karras_sigmas:
            smin, smax, rho = float(sig[-1]), float(sig[0]), 7.0
            ramp = np.linspace(0, 1, num_inference_steps)
            sig = (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho
            ts = np.array([self._sigma_to_t(s, log_sig) for s in sig])
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self._step_index = None

    @staticmethod
    def _sigma_to_t(sigma, log_sigmas):
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        return ((1 - w) * low_idx + w * high_idx).reshape(np.shape(sigma))

    def affine_step(self, timestep, s_churn: float = 0.0, s_tmin: float = 0.0, s_tmax: float = float("inf"),
                    s_noise: float = 1.0) -> AffineStep:
        """Scalars of `step` (scheduling_euler_discrete.py:107-166) at the current step index; advances the index."""
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self.sigmas[self._step_index])
        gamma = min(s_churn / (len(self.sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigma <= s_tmax else 0.0
        sigma_hat = sigma * (gamma + 1)
        c_n = s_noise * (sigma_hat ** 2 - sigma ** 2) ** 0.5 if gamma > 0 else 0.0
        dt = float(self.sigmas[self._step_index + 1]) - sigma_hat
        pt = self.config.prediction_type
        # x0 = a_x x' + a_e e  (x' = x + c_n noise);  prev = x' + (x' - x0) / sigma_hat * dt
        if pt in ("original_sample", "sample"):
            a_x, a_e = 0.0, 1.0
        elif pt == "epsilon":
            a_x, a_e = 1.0, -sigma_hat
        elif pt == "v_prediction":
            a_x, a_e = 1.0 / (sigma ** 2 + 1), -sigma / (sigma ** 2 + 1) ** 0.5
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, or `v_prediction`")
        r = dt / sigma_hat
        c_x = 1.0 + r * (1.0 - a_x)
        c_e = -r * a_e
        self._step_index += 1
        # the churn noise enters through x': prev = c_x (x + c_n z) + c_e e, aux = a_x (x + c_n z) + a_e e; with the default
        # s_churn = 0 there is no noise. (aux ignores the churn term: it is only reported, never fed back.)
        return AffineStep(c_x, c_e, c_x * c_n, a_x, a_e)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, s_churn: float = 0.0, s_tmin: float = 0.0,
             s_tmax: float = float("inf"), s_noise: float = 1.0, generator=None, return_dict: bool = True,
             w_ind_noise: float = 0.5, noise_type: str = "random"):
        if isinstance(timestep, int) or isinstance(timestep, (torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                             " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                             " one of the `scheduler.timesteps` as a timestep.")
        a = self.affine_step(timestep, s_churn, s_tmin, s_tmax, s_noise)
        # the reference draws the noise on every step (scheduling_euler_discrete.py:116-127) even when gamma = 0, which
        # advances the generator; keep that side effect
        noise = _variance_noise(model_output, generator, noise_type, w_ind_noise)
        prev, x0 = _run_affine(a, model_output, sample, noise if a.c_n != 0.0 else None)
        if not return_dict:
            return (prev,)
        return EulerDiscreteSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
        sched_t = self.timesteps.to(original_samples.device)
        idx = [(sched_t == t).nonzero().item() for t in timesteps.to(original_samples.device)]
        sigma = sigmas[idx].flatten()
        while sigma.dim() < original_samples.dim():
            sigma = sigma.unsqueeze(-1)
        return original_samples + noise * sigma

    def __len__(self):
        return self.config.num_train_timesteps


class LCMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", trained_betas=None, original_inference_steps: int = 50,
                 clip_sample: bool = False, clip_sample_range: float = 1.0, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, sample_max_value: float = 1.0,
                 timestep_spacing: str = "leading", timestep_scaling: float = 10.0, rescale_betas_zero_snr: bool = False):
        if thresholding or clip_sample:
            raise NotImplementedError("musev_b200.LCMScheduler runs the affine fused step: clip_sample / thresholding are not supported")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, trained_betas=trained_betas,
                                      original_inference_steps=original_inference_steps, clip_sample=clip_sample,
                                      clip_sample_range=clip_sample_range, set_alpha_to_one=set_alpha_to_one,
                                      steps_offset=steps_offset, prediction_type=prediction_type, thresholding=thresholding,
                                      dynamic_thresholding_ratio=dynamic_thresholding_ratio, sample_max_value=sample_max_value,
                                      timestep_spacing=timestep_spacing, timestep_scaling=timestep_scaling,
                                      rescale_betas_zero_snr=rescale_betas_zero_snr)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps, trained_betas, self.__class__)
        if rescale_betas_zero_snr:
            self.betas = _rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def _init_step_index(self, timestep):
        cand = (self.timesteps == int(timestep)).nonzero()
        if len(cand) == 0:
            raise ValueError(f"timestep {int(timestep)} is not one of scheduler.timesteps")
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def model_input_scale(self, timestep) -> float:
        return 1.0

    def set_timesteps(self, num_inference_steps: int, device=None, original_inference_steps: Optional[int] = None,
                      strength: float = 1.0):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                             f" {c.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                             f" maximal {c.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        original_steps = original_inference_steps if original_inference_steps is not None else c.original_inference_steps
        if original_steps > c.num_train_timesteps:
            raise ValueError(f"`original_steps`: {original_steps} cannot be larger than `self.config.train_timesteps`:"
                             f" {c.num_train_timesteps}")
        if num_inference_steps > original_steps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than `original_inference_steps`:"
                             f" {original_steps}")
        k = c.num_train_timesteps // original_steps
        origin = np.asarray(list(range(1, int(original_steps * strength) + 1))) * k - 1
        if len(origin) // num_inference_steps < 1:
            raise ValueError(f"The combination of `original_steps x strength`: {original_steps} x {strength} is smaller than"
                             f" `num_inference_steps`: {num_inference_steps}.")
        origin = origin[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
        self.timesteps = torch.from_numpy(origin[idx]).to(dtype=torch.long)
        self._step_index = None

    def get_scalings_for_boundary_condition_discrete(self, timestep) -> Tuple[float, float]:
        sigma_data = 0.5
        st = float(timestep) * self.config.timestep_scaling
        return sigma_data ** 2 / (st ** 2 + sigma_data ** 2), st / (st ** 2 + sigma_data ** 2) ** 0.5

    def affine_step(self, timestep) -> AffineStep:
        """Scalars of `step` (musev/schedulers/scheduling_lcm.py:232-305); advances the step index."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        t = int(timestep)
        nxt = self._step_index + 1
        prev_t = int(self.timesteps[nxt]) if nxt < len(self.timesteps) else t
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        sa, sb = a_t ** 0.5, (1 - a_t) ** 0.5
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(t)
        pt = self.config.prediction_type
        if pt == "epsilon":
            x0_x, x0_e = 1.0 / sa, -sb / sa
        elif pt == "sample":
            x0_x, x0_e = 0.0, 1.0
        elif pt == "v_prediction":
            x0_x, x0_e = sa, -sb
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, `sample` or `v_prediction` for `LCMScheduler`.")
        d_x, d_e = c_out * x0_x + c_skip, c_out * x0_e                       # denoised = d_x x + d_e e
        last = self._step_index == self.num_inference_steps - 1
        self._step_index += 1
        if last:
            return AffineStep(d_x, d_e, 0.0, d_x, d_e)
        return AffineStep(a_p ** 0.5 * d_x, a_p ** 0.5 * d_e, (1 - a_p) ** 0.5, d_x, d_e)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None, return_dict: bool = True):
        a = self.affine_step(timestep)
        noise = None
        if a.c_n != 0.0:
            noise = torch.randn(model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
        prev, den = _run_affine(a, model_output, sample, noise)
        if not return_dict:
            return (prev, den)
        return LCMSchedulerOutput(prev_sample=prev, denoised=den)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        a = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)[timesteps.to(original_samples.device)]
        sa, sb = (a ** 0.5).flatten(), ((1 - a) ** 0.5).flatten()
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def __len__(self):
        return self.config.num_train_timesteps
