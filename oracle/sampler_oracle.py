"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two other samplers of SURVEY.md 8(f)-4, written in the reference's own
(non-affine) form so that it is an independent check of musev_b200/samplers.py.

  * EulerOracle   musev/schedulers/scheduling_euler_discrete.py:47-170 over diffusers
                  schedulers/scheduling_euler_discrete.py:135-330. Pinned by the upstream known-answer tests
                  diffusers/tests/schedulers/test_scheduler_euler.py:41-110 (full loop 10.0807 / 0.0131, v-prediction
                  0.0002 / 2.2676e-06) and by sigmas / timesteps / steps of the imported musev scheduler
                  (tests/golden/samplers_sd15.pt, oracle/make_golden.py).
  * LCMOracle     musev/schedulers/scheduling_lcm.py:196-312 over diffusers schedulers/scheduling_lcm.py:196-470. Pinned by
                  diffusers/tests/schedulers/test_scheduler_lcm.py:226-244 (one step 18.7097 / 0.0244, ten steps
                  197.7616 / 0.2575) and the same fixture.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


def _betas(schedule, b0, b1, n):
    if schedule == "linear":
        return torch.linspace(b0, b1, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(schedule)


class EulerOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", timestep_spacing="linspace", steps_offset=0):
        self.n = num_train_timesteps
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), dim=0)
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.sigmas = torch.from_numpy(np.concatenate([sig[::-1], [0.0]]).astype(np.float32))
        self.step_index = None

    @property
    def init_noise_sigma(self):
        if self.timestep_spacing in ("linspace", "trailing"):
            return self.sigmas.max()
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    def set_timesteps(self, n):
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, self.n - 1, n, dtype=np.float32)[::-1].copy()
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (self.n // n)).round()[::-1].copy().astype(np.float32) + self.steps_offset
        else:
            ts = (np.arange(self.n, 0, -self.n / n)).round().copy().astype(np.float32) - 1
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.step_index = None

    def _index(self, t):
        if self.step_index is None:
            c = (self.timesteps == t).nonzero()
            self.step_index = (c[1] if len(c) > 1 else c[0]).item()

    def scale_model_input(self, x, t):
        self._index(t)
        sigma = self.sigmas[self.step_index]
        return x / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, t, sample):
        self._index(t)
        sigma = self.sigmas[self.step_index]
        if self.prediction_type == "epsilon":
            x0 = sample - sigma * model_output
        elif self.prediction_type == "v_prediction":
            x0 = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        else:
            x0 = model_output
        derivative = (sample - x0) / sigma
        dt = self.sigmas[self.step_index + 1] - sigma
        self.step_index += 1
        return sample + derivative * dt, x0


class LCMOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 original_inference_steps=50, set_alpha_to_one=True, prediction_type="epsilon", timestep_scaling=10.0):
        self.n = num_train_timesteps
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.original_inference_steps = original_inference_steps
        self.prediction_type = prediction_type
        self.timestep_scaling = timestep_scaling
        self.step_index = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        k = self.n // self.original_inference_steps
        origin = (np.asarray(list(range(1, self.original_inference_steps + 1))) * k - 1)[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=n, endpoint=False)).astype(np.int64)
        self.timesteps = torch.from_numpy(origin[idx]).long()
        self.step_index = None

    def step(self, model_output, t, sample, generator=None):
        if self.step_index is None:
            c = (self.timesteps == t).nonzero()
            self.step_index = (c[1] if len(c) > 1 else c[0]).item()
        nxt = self.step_index + 1
        prev_t = self.timesteps[nxt] if nxt < len(self.timesteps) else t
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        st = t * self.timestep_scaling
        c_skip = 0.5 ** 2 / (st ** 2 + 0.5 ** 2)
        c_out = st / (st ** 2 + 0.5 ** 2) ** 0.5
        if self.prediction_type == "epsilon":
            x0 = (sample - (1 - a_t).sqrt() * model_output) / a_t.sqrt()
        elif self.prediction_type == "v_prediction":
            x0 = a_t.sqrt() * sample - (1 - a_t).sqrt() * model_output
        else:
            x0 = model_output
        denoised = c_out * x0 + c_skip * sample
        if self.step_index != self.num_inference_steps - 1:
            noise = torch.randn(model_output.shape, generator=generator, dtype=denoised.dtype)
            prev = a_p.sqrt() * denoised + (1 - a_p).sqrt() * noise
        else:
            noise, prev = None, denoised
        self.step_index += 1
        return prev, denoised, noise
