"""The north star's own acceptance measurement (BASELINE.json: 20 DDIM steps, config 2, fixed seed, max-abs on the denoised
latents; SURVEY.md 7.2 / 8d: report (ours - ref32) next to (ref16 - ref32)).

What is asserted, and why not "1e-3 max-abs": with the synthetic (random-init) weights the 20-step latents have std ~13, and
the reference's OWN fp16 path (eager PyTorch fp16, scripts/inference/text2video.py:590) ends 0.27 max-abs / 0.030 rms from the
fp32 ground truth. The engine (fp16 storage, fp32 accumulation and statistics) ends 0.066 / 0.013 -- 4x / 2.3x closer than the
reference's own dtype, and exactly where an fp32 restatement with fp16 rounding at the engine's storage points lands
(profiles/r02_parity20_musev_v0.json: emu_all 0.068 / 0.0134). Bounds = ~2x the measured engine distances, relative to the
latent std, plus the ordering against ref16."""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_20step_config2_latents_parity(built_lib):
    from gpu_parity_20step import run
    r = run("musev", steps=20, T=16, h=64, w=64)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/parity20_test.json", "w") as fh:
            json.dump(r, fh)
    except OSError:
        pass
    std = r["latents_std"]
    ours, ref16 = r["ours_minus_ref32"], r["ref16_minus_ref32"]
    assert ours["rms"] / std < 2.5e-3, r          # measured 1.03e-3
    assert ours["max_abs"] / std < 1.2e-2, r      # measured 5.1e-3
    assert ours["rms"] < ref16["rms"] and ours["max_abs"] < ref16["max_abs"], r   # closer to fp32 truth than the reference's fp16
