"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (TMElyralab/MuseV at /root/reference) on CPU.

Only usable in the build container (where /root/reference is mounted); it is used to pin the CPU restatement in
oracle/unet3d_oracle.py and to generate the golden vectors under tests/golden/ (oracle/make_golden.py).
Nothing in the product path (musev_b200/) may import this module.

The reference targets diffusers 0.24 / huggingface_hub <1.0 / xformers 0.0.21; the shims below only stand in
for symbols that moved or for packages that are absent offline:
  * huggingface_hub.constants.hf_cache_home, HfFolder, cached_download  (diffusers/utils/constants.py:17,
    dynamic_modules_utils.py:28)
  * xformers.ops.memory_efficient_attention -> torch SDPA (same math on (B*H, N, d) tensors), installed AFTER
    `import diffusers` so that diffusers itself sees xformers as unavailable
  * mmcm.utils.gpu_util.get_gpu_status (imported, never called: musev/models/temporal_transformer.py:35)
"""
from __future__ import annotations

import importlib.machinery
import logging
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "musev"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = False


def load():
    """Returns (UNet3DConditionModel, DDIMScheduler) classes of the reference."""
    global _loaded
    import torch

    if not available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    if not _loaded:
        import huggingface_hub
        import huggingface_hub.constants as hc

        if not hasattr(hc, "hf_cache_home"):
            hc.hf_cache_home = os.path.expanduser("~/.cache/huggingface")
        for n in ("HfFolder", "cached_download"):
            if not hasattr(huggingface_hub, n):
                setattr(huggingface_hub, n, type(n, (), {}))
        import transformers.utils as tu

        if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
            tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, "diffusers", "src"))
        import diffusers  # noqa: F401  (0.24.0.dev0 fork) -- must be imported before xformers is stubbed

        def mea(q, k, v, attn_bias=None, op=None, scale=None, p=0.0):
            return torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=attn_bias, scale=scale)

        _stub("xformers").ops = _stub("xformers.ops", memory_efficient_attention=mea)
        _stub("mmcm")
        _stub("mmcm.utils")
        _stub("mmcm.utils.gpu_util", get_gpu_status=lambda *a, **k: None)
        sys.path.insert(0, REFERENCE_ROOT)
        import musev  # noqa: F401

        logging.getLogger("musev").setLevel(logging.WARNING)
        _loaded = True
    from musev.models.unet_3d_condition import UNet3DConditionModel
    from musev.schedulers import DDIMScheduler

    return UNet3DConditionModel, DDIMScheduler


# the two released configurations (musev/models/unet_loader.py:232-268) on top of SD-1.5's unet/config.json
SD15_KW = dict(
    sample_size=64, in_channels=4, out_channels=4, cross_attention_dim=768, attention_head_dim=8,
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, norm_num_groups=32,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
)
PRESET_KW = {
    "musev": dict(
        need_t2i_ip_adapter=True, need_adain_temporal_cond=True,
        t2i_ip_adapter_attn_processor="NonParamReferenceIPXFormersAttnProcessor",
    ),
    "musev_referencenet": dict(
        need_transformer_in=False, use_anivv1_cfg=True, resnet_2d_skip_time_act=True,
        need_t2i_ip_adapter=True, need_adain_temporal_cond=True, keep_vision_condtion=True,
        t2i_ip_adapter_attn_processor="NonParamReferenceIPXFormersAttnProcessor",
        need_refer_emb=True, ip_adapter_cross_attn=True,
        t2i_crossattn_ip_adapter_attn_processor="T2IReferencenetIPAdapterXFormersAttnProcessor",
    ),
}
